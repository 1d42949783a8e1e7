"""The ten stencil / apply launches (`irc_gather`) of one generator forward + backward at B=16, 256x256, on the plan's own frames:
CUDA-event time per launch with a flushed L2, algorithmic MB and GB/s.  Also the ncu target for these kernels
(`ncu --set full -k regex:gather ... python scripts/prof_gathers.py 1`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import engine as E
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
be = CudaBackend(); dev = "cuda"
B, H, W = 16, 256, 256
H2, W2, H4, W4 = H // 2, W // 2, H // 4, W // 4
g = E.GeneratorEngine(be, B, H, W, dev)
for fr in (g.Z0, g.Z1, g.Z2, g.Z3, g.Z4, g.X[g.nb], g.Gcat1, g.Gcat2, g.Gx1, g.dOut[0]):
    fr.t.normal_()
for st in (g.st0, g.st1, g.st2, g.st3, g.st4):
    st[..., 0].normal_(); st[..., 1].fill_(3e4)
cur = g.dOut[0]
MB = lambda *elems: sum(elems) * 2 / 1e6
px = B * H * W
calls = [
    ("inc IN+ReLU -> cat2[128:192) (64 ch, identity)", MB(px * 64, px * 64),
     lambda: be.gather(g.Z0.view(), g.cat2.view(128), 64, B, H, W, 1, 0, **g._na(g.st0, H * W))),
    ("down1 IN+ReLU+Downsample (128 ch, 256^2 -> 128^2)", MB(px * 128, px * 32),
     lambda: be.gather(g.Z1.view(), g.cat1.view(256), 128, B, H2, W2, 1, 0, tables=g.t_down1, **g._na(g.st1, H * W))),
    ("down2 IN+ReLU+Downsample (256 ch, 128^2 -> 64^2, reflect ring)", MB(px * 64, px * 16),
     lambda: be.gather(g.Z2.view(), g.X[0].view(), 256, B, H4, W4, 1, 1, tables=g.t_down2, **g._na(g.st2, H2 * W2))),
    ("up1 UpsampleAA (256 ch, 64^2 -> 128^2)", MB(px * 16, px * 64),
     lambda: be.gather(g.X[g.nb].view(), g.cat1.view(0), 256, B, H2, W2, 1, 0, tables=g.t_up1)),
    ("up2 IN+ReLU+UpsampleAA (128 ch, 128^2 -> 256^2)", MB(px * 32, px * 128),
     lambda: be.gather(g.Z3.view(), g.cat2.view(0), 128, B, H, W, 1, 0, tables=g.t_up2, **g._na(g.st3, H2 * W2))),
    ("up2 conv IN+ReLU -> y4 (64 ch, reflect-3 ring, identity)", MB(px * 64, px * 64),
     lambda: be.gather(g.Z4.view(), g.y4.view(), 64, B, H, W, 3, 1, **g._na(g.st4, H * W))),
    ("UpsampleAA^T up2 (128 ch, 256^2 -> 128^2)", MB(px * 128, px * 32),
     lambda: be.gather(g.Gcat2.view(0), g.g3.view(), 128, B, H2, W2, 0, 0, tables=g.t_up2_T)),
    ("UpsampleAA^T up1 (256 ch, 128^2 -> 64^2)", MB(px * 64, px * 16),
     lambda: be.gather(g.Gcat1.view(0), cur.view(), 256, B, H4, W4, 1, 0, tables=g.t_up1_T)),
    ("Downsample^T down2 (256 ch, 64^2 -> 128^2)", MB(px * 16, px * 64),
     lambda: be.gather(cur.view(), g.g2.view(), 256, B, H2, W2, 0, 0, tables=g.t_down2_T)),
    ("Downsample^T down1, two sources (128 ch, 128^2 -> 256^2)", MB(px * 32, px * 32, px * 128),
     lambda: be.gather(g.Gcat1.view(256), g.g1.view(), 128, B, H, W, 0, 0, tables=g.t_down1_T, src2=g.Gx1.view())),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = 6544.0
tot = 0.0
for i, (name, mb, fn) in enumerate(calls):
    if only and str(i) not in only:
        continue
    fn(); torch.cuda.synchronize()
    best = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(best)[len(best) // 2]
    tot += us
    print(f"[{i}] {name:66s} {mb:7.1f} MB {us:7.1f} us {mb / us * 1e3:7.0f} GB/s  ({mb / us * 1e3 / peak:.2f} of the HBM copy peak)")
print(f"total {tot:.1f} us")
