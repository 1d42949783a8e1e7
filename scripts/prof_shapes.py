"""conv_gemm forward launches of chosen plan shapes (operands, epilogue arguments and frame strides as in the train step) with
CUDA-event timing - used to A/B two builds of the library (the script only needs the stable `be.conv_gemm` signature)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import layout as L
be = CudaBackend(); dev = "cuda"
tag = sys.argv[1] if len(sys.argv) > 1 else ""
#        name         n   H    cin  frame_C off cout  epilogue
SHAPES = [("V.7 fwd", 32, 128, 128, 128, 0, 128, "bias"),
          ("V.5 fwd", 32, 128, 64, 64, 0, 128, "bias"),
          ("V.2 fwd", 32, 256, 64, 64, 0, 64, "bias"),
          ("G.down2 fwd", 16, 128, 128, 384, 256, 256, "stats"),
          ("G.down1 fwd", 16, 256, 64, 192, 128, 128, "stats"),
          ("G.up1 fwd", 16, 128, 384, 384, 0, 128, "stats"),
          ("V.12 fwd", 32, 64, 256, 256, 0, 256, "bias"),
          ("G.res fwd", 16, 64, 256, 256, 0, 256, "none")]
for name, n, H, cin, fc, off, cout, epi in SHAPES:
    fr = L.Frame(n, H, H, 1, fc, dev); fr.t.normal_()
    out = L.Frame(n, H, H, 1, cout, dev)
    w = (torch.randn(cout, 9 * cin, device=dev) * 0.02).bfloat16()
    bias = torch.randn(cout, device=dev)
    ri = torch.zeros(fr.rows, device=dev, dtype=torch.int16)
    be.row_index(ri, n, H + 2, H + 2, 1, H + 1, 1, H + 1)
    st = torch.zeros(n, cout, 2, device=dev)
    taps = L.taps_centered(3, 3, fr.wp)
    kw = dict(bias=bias, act=1, row_img=ri) if epi == "bias" else (dict(row_img=ri, in_stats=(st, n, fr.hp * fr.wp)) if epi == "stats" else {})
    flops = 2.0 * n * H * H * cout * 9 * cin
    for _ in range(3):
        be.conv_gemm(fr.t, off, cin, taps, w, cout, out.t, **kw)
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            be.conv_gemm(fr.t, off, cin, taps, w, cout, out.t, **kw)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 10)
    ms = sorted(ts)[1]
    print(f"{tag:12s} {name:12s} {ms * 1e3:7.1f} us  {flops / ms / 1e9:6.0f} TFLOP/s  ({flops / ms / 1e9 / 1370:.3f})", flush=True)
    del fr, out
