"""The memory-bound kernels where they are truly HBM-bound: BASELINE config-5 sized tensors (64 x 3 x 512 x 640 fp32 = 252 MB)
for the losses / metrics, and B=16 full-resolution feature maps (128 ch x 256 x 256) for the stencils.  CUDA events, L2
evicted with clean lines between repetitions, median of 5.  Writes one JSON document (--out)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import layout as L
from irc_b200.train_step import gaussian_window

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); args = ap.parse_args()
be = CudaBackend()
dev = "cuda"
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6544.0
flush = torch.zeros(512 << 20, dtype=torch.uint8, device=dev)
rows = []


def timeit(name, fn, nbytes, note="", reps=5, fma=None):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    r = dict(kernel=name, us=round(t * 1e3, 1), algorithmic_mb=round(nbytes / 1e6, 1), gb_s=round(nbytes / t / 1e6), frac_of_hbm_peak=round(nbytes / t / 1e6 / PEAK, 3), note=note)
    if fma:
        r["tfma_s"] = round(fma / t / 1e9, 2)
    rows.append(r)
    print(f"{name:58s} {t * 1e3:9.1f} us {nbytes / t / 1e6:7.0f} GB/s ({nbytes / t / 1e6 / PEAK:.2f})", flush=True)


# ---- losses / metrics at the config-5 size
N, C, H, W = 64, 3, 512, 640
f = torch.tanh(torch.randn(N, C, H, W, device=dev)); r = torch.rand(N, C, H, W, device=dev) * 2 - 1
sums = torch.zeros(8 + N, device=dev); d = torch.empty_like(f)
nb = f.numel() * 4
timeit("tv_loss + L1, value only (2 reads)", lambda: be.pixel_loss(f, r, 1.0, 1.0, 1.0, sums[:3], None), 2 * nb, "irc:686-694, :1664")
timeit("tv_loss + L1, value and gradient (2 reads + 1 write)", lambda: be.pixel_loss(f, r, 1.0, 1.0, 1.0, sums[:3], d), 3 * nb)
timeit("tv_loss alone, value and gradient (1 read + 1 write)", lambda: be.pixel_loss(f, None, 0.0, 1.0, 1.0, sums[:3], d), 2 * nb)
win = gaussian_window()
ga, gb, gc = (torch.empty_like(f) for _ in range(3))
px = f.numel()
timeit("ssim forward, loss value only (2 reads)", lambda: be.ssim_fwd(f, r, .5, .5, win, sums[8:]), 2 * nb, "FP32-pipe bound: 110 FMA per pixel-channel", fma=110 * px)
timeit("ssim forward + 3 saved maps (2 reads + 3 writes)", lambda: be.ssim_fwd(f, r, .5, .5, win, sums[8:], ga, gb, gc), 5 * nb, fma=110 * px)
timeit("ssim backward (5 reads + 1 write)", lambda: be.ssim_bwd(f, r, .5, .5, win, ga, gb, gc, 1.0, d, False), 6 * nb, fma=66 * px)
u8 = torch.empty(N, H, W, C, device=dev, dtype=torch.uint8); ms = torch.zeros(N, 2, device=dev, dtype=torch.float64)
timeit("quantise + MAE/MSE (2 reads + u8 write)", lambda: be.quantize_metrics(f, r, u8, ms), 2 * nb + f.numel(), "irc:865-876, :1197-1205")
del ga, gb, gc, d
# ---- fp32 NCHW stand-alone stencils (the drop-in Downsample / UpsampleAA modules) at B=16, 128 ch, 256^2 <-> 128^2
B, Cc, Hh, Ww = 16, 128, 256, 256
x = torch.randn(B, Cc, Hh, Ww, device=dev); o = torch.empty(B, Cc, Hh // 2, Ww // 2, device=dev)
tdn = L.make_tables(L.down_matrix(Hh), L.down_matrix(Ww), dev); tup = L.make_tables(L.up_matrix(Hh // 2), L.up_matrix(Ww // 2), dev)
tdnT = L.make_tables(L.down_matrix(Hh).T, L.down_matrix(Ww).T, dev); tupT = L.make_tables(L.up_matrix(Hh // 2).T, L.up_matrix(Ww // 2).T, dev)
timeit("Downsample fp32 NCHW 128ch 256^2 -> 128^2", lambda: be.stencil_nchw(x, o, tdn), (x.numel() + o.numel()) * 4, "irc:307-310")
timeit("Downsample^T fp32 NCHW (backward)", lambda: be.stencil_nchw(o, x, tdnT), (x.numel() + o.numel()) * 4)
timeit("UpsampleAA fp32 NCHW 128ch 128^2 -> 256^2", lambda: be.stencil_nchw(o, x, tup), (x.numel() + o.numel()) * 4, "irc:350-355")
timeit("UpsampleAA^T fp32 NCHW (backward)", lambda: be.stencil_nchw(x, o, tupT), (x.numel() + o.numel()) * 4)
del x, o
# ---- the fused bf16 NHWC stencils of the train step at the same shapes
def F(h, w, p, c):
    fr = L.Frame(B, h, w, p, c, dev); fr.t.normal_(); return fr
st = torch.rand(B, 128, 2, device=dev) + 1.0
Z1, cat1 = F(Hh, Ww, 1, 128), F(Hh // 2, Ww // 2, 1, 384)
be.in_stats(Z1.view(), 128, B, Hh, Ww, st)
n = B * Hh * Ww
timeit("IN + ReLU + Downsample, bf16 NHWC 128ch (fused, train step)", lambda: be.gather(Z1.view(), cat1.view(256), 128, B, Hh // 2, Ww // 2, 1, 0, tables=tdn, stats=st, cnt=Hh * Ww, act=1), (n + n // 4) * 128 * 2)
Z3, cat2 = F(Hh // 2, Ww // 2, 1, 128), F(Hh, Ww, 1, 192)
be.in_stats(Z3.view(), 128, B, Hh // 2, Ww // 2, st)
timeit("IN + ReLU + UpsampleAA, bf16 NHWC 128ch (fused, train step)", lambda: be.gather(Z3.view(), cat2.view(0), 128, B, Hh, Ww, 1, 0, tables=tup, stats=st, cnt=Hh * Ww // 4, act=1), (n + n // 4) * 128 * 2)
g3 = F(Hh // 2, Ww // 2, 0, 128)
timeit("UpsampleAA^T, bf16 NHWC 128ch (train step backward)", lambda: be.gather(cat2.view(0), g3.view(), 128, B, Hh // 2, Ww // 2, 0, 0, tables=tupT), (n + n // 4) * 128 * 2)
Gx1, g1 = F(Hh // 2, Ww // 2, 1, 128), F(Hh, Ww, 0, 128)
timeit("Downsample^T, two sources, bf16 NHWC 128ch (backward)", lambda: be.gather(cat1.view(256), g1.view(), 128, B, Hh, Ww, 0, 0, tables=tdnT, src2=Gx1.view()), (n + 2 * (n // 4)) * 128 * 2)
if args.out:
    json.dump(dict(hbm_peak_gb_s=PEAK, how="CUDA events, clean-line L2 eviction between repetitions, median of 5", kernels=rows), open(args.out, "w"), indent=1)
