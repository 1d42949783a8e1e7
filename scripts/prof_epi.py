"""What the conv_gemm epilogue arguments cost at the VGG conv1_2 shape (32 x 258 x 258 rows, 64 -> 64) and conv2_2 (128 -> 128, 130 x 130)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend, View
from irc_b200 import layout as L
be = CudaBackend()
for (n_img, H, cin, cout) in ((32, 256, 64, 64), (32, 128, 128, 128)):
    fr = L.Frame(n_img, H, H, 1, cin, "cuda"); fr.t.normal_()
    out = L.Frame(n_img, H, H, 1, cout, "cuda")
    msk = L.Frame(n_img, H, H, 1, cout, "cuda"); msk.t.normal_()
    w = (torch.randn(cout, 9 * cin, device="cuda") * 0.02).bfloat16()
    bias = torch.randn(cout, device="cuda")
    ri = torch.zeros(fr.rows, device="cuda", dtype=torch.int16)
    be.row_index(ri, n_img, H + 2, H + 2, 1, H + 1, 1, H + 1)
    taps = L.taps_centered(3, 3, fr.wp)
    flops = 2.0 * n_img * H * H * cout * 9 * cin
    for name, kw in (("plain", {}), ("bias", dict(bias=bias)), ("bias+relu", dict(bias=bias, act=1)), ("row_img", dict(row_img=ri)),
                     ("bias+relu+row_img", dict(bias=bias, act=1, row_img=ri)), ("mask (dgrad)", dict(mask=View(msk.t, 0, 0, 0), mask_slope=0.0))):
        for _ in range(3):
            be.conv_gemm(fr.t, 0, cin, taps, w, cout, out.t, **kw)
        ts = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                be.conv_gemm(fr.t, 0, cin, taps, w, cout, out.t, **kw)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 10)
        ms = sorted(ts)[1]
        print(f"{cin:4d}->{cout:4d} {H}^2  {name:20s} {ms * 1e3:7.1f} us  {flops / ms / 1e9:6.0f} TFLOP/s", flush=True)
