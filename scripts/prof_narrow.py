"""The N = 64 full-resolution conv_gemm shapes (VGG conv1_2 forward: 32 x 258 x 258 rows, 64 -> 64, 3 x 3; up2_conv forward:
16 x 258 x 258 rows, 192 -> 64) with CUDA-event timing per variant; `ncu --set full` captures the last launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import layout as L

be = CudaBackend()
which = sys.argv[1] if len(sys.argv) > 1 else "v2"
n_img, cin = (32, 64) if which == "v2" else (16, 192)
H = W = 256
fr = L.Frame(n_img, H, W, 1, cin, "cuda"); fr.t.normal_()
out = L.Frame(n_img, H, W, 1, 64, "cuda")
w = (torch.randn(64, 9 * cin, device="cuda") * 0.02).bfloat16()
bias = torch.randn(64, device="cuda")
ri = torch.zeros(fr.rows, device="cuda", dtype=torch.int16)
be.row_index(ri, n_img, H + 2, W + 2, 1, H + 1, 1, W + 1)
taps = L.taps_centered(3, 3, fr.wp)
flops = 2.0 * n_img * H * W * 64 * 9 * cin


def run(reuse, mt, epi, n=10):
    be.conv_reuse, be.conv_mt = reuse, mt
    kw = dict(bias=bias, act=1, row_img=ri) if epi else {}
    for _ in range(2):
        be.conv_gemm(fr.t, 0, cin, taps, w, 64, out.t, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        be.conv_gemm(fr.t, 0, cin, taps, w, 64, out.t, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{which} reuse={reuse} mt={mt} epilogue={'bias+relu+ring' if epi else 'plain'}: {ms * 1e3:7.1f} us  {flops / ms / 1e9:6.0f} TFLOP/s", flush=True)


if len(sys.argv) > 2:      # one configuration: reuse mt epilogue  (for ncu)
    run(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), n=3)
else:
    for epi in (1, 0):
        for reuse in (0, 1):
            for mt in (1, 2, 4):
                run(reuse, mt, epi)
        run(2, 0, epi)          # packed taps
