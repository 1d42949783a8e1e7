"""One resblock-shaped conv_gemm (M=16*66*66, N=256, K=9*256) and one tn_gemm launch loop for ncu --set full."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import layout as L

be = CudaBackend()
B, H, W, C = 16, 64, 64, 256
fr = L.Frame(B, H, W, 1, C, "cuda"); fr.t.normal_()
out = L.Frame(B, H, W, 1, C, "cuda")
w = (torch.randn(C, 9 * C, device="cuda") * 0.02).bfloat16()
taps = L.taps_centered(3, 3, fr.wp)
part = torch.zeros(16 * C * 9 * C, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(n):
    be.conv_gemm(fr.t, 0, C, taps, w, C, out.t)
    be.tn_gemm(out.t, 0, C, fr.t, 0, C, fr.rows, [0] * 9, taps, part, C, 9 * C, 1, 16, C * 9 * C)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    be.conv_gemm(fr.t, 0, C, taps, w, C, out.t)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("conv_gemm resblock: %.1f us, %.0f TFLOP/s" % (ms * 1e3, 2 * B * H * W * C * 9 * C / ms / 1e9))
