"""ncu target: the loss / metric kernels at the config-5 size (64 x 3 x 512 x 640 fp32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200.train_step import gaussian_window
be = CudaBackend()
N, C, H, W = 64, 3, 512, 640
f = torch.tanh(torch.randn(N, C, H, W, device="cuda")); r = torch.rand(N, C, H, W, device="cuda") * 2 - 1
sums = torch.zeros(8 + N, device="cuda"); d = torch.empty_like(f)
win = gaussian_window()
ga, gb, gc = (torch.empty_like(f) for _ in range(3))
u8 = torch.empty(N, H, W, C, device="cuda", dtype=torch.uint8); ms = torch.zeros(N, 2, device="cuda", dtype=torch.float64)
for _ in range(2):
    be.pixel_loss(f, r, 1.0, 1.0, 1.0, sums[:3], d)
    be.ssim_fwd(f, r, .5, .5, win, sums[8:], ga, gb, gc)
    be.ssim_bwd(f, r, .5, .5, win, ga, gb, gc, 1.0, d, False)
    be.quantize_metrics(f, r, u8, ms)
torch.cuda.synchronize()
