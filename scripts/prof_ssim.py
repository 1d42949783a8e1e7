"""SSIM forward / backward and L1+TV at B=16, 3x256x256 (for ncu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200.train_step import gaussian_window
be = CudaBackend()
B, H, W = 16, 256, 256
f = torch.tanh(torch.randn(B, 3, H, W, device="cuda")); r = torch.rand(B, 3, H, W, device="cuda") * 2 - 1
sums = torch.zeros(8 + B, device="cuda"); d = torch.empty_like(f)
win = gaussian_window(); ga, gb, gc = (torch.empty_like(f) for _ in range(3))
for _ in range(3):
    be.ssim_fwd(f, r, .5, .5, win, sums[8:], ga, gb, gc)
    be.ssim_bwd(f, r, .5, .5, win, ga, gb, gc, 1.0, d, True)
    be.pixel_loss(f, r, 1.0, 1.0, 1.0, sums[:3], d)
torch.cuda.synchronize(); print("ok")
