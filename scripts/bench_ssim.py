"""SSIM forward / backward at the train-step size (B=16, 3x256x256): CUDA-event time with a flushed L2, both kernel families
(IRC_SSIM=tiled selects the older 32 x 16 tiles)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200.train_step import gaussian_window
be = CudaBackend()
B, H, W = 16, 256, 256
f = torch.tanh(torch.randn(B, 3, H, W, device="cuda")); r = torch.rand(B, 3, H, W, device="cuda") * 2 - 1
sums = torch.zeros(B, device="cuda"); d = torch.empty_like(f)
win = gaussian_window(); ga, gb, gc = (torch.empty_like(f) for _ in range(3))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[3]
npc = B * 3 * H * W
a = t(lambda: be.ssim_fwd(f, r, .5, .5, win, sums, ga, gb, gc))
b = t(lambda: be.ssim_bwd(f, r, .5, .5, win, ga, gb, gc, 1.0, d, True))
print(f"mode={os.environ.get('IRC_SSIM', 'default')}  ssim_fwd {a:6.1f} us ({npc * 110 / a / 1e6:5.2f} TFMA/s, {npc * 4 * 5 / a / 1e3:6.0f} GB/s)   ssim_bwd {b:6.1f} us ({npc * 66 / b / 1e6:5.2f} TFMA/s, {npc * 4 * 7 / b / 1e3:6.0f} GB/s)")
