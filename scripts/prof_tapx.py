"""outc backward expansion (tap_expand + bias gradient) at the bench shape, for ncu and CUDA-event timing"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
be = CudaBackend()
B, H, W, p = 16, 256, 256, 3
hp, wp = H + 2 * p, W + 2 * p
g = torch.randn(B, 3, H, W, device="cuda"); y = torch.tanh(torch.randn(B, 3, H, W, device="cuda"))
E = torch.zeros(B * hp * wp, 64, device="cuda", dtype=torch.bfloat16); db = torch.zeros(3, device="cuda")
sh = [(0, s - 3) for s in range(7)]
for live in (False, True):
    for _ in range(3):
        be.tap_expand(g, y, sh, 3, B, H, W, hp, wp, p, p, E, dbias=db, live_cols_only=live)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        be.tap_expand(g, y, sh, 3, B, H, W, hp, wp, p, p, E, dbias=db, live_cols_only=live)
    e1.record(); torch.cuda.synchronize()
    print(f"tap_expand live_cols_only={live}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
