"""Direct small-K convolutions at the bench shapes (VGG conv1_1: 16 x 3 x 256 x 256 -> framed 64 ch; inc: 16 x 1 x 256 x 256, 7 x 7 reflect;
D model.0: 16 x (1 + 3) x 256 x 256, 4 x 4 stride 2, space-to-depth row order) with CUDA-event timing; for ncu --set full."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend

be = CudaBackend()
n, H, W = 16, 256, 256
g = torch.Generator(device="cuda").manual_seed(0)
x1 = torch.randn(n, 1, H, W, device="cuda", generator=g)
x3 = torch.randn(n, 3, H, W, device="cuda", generator=g)
w = (torch.randn(64, 64, device="cuda", generator=g) * 0.1).bfloat16()
bias = torch.randn(64, device="cuda", generator=g)
sc = torch.ones(3, device="cuda"); sh = torch.zeros(3, device="cuda")
cases = {
    "v0": lambda out, E: be.smallk_conv_fwd(x3, None, sc, sh, n, H, W, 3, 1, 1, 0, H, W, 1, w, out, bias=bias, act=1, E=E),
    "inc": lambda out, E: be.smallk_conv_fwd(x1, None, None, None, n, H, W, 7, 1, 3, 1, H, W, 0, w, out, E=E),
    "d0": lambda out, E: be.smallk_conv_fwd(x1, x3, None, None, n, H, W, 4, 2, 1, 0, H // 2, W // 2, 2, w, out, bias=bias, act=2, slope=0.2, E=E),
}
rows = {"v0": be.im2col_rows(1, n, H, W), "inc": be.im2col_rows(0, n, H, W), "d0": be.im2col_rows(2, n, H // 2, W // 2)}
which = sys.argv[1:] or list(cases)
for name in which:
    for keep in (0, 1):
        out = torch.zeros(rows[name], 64, device="cuda", dtype=torch.bfloat16)
        E = torch.zeros(rows[name], 64, device="cuda", dtype=torch.bfloat16) if keep else None
        for _ in range(2):
            cases[name](out, E)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            cases[name](out, E)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        mb = rows[name] * 128 * (1 + keep) / 1e6
        print(f"{name} keep_operand={keep}: {us:7.1f} us  ({mb:.0f} MB written, {mb / us * 1e3:.0f} GB/s)", flush=True)
