"""ncu target: the L2-resident single-launch InstanceNorm backward at the full-resolution shapes (B=16, 64 / 128 ch, 256 x 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import layout as L
be = CudaBackend(); B = 16
def F(h, w, p, c):
    f = L.Frame(B, h, w, p, c, "cuda"); f.t.normal_(); return f
flush = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda")
for C in (64, 128):
    Z, G, dZ = F(256, 256, 1, C), F(256, 256, 1, C), F(256, 256, 1, C)
    st = torch.rand(B, C, 2, device="cuda") + 1.0; bs = torch.zeros(B, 512, 2, device="cuda")
    be.in_stats(Z.view(), C, B, 256, 256, st)
    for _ in range(2):
        flush.sum()
        be.in_bwd(Z.view(), G.view(), dZ.view(), C, B, 256, 256, stats=st, cnt=65536, act=1, bsum=bs)
torch.cuda.synchronize(); print("ok")
