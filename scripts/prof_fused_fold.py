"""in_bwd_fused with and without the reflection fold at the ResNet-bottleneck shape, for ncu (source-level stall samples)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import layout as L
be = CudaBackend()
B = 16
def F(h, w, p, c):
    f = L.Frame(B, h, w, p, c, "cuda"); f.t.normal_(); return f
Z, G, dZ = (F(64, 64, 1, 256) for _ in range(3))
st = torch.rand(B, 256, 2, device="cuda") + 1.0; st[..., 1] += 4096; bs = torch.zeros(B, 512, 2, device="cuda")
flush = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.sum()
    be.in_bwd(Z.view(), G.view(), dZ.view(), 256, B, 64, 64, stats=st, cnt=4096, act=1, bsum=bs, fold_pad=1)
    flush.sum()
    be.in_bwd(Z.view(), G.view(), dZ.view(), 256, B, 64, 64, stats=st, cnt=4096, act=1, bsum=bs, fold_pad=0)
torch.cuda.synchronize(); print("ok")
