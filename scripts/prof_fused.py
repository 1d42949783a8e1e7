"""cluster-resident InstanceNorm forward / backward at the ResNet-bottleneck shape (B=16, 64x64, 256 ch), for ncu"""
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
Z, G, dZ, X, Y = (F(64, 64, 1, 256) for _ in range(5))
st = torch.zeros(B, 256, 2, device="cuda"); bs = torch.zeros(B, 512, 2, device="cuda")
flush = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.sum()
    be.in_apply(Z.view(), Y.view(), 256, B, 64, 64, 1, 1, st, act=0, res=X.view())
    flush.sum()
    be.in_bwd(Z.view(), G.view(), dZ.view(), 256, B, 64, 64, stats=st, cnt=4096, act=1, bsum=bs, fold_pad=1)
torch.cuda.synchronize(); print("ok")
