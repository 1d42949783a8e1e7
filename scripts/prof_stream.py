"""the four streaming stencil launches of the generator plan at B=16, 256x256 (for ncu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import layout as L
be = CudaBackend()
B, H, W = 16, 256, 256
def F(h, w, p, c):
    f = L.Frame(B, h, w, p, c, "cuda"); f.t.normal_(); return f
cat2 = F(H, W, 1, 192); g3 = F(H // 2, W // 2, 0, 128); Z1 = F(H, W, 1, 128); cat1 = F(H // 2, W // 2, 1, 384)
Gx1 = F(H // 2, W // 2, 1, 128); g1 = F(H, W, 0, 128)
Z3 = F(H // 2, W // 2, 1, 128); st = torch.rand(B, 128, 2, device="cuda") + 1
mk = lambda a, b: L.make_tables(a, b, "cuda")
tup, tupT = mk(L.up_matrix(H // 2), L.up_matrix(W // 2)), mk(L.up_matrix(H // 2).T, L.up_matrix(W // 2).T)
tdn, tdnT = mk(L.down_matrix(H), L.down_matrix(W)), mk(L.down_matrix(H).T, L.down_matrix(W).T)
for _ in range(2):
    be.gather(cat2.view(0), g3.view(), 128, B, H // 2, W // 2, 0, 0, tables=tupT)
    be.gather(Z3.view(), cat2.view(0), 128, B, H, W, 1, 0, tables=tup, stats=st, cnt=H * W // 4, act=1)
    be.gather(Z1.view(), cat1.view(256), 128, B, H // 2, W // 2, 1, 0, tables=tdn, stats=st, cnt=H * W, act=1)
    be.gather(cat1.view(256), g1.view(), 128, B, H, W, 0, 0, tables=tdnT, src2=Gx1.view())
torch.cuda.synchronize(); print("ok")
