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
cat2 = F(H, W, 1, 192); g3 = F(H // 2, W // 2, 0, 128)
tupT = L.make_tables(L.up_matrix(H // 2).T, L.up_matrix(W // 2).T, "cuda")
Z3 = F(H // 2, W // 2, 1, 128); st = torch.rand(B, 128, 2, device="cuda") + 1
tup = L.make_tables(L.up_matrix(H // 2), L.up_matrix(W // 2), "cuda")
for _ in range(3):
    be.gather(cat2.view(0), g3.view(), 128, B, H // 2, W // 2, 0, 0, tables=tupT)
    be.gather(Z3.view(), cat2.view(0), 128, B, H, W, 1, 0, tables=tup, stats=st, cnt=H * W // 4, act=1)
torch.cuda.synchronize(); print("ok")
