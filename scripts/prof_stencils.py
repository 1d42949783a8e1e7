"""ncu target: the stencil kernels (fused bf16 NHWC of the train step, stand-alone fp32 NCHW modules) at B=16, 128 ch, 256^2 <-> 128^2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import layout as L
be = CudaBackend(); dev = "cuda"
B, Cc, Hh, Ww = 16, 128, 256, 256
tdn = L.make_tables(L.down_matrix(Hh), L.down_matrix(Ww), dev); tup = L.make_tables(L.up_matrix(Hh // 2), L.up_matrix(Ww // 2), dev)
tdnT = L.make_tables(L.down_matrix(Hh).T, L.down_matrix(Ww).T, dev); tupT = L.make_tables(L.up_matrix(Hh // 2).T, L.up_matrix(Ww // 2).T, dev)
def F(h, w, p, c):
    fr = L.Frame(B, h, w, p, c, dev); fr.t.normal_(); return fr
st = torch.rand(B, 128, 2, device=dev) + 1.0
Z1, cat1 = F(Hh, Ww, 1, 128), F(Hh // 2, Ww // 2, 1, 384)
Z3, cat2 = F(Hh // 2, Ww // 2, 1, 128), F(Hh, Ww, 1, 192)
g3 = F(Hh // 2, Ww // 2, 0, 128); Gx1, g1 = F(Hh // 2, Ww // 2, 1, 128), F(Hh, Ww, 0, 128)
be.in_stats(Z1.view(), 128, B, Hh, Ww, st)
x = torch.randn(B, Cc, Hh, Ww, device=dev); o = torch.empty(B, Cc, Hh // 2, Ww // 2, device=dev)
for _ in range(2):
    be.gather(Z1.view(), cat1.view(256), 128, B, Hh // 2, Ww // 2, 1, 0, tables=tdn, stats=st, cnt=Hh * Ww, act=1)
    be.gather(Z3.view(), cat2.view(0), 128, B, Hh, Ww, 1, 0, tables=tup, stats=st, cnt=Hh * Ww // 4, act=1)
    be.gather(cat2.view(0), g3.view(), 128, B, Hh // 2, Ww // 2, 0, 0, tables=tupT)
    be.gather(cat1.view(256), g1.view(), 128, B, Hh, Ww, 0, 0, tables=tdnT, src2=Gx1.view())
    be.stencil_nchw(x, o, tdn); be.stencil_nchw(o, x, tdnT); be.stencil_nchw(o, x, tup); be.stencil_nchw(x, o, tupT)
torch.cuda.synchronize()
