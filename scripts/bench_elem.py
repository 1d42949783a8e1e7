"""Micro-benchmarks of the memory-bound kernels at the B=16, 256x256 shapes of the train step (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend, View
from irc_b200 import layout as L

be = CudaBackend()
B = 16
dev = "cuda"
flush = torch.zeros(512 << 20, dtype=torch.uint8, device=dev)


def timeit(name, fn, bytes_, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.sum()          # evict L2 with clean lines (a write flush would leave dirty lines the timed kernel has to write back)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:42s} {t * 1e3:8.1f} us  {bytes_ / t / 1e6:7.0f} GB/s  ({bytes_ / t / 1e6 / 6544:.2f} of HBM peak)")


def F(h, w, p, c):
    f = L.Frame(B, h, w, p, c, dev); f.t.normal_(); return f


st = lambda c: torch.rand(B, c, 2, device=dev) + 1.0
H = W = 256
bs = torch.zeros(B, 512, 2, device=dev)
# in_bwd identity, two sources (inc layer)
Z0, G1, G2, dZ0 = F(H, W, 0, 64), F(H, W, 1, 192), F(H, W, 1, 64), F(H, W, 0, 64)
s64 = st(64)
be.in_stats(Z0.view(), 64, B, H, W, s64)
n = B * H * W
timeit("in_bwd identity 2src 64ch 256^2", lambda: be.in_bwd(Z0.view(), G1.view(128), dZ0.view(), 64, B, H, W, stats=s64, cnt=H * W, act=1, g2=G2.view(), bsum=bs),
       n * 64 * 2 * (2 * 3 + 1))
# in_bwd fold3 (up2)
Z4, G4, dZ4 = F(H, W, 1, 64), F(H, W, 3, 64), F(H, W, 1, 64)
tf3 = L.make_tables(L.fold_matrix(H, 3), L.fold_matrix(W, 3), dev)
timeit("in_bwd fold3 64ch 256^2", lambda: be.in_bwd(Z4.view(), G4.pview(), dZ4.view(), 64, B, H, W, stats=s64, cnt=H * W, act=1, tables=tf3, bsum=bs),
       n * 64 * 2 * (2 * 2 + 1))
# in_bwd down^T two sources (down1)
Z1, Gc1, Gx1, dZ1 = F(H, W, 1, 128), F(H // 2, W // 2, 1, 384), F(H // 2, W // 2, 1, 128), F(H, W, 1, 128)
s128 = st(128); be.in_stats(Z1.view(), 128, B, H, W, s128)
tdT = L.make_tables(L.down_matrix(H).T, L.down_matrix(W).T, dev)
timeit("in_bwd down^T 2src 128ch 256^2", lambda: be.in_bwd(Z1.view(), Gc1.view(256), dZ1.view(), 128, B, H, W, stats=s128, cnt=H * W, act=1, tables=tdT, g2=Gx1.view(), bsum=bs),
       n * 128 * 2 * (2 * 1 + 1) + 2 * 2 * n // 4 * 128 * 2 * 2)
# forward fused IN+ReLU+upsample (up2)
Z3, cat2 = F(H // 2, W // 2, 1, 128), F(H, W, 1, 192)
be.in_stats(Z3.view(), 128, B, H // 2, W // 2, s128)
tup = L.make_tables(L.up_matrix(H // 2), L.up_matrix(W // 2), dev)
timeit("gather IN+ReLU+UpsampleAA 128ch ->256^2", lambda: be.gather(Z3.view(), cat2.view(0), 128, B, H, W, 1, 0, tables=tup, stats=s128, cnt=H * W // 4, act=1),
       n // 4 * 128 * 2 + n * 128 * 2)
# forward fused IN+ReLU+downsample (down1)
cat1 = F(H // 2, W // 2, 1, 384)
tdn = L.make_tables(L.down_matrix(H), L.down_matrix(W), dev)
timeit("gather IN+ReLU+Downsample 128ch 256^2->", lambda: be.gather(Z1.view(), cat1.view(256), 128, B, H // 2, W // 2, 1, 0, tables=tdn, stats=s128, cnt=H * W, act=1),
       n * 128 * 2 + n // 4 * 128 * 2)
# plain apply (y4) and resblock apply
y4 = F(H, W, 3, 64)
timeit("gather IN+ReLU apply 64ch 256^2 reflect3", lambda: be.gather(Z4.view(), y4.view(), 64, B, H, W, 3, 1, stats=s64, cnt=H * W, act=1), n * 64 * 2 * 2)
X, Zb, Y = F(64, 64, 1, 256), F(64, 64, 1, 256), F(64, 64, 1, 256)
s256 = st(256); be.in_stats(Zb.view(), 256, B, 64, 64, s256)
timeit("gather IN+res apply 256ch 64^2 (resblock)", lambda: be.gather(Zb.view(), Y.view(), 256, B, 64, 64, 1, 1, stats=s256, cnt=4096, act=0, res=X.view()), B * 4096 * 256 * 2 * 3)
timeit("in_stats 256ch 64^2", lambda: be.in_stats(Zb.view(), 256, B, 64, 64, s256), B * 4096 * 256 * 2)
timeit("in_stats 128ch 256^2", lambda: be.in_stats(Z1.view(), 128, B, H, W, s128), n * 128 * 2)
dZb = F(64, 64, 1, 256)
timeit("in_bwd identity 256ch 64^2 (resblock)", lambda: be.in_bwd(Zb.view(), X.view(), dZb.view(), 256, B, 64, 64, stats=s256, cnt=4096, act=0, bsum=bs), B * 4096 * 256 * 2 * 5)
# transposed stencils (backward), materialised by the tiled gather
g1 = F(H, W, 0, 128)
timeit("gather down^T 2src 128ch ->256^2 ", lambda: be.gather(Gc1.view(256), g1.view(), 128, B, H, W, 0, 0, tables=tdT, src2=Gx1.view()),
       n * 128 * 2 + 2 * n // 4 * 128 * 2)
timeit("in_bwd identity 1src 128ch 256^2", lambda: be.in_bwd(Z1.view(), g1.view(), dZ1.view(), 128, B, H, W, stats=s128, cnt=H * W, act=1, bsum=bs),
       n * 128 * 2 * 5)
g3 = F(H // 2, W // 2, 0, 128)
tupT = L.make_tables(L.up_matrix(H // 2).T, L.up_matrix(W // 2).T, dev)
timeit("gather up^T 128ch 256^2->128^2 ", lambda: be.gather(cat2.view(0), g3.view(), 128, B, H // 2, W // 2, 0, 0, tables=tupT),
       n * 128 * 2 + n // 4 * 128 * 2)
g4 = F(H, W, 0, 64)
timeit("gather fold3 64ch 256^2 ", lambda: be.gather(G4.pview(), g4.view(), 64, B, H, W, 0, 0, tables=tf3), n * 64 * 2 * 2)
Gh = F(64, 64, 1, 256); dZa = F(64, 64, 1, 256)
tf1 = L.make_tables(L.fold_matrix(64, 1), L.fold_matrix(64, 1), dev)
timeit("in_bwd fold1 256ch 64^2 (resblock, fused)", lambda: be.in_bwd(Zb.view(), Gh.pview(), dZa.view(), 256, B, 64, 64, stats=s256, cnt=4096, act=1, tables=tf1, bsum=bs), B * 4096 * 256 * 2 * 5)
timeit("gather fold1+res 256ch 64^2 (resblock)", lambda: be.gather(Gh.pview(), Y.view(), 256, B, 64, 64, 1, 0, tables=tf1, res=X.view()), B * 4096 * 256 * 2 * 3)
# im2col / tap_expand
ir = torch.randn(B, 1, H, W, device=dev)
E = torch.zeros(n, 64, device=dev, dtype=torch.bfloat16)
timeit("im2col inc 7x7 reflect", lambda: be.im2col(ir, None, None, None, B, H, W, 7, 1, 3, 1, H, W, 0, E), n * 64 * 2 + n * 4)
g = torch.randn(B, 3, H, W, device=dev); yv = torch.tanh(torch.randn(B, 3, H, W, device=dev))
Eo = torch.zeros(y4.rows, 64, device=dev, dtype=torch.bfloat16)
db = torch.zeros(3, device=dev)
timeit("tap_expand outc (+dbias)", lambda: be.tap_expand(g, yv, [(0, s - 3) for s in range(7)], 3, B, H, W, y4.hp, y4.wp, 3, 3, Eo, dbias=db), y4.rows * 128 + n * 3 * 8 * 2)
# fp32 stand-alone stencils (north star: >= 70% of HBM peak)
x = torch.randn(B, 128, H, W, device=dev); o = torch.empty(B, 128, H // 2, W // 2, device=dev)
timeit("Downsample fp32 NCHW 128ch 256^2", lambda: be.stencil_nchw(x, o, tdn), x.numel() * 4 + o.numel() * 4)
x2 = torch.randn(B, 128, H // 2, W // 2, device=dev); o2 = torch.empty(B, 128, H, W, device=dev)
timeit("UpsampleAA fp32 NCHW 128ch ->256^2", lambda: be.stencil_nchw(x2, o2, tup), x2.numel() * 4 + o2.numel() * 4)
f = torch.tanh(torch.randn(B, 3, H, W, device=dev)); r = torch.rand(B, 3, H, W, device=dev) * 2 - 1
sums = torch.zeros(8 + B, device=dev); d = torch.empty_like(f)
timeit("pixel_loss (L1+TV value+grad)", lambda: be.pixel_loss(f, r, 1.0, 1.0, 1.0, sums[:3], d), f.numel() * 4 * 3)
from irc_b200.train_step import gaussian_window
win = gaussian_window(); ga, gb, gc = (torch.empty_like(f) for _ in range(3))
timeit("ssim_fwd (+3 maps)", lambda: be.ssim_fwd(f, r, .5, .5, win, sums[8:], ga, gb, gc), f.numel() * 4 * 5)
timeit("ssim_bwd", lambda: be.ssim_bwd(f, r, .5, .5, win, ga, gb, gc, 1.0, d, True), f.numel() * 4 * 7)

# ---- round 2: L2-resident single-launch InstanceNorm backward (image groups) against the two-pass version
print("\n# in_bwd large maps: two-pass (groups=0) vs L2-resident single launch with K image groups in flight")
Z2f, g2f, dZ2f = F(H // 2, W // 2, 1, 256), F(H // 2, W // 2, 0, 256), F(H // 2, W // 2, 1, 256)
s256b = st(256); be.in_stats(Z2f.view(), 256, B, H // 2, W // 2, s256b)
Z3f, g3f, dZ3f = F(H // 2, W // 2, 1, 128), F(H // 2, W // 2, 0, 128), F(H // 2, W // 2, 1, 128)
s128b = st(128); be.in_stats(Z3f.view(), 128, B, H // 2, W // 2, s128b)
for grp in (0, 4, 8, 16):
    be.inbwd_l2_groups = grp
    timeit(f"[K={grp}] in_bwd 2src 64ch 256^2", lambda: be.in_bwd(Z0.view(), G1.view(128), dZ0.view(), 64, B, H, W, stats=s64, cnt=H * W, act=1, g2=G2.view(), bsum=bs),
           n * 64 * 2 * 4)
    timeit(f"[K={grp}] in_bwd 1src 64ch 256^2", lambda: be.in_bwd(Z4.view(), G4.view(), dZ4.view(), 64, B, H, W, stats=s64, cnt=H * W, act=1, bsum=bs), n * 64 * 2 * 3)
    timeit(f"[K={grp}] in_bwd 1src 128ch 256^2", lambda: be.in_bwd(Z1.view(), g1.view(), dZ1.view(), 128, B, H, W, stats=s128, cnt=H * W, act=1, bsum=bs), n * 128 * 2 * 3)
    timeit(f"[K={grp}] in_bwd 1src 256ch 128^2", lambda: be.in_bwd(Z2f.view(), g2f.view(), dZ2f.view(), 256, B, H // 2, W // 2, stats=s256b, cnt=H * W // 4, act=1, bsum=bs), n // 4 * 256 * 2 * 3)
    timeit(f"[K={grp}] in_bwd 1src 128ch 128^2", lambda: be.in_bwd(Z3f.view(), g3f.view(), dZ3f.view(), 128, B, H // 2, W // 2, stats=s128b, cnt=H * W // 4, act=1, bsum=bs), n // 4 * 128 * 2 * 3)
be.inbwd_l2_groups = 4
print("\n# cluster InstanceNorm backward of the ResNet bottleneck with / without the reflection fold (algorithmic: 3 units)")
timeit("in_bwd fused 256ch 64^2, fold_pad=1", lambda: be.in_bwd(Zb.view(), Gh.view(), dZa.view(), 256, B, 64, 64, stats=s256, cnt=4096, act=1, bsum=bs, fold_pad=1), B * 4096 * 256 * 2 * 3)
timeit("in_bwd fused 256ch 64^2, no fold", lambda: be.in_bwd(Zb.view(), Gh.view(), dZa.view(), 256, B, 64, 64, stats=s256, cnt=4096, act=1, bsum=bs), B * 4096 * 256 * 2 * 3)
timeit("in_apply fused 256ch 64^2 (+res, reflect ring)", lambda: be.in_apply(Zb.view(), Y.view(), 256, B, 64, 64, 1, 1, s256, act=0, res=X.view()), B * 4096 * 256 * 2 * 3)
timeit("in_apply fused 256ch 64^2 (ReLU, no residual, reflect ring)", lambda: be.in_apply(Zb.view(), Y.view(), 256, B, 64, 64, 1, 1, s256, act=1), B * 4096 * 256 * 2 * 2)
be.fused_in_bwd = False
for grp in (4, 8, 16):
    be.inbwd_l2_groups = grp
    timeit(f"[K={grp}] in_bwd L2 kernel 256ch 64^2, no fold", lambda: be.in_bwd(Zb.view(), Gh.view(), dZa.view(), 256, B, 64, 64, stats=s256, cnt=4096, act=1, bsum=bs), B * 4096 * 256 * 2 * 3)
    timeit(f"[K={grp}] fold_inplace + in_bwd L2 kernel 256ch 64^2", lambda: be.in_bwd(Zb.view(), Gh.view(), dZa.view(), 256, B, 64, 64, stats=s256, cnt=4096, act=1, bsum=bs, fold_pad=1), B * 4096 * 256 * 2 * 3)
be.fused_in_bwd = True; be.inbwd_l2_groups = 16
# warm L2 variant: the gradient was just written by the previous kernel (as inside the step): no flush between producer and consumer
def warm(name, prod, fn, bytes_, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.sum(); prod()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:52s} {t * 1e3:8.1f} us  {bytes_ / t / 1e6:7.0f} GB/s")
warm("in_bwd fused, fold, g just written (L2-warm)", lambda: Gh.t.mul_(1.0), lambda: be.in_bwd(Zb.view(), Gh.view(), dZa.view(), 256, B, 64, 64, stats=s256, cnt=4096, act=1, bsum=bs, fold_pad=1), B * 4096 * 256 * 2 * 3)
be.fused_in_bwd = False
warm("in_bwd L2 kernel (+fold pass), g just written", lambda: Gh.t.mul_(1.0), lambda: be.in_bwd(Zb.view(), Gh.view(), dZa.view(), 256, B, 64, 64, stats=s256, cnt=4096, act=1, bsum=bs, fold_pad=1), B * 4096 * 256 * 2 * 3)
be.fused_in_bwd = True
