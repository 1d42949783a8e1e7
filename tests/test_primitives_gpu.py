"""Every memory-bound primitive of the C ABI vs its torch restatement (tests/ref_backend.py) on the
same device buffers, plus the golden vectors of the reference for the stencils / losses / metrics."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_small.npz")


@pytest.fixture(scope="module")
def bes():
    import irc_b200
    from irc_b200._native import CudaBackend
    from ref_backend import RefBackend
    return CudaBackend(), RefBackend()


def gen(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


def rnd(shape, g, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(dtype)


def close(a, b, tol, what=""):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    ref = max(b.abs().max().item(), 1e-6)
    assert err <= tol * ref, f"{what}: max err {err} vs scale {ref}"


def frame(n, h, w, p, c, g):
    from irc_b200 import layout as L
    f = L.Frame(n, h, w, p, c, "cuda")
    f.t.copy_(rnd(f.t.shape, g))
    return f


def both(bes, fn, outs):
    """run fn(backend) for both backends on clones of the output buffers; return the two output lists"""
    res = []
    for be in bes:
        bufs = [o.clone() for o in outs]
        fn(be, *bufs)
        torch.cuda.synchronize()
        res.append(bufs)
    return res


@pytest.mark.parametrize("C,H,W", [(64, 16, 12), (256, 8, 8), (192, 12, 20)])
def test_in_stats(bes, C, H, W):
    z = frame(3, H, W, 1, C, gen(1))
    a, b = both(bes, lambda be, st: be.in_stats(z.view(), C, 3, H, W, st), [torch.zeros(3, C, 2, device="cuda")])
    close(a[0], b[0], 2e-4, "in_stats")


@pytest.mark.parametrize("mode", ["apply_reflect", "apply_res", "down", "up", "s2d", "fold"])
def test_gather(bes, mode):
    from irc_b200 import layout as L
    from irc_b200._native import View
    g = gen(2)
    n, C, H, W = 2, 64, 12, 8
    src = frame(n, H, W, 1, C, g)
    st = torch.zeros(n, C, 2, device="cuda")
    bes[1].in_stats(src.view(), C, n, H, W, st)
    if mode == "apply_reflect":
        dst = frame(n, H, W, 3, 128, g)
        fn = lambda be, d: be.gather(src.view(), View(d, 64, dst.hp, dst.wp, 3, 3), C, n, H, W, 3, 1, stats=st, cnt=H * W, act=1)
    elif mode == "apply_res":
        dst = frame(n, H, W, 1, C, g); res = frame(n, H, W, 1, C, g)
        fn = lambda be, d: be.gather(src.view(), View(d, 0, dst.hp, dst.wp, 1, 1), C, n, H, W, 1, 1, stats=st, cnt=H * W, act=0, res=res.view())
    elif mode == "down":
        dst = frame(n, H // 2, W // 2, 1, C, g)
        t = L.make_tables(L.down_matrix(H), L.down_matrix(W), "cuda")
        fn = lambda be, d: be.gather(src.view(), View(d, 0, dst.hp, dst.wp, 1, 1), C, n, H // 2, W // 2, 1, 0, tables=t, stats=st, cnt=H * W, act=1)
    elif mode == "up":
        dst = frame(n, 2 * H, 2 * W, 1, C, g)
        t = L.make_tables(L.up_matrix(H), L.up_matrix(W), "cuda")
        fn = lambda be, d: be.gather(src.view(), View(d, 0, dst.hp, dst.wp, 1, 1), C, n, 2 * H, 2 * W, 1, 0, tables=t)
    elif mode == "s2d":
        dst = frame(n, (H + 2) // 2, (W + 2) // 2, 0, 4 * C, g)
        fn = lambda be, d: be.gather(src.view(), View(d, 0, dst.hp, dst.wp), C, n, H, W, 1, 0, stats=st, cnt=H * W, act=2, slope=0.2, dst_s2d=1)
    else:
        dst = frame(n, H, W, 1, C, g); src2 = frame(n, H, W, 1, C, g)
        t = L.make_tables(L.fold_matrix(H, 1), L.fold_matrix(W, 1), "cuda")
        fn = lambda be, d: be.gather(src.pview(), View(d, 0, dst.hp, dst.wp, 1, 1), C, n, H, W, 1, 0, tables=t, src2=src2.pview())
    a, b = both(bes, fn, [dst.t])
    close(a[0], b[0], 1e-2, mode)


@pytest.mark.parametrize("mode", ["down_zero", "down_reflect", "up_plain", "up_norm", "upT", "upT_pad", "downT", "downT_two"])
@pytest.mark.parametrize("C,H,W", [(128, 24, 40), (256, 16, 12), (64, 36, 20)])
def test_gather_stream(bes, mode, C, H, W):
    """the streaming separable kernel (anti-aliased Downsample / UpsampleAA and their transposes, irc:307-310, :350-355) in
    every configuration the generator plans use, over several strips and column chunks, against the torch restatement"""
    from irc_b200 import layout as L
    from irc_b200._native import View
    g = gen(5)
    n = 2
    h2, w2 = H // 2, W // 2
    mk = lambda my, mx: L.make_tables(my, mx, "cuda")
    if mode.startswith("down") and not mode.startswith("downT"):
        src = frame(n, H, W, 1, C, g); dst = frame(n, h2, w2, 1, C + 64, g)
        st = torch.zeros(n, C, 2, device="cuda"); bes[1].in_stats(src.view(), C, n, H, W, st)
        t = mk(L.down_matrix(H), L.down_matrix(W))
        halo = 1 if mode == "down_reflect" else 0
        fn = lambda be, d: be.gather(src.view(), View(d, 64, dst.hp, dst.wp, 1, 1), C, n, h2, w2, 1, halo, tables=t, stats=st, cnt=H * W, act=1)
    elif mode in ("up_plain", "up_norm"):
        src = frame(n, h2, w2, 1, C, g); dst = frame(n, H, W, 1, C, g)
        st = torch.zeros(n, C, 2, device="cuda"); bes[1].in_stats(src.view(), C, n, h2, w2, st)
        t = mk(L.up_matrix(h2), L.up_matrix(w2))
        kw = dict(stats=st, cnt=h2 * w2, act=2, slope=0.2) if mode == "up_norm" else {}
        fn = lambda be, d: be.gather(src.view(), View(d, 0, dst.hp, dst.wp, 1, 1), C, n, H, W, 1, 0, tables=t, **kw)
    elif mode in ("upT", "upT_pad"):
        p = 1 if mode == "upT_pad" else 0
        src = frame(n, H, W, 1, C + 64, g); dst = frame(n, h2, w2, p, C, g)
        t = mk(L.up_matrix(h2).T, L.up_matrix(w2).T)
        fn = lambda be, d: be.gather(src.view(64), View(d, 0, dst.hp, dst.wp, p, p), C, n, h2, w2, p, 0, tables=t)
    else:
        src = frame(n, h2, w2, 1, C, g); src2 = frame(n, h2, w2, 1, C, g); dst = frame(n, H, W, 0, C, g)
        t = mk(L.down_matrix(H).T, L.down_matrix(W).T)
        kw = dict(src2=src2.view()) if mode == "downT_two" else {}
        fn = lambda be, d: be.gather(src.view(), View(d, 0, dst.hp, dst.wp, 0, 0), C, n, H, W, 0, 0, tables=t, **kw)
    assert t.stream_window() > 0
    a, b = both(bes, fn, [dst.t])
    close(a[0], b[0], 1e-2, mode)
    # the same call through the other kernels of irc_gather gives the same frame (ring included)
    be = bes[0]
    old = be.gather_mode
    try:
        be.gather_mode = "lean"
        d2 = dst.t.clone(); fn(be, d2); torch.cuda.synchronize()
    finally:
        be.gather_mode = old
    close(a[0], d2, 1e-2, mode + " vs lean")


@pytest.mark.parametrize("kind", ["down", "up", "downT", "upT"])
@pytest.mark.parametrize("n,c,h,w", [(2, 3, 16, 24), (1, 5, 40, 72), (3, 2, 10, 300)])
def test_stencil_nchw_stream(bes, kind, n, c, h, w):
    """stand-alone fp32 Downsample / UpsampleAA (irc:307-310, :350-355) and their transposes: streaming kernel == row kernel ==
    torch restatement; odd plane counts, several strips / column chunks, accumulate mode"""
    from irc_b200 import layout as L
    g = gen(11)
    if kind == "down":
        t = L.make_tables(L.down_matrix(h), L.down_matrix(w), "cuda"); ho, wo, hi, wi = h // 2, w // 2, h, w
    elif kind == "up":
        t = L.make_tables(L.up_matrix(h), L.up_matrix(w), "cuda"); ho, wo, hi, wi = 2 * h, 2 * w, h, w
    elif kind == "downT":
        t = L.make_tables(L.down_matrix(h).T, L.down_matrix(w).T, "cuda"); ho, wo, hi, wi = h, w, h // 2, w // 2
    else:
        t = L.make_tables(L.up_matrix(h).T, L.up_matrix(w).T, "cuda"); ho, wo, hi, wi = h, w, 2 * h, 2 * w
    assert t.stream_window(8) > 0
    x = torch.randn(n, c, hi, wi, device="cuda", generator=g)
    o0 = torch.randn(n, c, ho, wo, device="cuda", generator=g)
    for acc in (False, True):
        a, b = both(bes, lambda be, o: be.stencil_nchw(x, o, t, accumulate=acc), [o0])
        close(a[0], b[0], 1e-5, f"{kind} stream acc={acc}")
        be = bes[0]
        old = be.gather_mode
        try:
            be.gather_mode = "lean"
            o2 = o0.clone(); be.stencil_nchw(x, o2, t, accumulate=acc); torch.cuda.synchronize()
        finally:
            be.gather_mode = old
        close(a[0], o2, 1e-5, f"{kind} stream vs row kernel acc={acc}")


@pytest.mark.parametrize("C,H,W,act,pad,halo,with_res", [(256, 64, 64, 1, 1, 1, False), (256, 64, 64, 0, 1, 1, True), (512, 31, 31, 2, 1, 0, False),
                                                        (128, 24, 40, 1, 3, 1, False), (64, 8, 12, 2, 0, 0, True), (32, 5, 7, 1, 1, 0, False)])
def test_in_apply_fused_cluster(bes, C, H, W, act, pad, halo, with_res):
    """single-pass InstanceNorm statistics + apply (+ residual, + ring) == irc_in_stats + irc_gather == torch restatement"""
    from irc_b200._native import View
    g = gen(14)
    n = 3
    z = frame(n, H, W, 1, C, g)
    res = frame(n, H, W, 1, C, g) if with_res else None
    dst = frame(n, H, W, pad, C + 32, g)
    outs = []
    for be, fused in ((bes[0], True), (bes[0], False), (bes[1], False)):
        d = dst.t.clone(); st = torch.zeros(n, C, 2, device="cuda")
        old = getattr(be, "fused_in_apply", None)
        if old is not None:
            be.fused_in_apply = fused
        try:
            be.in_apply(z.view(), View(d, 32, dst.hp, dst.wp, pad, pad), C, n, H, W, pad, halo, st, act=act, slope=0.2,
                        res=None if res is None else res.view())
        finally:
            if old is not None:
                be.fused_in_apply = old
        torch.cuda.synchronize()
        outs.append((d, st))
    close(outs[0][1], outs[2][1], 2e-4, "stats fused vs torch")
    close(outs[0][1], outs[1][1], 2e-4, "stats fused vs two-pass")
    close(outs[0][0], outs[2][0], 1e-2, "frame fused vs torch")
    close(outs[0][0], outs[1][0], 8e-3, "frame fused vs two-pass")
    assert torch.equal(outs[0][0][:, :32], dst.t[:, :32])          # channels outside the view untouched


@pytest.mark.parametrize("C,H,W,act,fold", [(256, 64, 64, 1, 1), (256, 64, 64, 0, 0), (512, 31, 31, 2, 0), (128, 24, 40, 1, 1),
                                             (64, 8, 12, 2, 3), (32, 5, 7, 1, 0)])
def test_in_bwd_fused_cluster(bes, C, H, W, act, fold):
    """single-pass cluster kernel (InstanceNorm + activation backward, optional ReflectionPad^T on load) == the two-pass
    kernels (+ in-place fold) == the torch restatement; cluster sizes 1..8, ragged tails, LeakyReLU / ReLU / none"""
    from irc_b200._native import View
    g = gen(13)
    n = 3
    pz = 1
    pg = max(fold, 1)
    z = frame(n, H, W, pz, C, g)
    st = torch.zeros(n, C, 2, device="cuda")
    bes[1].in_stats(z.view(), C, n, H, W, st)
    gsrc = frame(n, H, W, pg, C, g)
    if not fold:
        gsrc.t.view(n, gsrc.hp, gsrc.wp, C)[:, :pg] = 7.0       # ring garbage must be ignored without a fold
    dz = frame(n, H, W, 1, C, g)
    bs = torch.zeros(n, 512, 2, device="cuda")

    def run(be, d, fused):
        gc = gsrc.t.clone()                                       # the two-pass path folds in place
        gv = View(gc, 0, gsrc.hp, gsrc.wp, pg, pg)
        old = getattr(be, "fused_in_bwd", None)
        if old is not None:
            be.fused_in_bwd = fused
        try:
            be.in_bwd(z.view(), gv, View(d, 0, dz.hp, dz.wp, 1, 1), C, n, H, W, stats=st, cnt=H * W, act=act, slope=0.2, bsum=bs,
                      fold_pad=fold)
        finally:
            if old is not None:
                be.fused_in_bwd = old
        torch.cuda.synchronize()

    outs = []
    for be, fused in ((bes[0], True), (bes[0], False), (bes[1], False)):
        d = dz.t.clone(); run(be, d, fused); outs.append(d)
    close(outs[0], outs[2], 1e-2, "fused vs torch")
    close(outs[0], outs[1], 4e-3, "fused vs two-pass")
    assert torch.equal(outs[0].view(n, dz.hp, dz.wp, C)[:, 0], dz.t.view(n, dz.hp, dz.wp, C)[:, 0])     # ring of dz untouched


@pytest.mark.parametrize("C,H,W,two,groups,n", [(64, 80, 72, True, 3, 5), (128, 72, 80, False, 4, 5), (256, 66, 70, False, 2, 3), (64, 128, 96, False, 8, 16)])
def test_in_bwd_l2_resident(bes, C, H, W, two, groups, n):
    """single-launch, L2-resident InstanceNorm backward for large maps (image groups + counter barrier) == the two-pass
    kernels bit for bit in the reductions' inputs (same formula; summation order differs) == the torch restatement;
    uneven images per group, launched twice (the barrier counters must re-arm), one and two gradient sources"""
    from irc_b200._native import View
    g = gen(17)
    z = frame(n, H, W, 1, C, g)
    st = torch.zeros(n, C, 2, device="cuda")
    bes[1].in_stats(z.view(), C, n, H, W, st)
    g1 = frame(n, H, W, 1, 2 * C, g); g2 = frame(n, H, W, 0, C, g)
    dz = frame(n, H, W, 1, C, g)
    outs, sums = [], []
    for be, grp in ((bes[0], groups), (bes[0], groups), (bes[0], 0), (bes[1], 0)):
        d = dz.t.clone(); bs = torch.zeros(n, C, 2, device="cuda")
        old = getattr(be, "inbwd_l2_groups", None)
        if old is not None:
            be.inbwd_l2_groups = grp
        try:
            be.in_bwd(z.view(), g1.view(C), View(d, 0, dz.hp, dz.wp, 1, 1), C, n, H, W, stats=st, cnt=H * W, act=1, bsum=bs,
                      g2=g2.view() if two else None)
        finally:
            if old is not None:
                be.inbwd_l2_groups = old
        torch.cuda.synchronize()
        outs.append(d); sums.append(bs)
    assert torch.equal(outs[0], outs[1]) and torch.equal(sums[0], sums[1])          # deterministic, counters re-armed
    close(sums[0], sums[2], 1e-4, "bsum l2 vs two-pass"); close(sums[0], sums[3], 1e-3, "bsum l2 vs torch")
    close(outs[0], outs[2], 4e-3, "l2 vs two-pass"); close(outs[0], outs[3], 1e-2, "l2 vs torch")
    assert torch.equal(outs[0].view(n, dz.hp, dz.wp, C)[:, 0], dz.t.view(n, dz.hp, dz.wp, C)[:, 0])     # ring of dz untouched


@pytest.mark.parametrize("mode", ["norm_relu_fold", "norm_none_two", "plain_lrelu", "upT", "s2d_src"])
def test_in_bwd(bes, mode):
    from irc_b200 import layout as L
    from irc_b200._native import View
    g = gen(3)
    n, C, H, W = 2, 128, 8, 12
    z = frame(n, H, W, 1, C, g)
    st = torch.zeros(n, C, 2, device="cuda")
    bes[1].in_stats(z.view(), C, n, H, W, st)
    dz = frame(n, H, W, 1, C, g)
    bs = torch.zeros(n, 512, 2, device="cuda")
    kw = dict(stats=st, cnt=H * W, bsum=bs)
    if mode == "norm_relu_fold":
        gsrc = frame(n, H, W, 1, C, g)
        t = L.make_tables(L.fold_matrix(H, 1), L.fold_matrix(W, 1), "cuda")
        fn = lambda be, d: be.in_bwd(z.view(), gsrc.pview(), View(d, 0, dz.hp, dz.wp, 1, 1), C, n, H, W, act=1, tables=t, **kw)
    elif mode == "norm_none_two":
        g1 = frame(n, H, W, 1, 256, g); g2 = frame(n, H, W, 1, C, g)
        fn = lambda be, d: be.in_bwd(z.view(), g1.view(128), View(d, 0, dz.hp, dz.wp, 1, 1), C, n, H, W, act=0, g2=g2.view(), **kw)
    elif mode == "plain_lrelu":
        g1 = frame(n, H, W, 1, C, g)
        fn = lambda be, d: be.in_bwd(z.view(), g1.view(), View(d, 0, dz.hp, dz.wp, 1, 1), C, n, H, W, act=2, slope=0.2)
    elif mode == "upT":
        g1 = frame(n, 2 * H, 2 * W, 1, C, g)
        t = L.make_tables(L.up_matrix(H).T, L.up_matrix(W).T, "cuda")
        fn = lambda be, d: be.in_bwd(z.view(), g1.view(), View(d, 0, dz.hp, dz.wp, 1, 1), C, n, H, W, act=1, tables=t, **kw)
    else:
        g1 = frame(n, (H + 2) // 2, (W + 2) // 2, 0, 4 * C, g)
        fn = lambda be, d: be.in_bwd(z.view(), View(g1.t, 0, g1.hp, g1.wp, 1, 1, C), View(d, 0, dz.hp, dz.wp, 1, 1), C, n, H, W, act=2, slope=0.2, **kw)
    a, b = both(bes, fn, [dz.t])
    close(a[0], b[0], 1.5e-2, mode)


@pytest.mark.parametrize("p,C,H,W", [(1, 64, 8, 10), (3, 64, 12, 16), (1, 256, 64, 64)])
def test_fold_inplace(bes, p, C, H, W):
    g = gen(11)
    n = 2
    fr = frame(n, H, W, p, C + 64, g)
    a, b = both(bes, lambda be, t: be.fold_inplace(t, 64, C, n, H, W, p), [fr.t])
    close(a[0], b[0], 1e-2, "fold_inplace")
    ring = a[0].view(n, H + 2 * p, W + 2 * p, -1)[..., 64:].clone()
    ring[:, p:p + H, p:p + W] = 0
    assert (ring == 0).all()
    assert torch.equal(a[0][:, :64], fr.t[:, :64])          # other channels untouched


def test_maxpool(bes):
    g = gen(4)
    n, C, H, W = 3, 64, 8, 12
    src = frame(n, H, W, 1, C, g); src.t.copy_(torch.relu(src.t))
    from irc_b200._native import View
    dst = frame(n, H // 2, W // 2, 1, C, g)
    a, b = both(bes, lambda be, d: be.maxpool2(src.view(), View(d, 0, dst.hp, dst.wp, 1, 1), C, n, H // 2, W // 2), [dst.t])
    assert torch.equal(a[0], b[0])
    gg = frame(2, H // 2, W // 2, 1, C, g); dsrc = frame(2, H, W, 1, C, g)
    sv = View(src.t[:src.rows_of(2)], 0, src.hp, src.wp, 1, 1)
    a, b = both(bes, lambda be, d: be.maxpool2_bwd(sv, gg.view(), View(d, 0, dsrc.hp, dsrc.wp, 1, 1), C, 2, H // 2, W // 2), [dsrc.t])
    assert torch.equal(a[0], b[0])


def test_colsum(bes):
    g = gen(5)
    a_ = rnd((5000, 64), g)
    ri = (torch.arange(5000, device="cuda") % 3 - 1).short()
    for r in (None, ri):
        a, b = both(bes, lambda be, o: be.colsum(a_, 0, 64, o, row_img=r), [torch.zeros(64, device="cuda")])
        close(a[0], b[0], 1e-4, "colsum")


@pytest.mark.parametrize("cfg", [dict(c1=1, c2=0, k=7, s=1, p=3, pm=1, rm=0), dict(c1=3, c2=0, k=3, s=1, p=1, pm=0, rm=1, aff=True),
                                 dict(c1=1, c2=3, k=4, s=2, p=1, pm=0, rm=2)])
def test_im2col_col2im(bes, cfg):
    g = gen(6)
    n, H, W = 2, 16, 24
    s1 = torch.randn(n, cfg["c1"], H, W, device="cuda", generator=g)
    s2 = torch.randn(n, cfg["c2"], H, W, device="cuda", generator=g) if cfg["c2"] else None
    C = cfg["c1"] + cfg["c2"]
    Ho, Wo = H // cfg["s"], W // cfg["s"]
    sc = torch.rand(C, device="cuda", generator=g) + 0.5 if cfg.get("aff") else None
    sh = torch.randn(C, device="cuda", generator=g) if cfg.get("aff") else None
    rows = bes[0].im2col_rows(cfg["rm"], n, Ho, Wo)
    assert rows == bes[1].im2col_rows(cfg["rm"], n, Ho, Wo)
    fn = lambda be, d, ri: be.im2col(s1, s2, sc, sh, n, H, W, cfg["k"], cfg["s"], cfg["p"], cfg["pm"], Ho, Wo, cfg["rm"], d, row_img=ri)
    a, b = both(bes, fn, [torch.full((rows, 64), 7.0, device="cuda", dtype=torch.bfloat16), torch.zeros(rows, device="cuda", dtype=torch.int16)])
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    if cfg["pm"] == 0:
        de = rnd((rows, 64), g)
        c_first, c_out = (1, 3) if C == 4 else (0, C)
        out0 = torch.randn(n, c_out, H, W, device="cuda", generator=g)
        fn = lambda be, o: be.col2im(de, C, c_first, c_out, n, H, W, cfg["k"], cfg["s"], cfg["p"], Ho, Wo, cfg["rm"], sc, o, True)
        a, b = both(bes, fn, [out0])
        close(a[0], b[0], 1e-5, "col2im")


@pytest.mark.parametrize("cfg", [dict(c1=1, c2=0, k=7, s=1, p=3, pm=1, rm=0, act=0), dict(c1=3, c2=0, k=3, s=1, p=1, pm=0, rm=1, aff=True, act=1, bias=True),
                                 dict(c1=1, c2=3, k=4, s=2, p=1, pm=0, rm=2, act=2, bias=True)])
@pytest.mark.parametrize("keep", [False, True])
@pytest.mark.parametrize("size", [(2, 16, 24), (3, 36, 20), (1, 8, 644)])      # the last one: wide rows, fewer lines per block
def test_smallk_conv_fwd(bes, cfg, keep, size):
    """direct small-K convolution (inc irc:458-463, VGG conv1_1 irc:664, D model.0 irc:600) == im2col operand x weights in fp32 of
    the same bf16 values, to one bf16 rounding; the operand by-product is bit-identical to irc_im2col's"""
    g = gen(16)
    n, H, W = size
    s1 = torch.randn(n, cfg["c1"], H, W, device="cuda", generator=g)
    s2 = torch.randn(n, cfg["c2"], H, W, device="cuda", generator=g) if cfg["c2"] else None
    C = cfg["c1"] + cfg["c2"]
    Ho, Wo = H // cfg["s"], W // cfg["s"]
    sc = torch.rand(C, device="cuda", generator=g) + 0.5 if cfg.get("aff") else None
    sh = torch.randn(C, device="cuda", generator=g) if cfg.get("aff") else None
    rows = bes[0].im2col_rows(cfg["rm"], n, Ho, Wo)
    w = (torch.randn(64, 64, device="cuda", generator=g) * 0.2).bfloat16()
    w[:, cfg["k"] * cfg["k"] * C:] = 0
    bias = torch.randn(64, device="cuda", generator=g) if cfg.get("bias") else None

    def fn(be, out, E, ri):
        be.smallk_conv_fwd(s1, s2, sc, sh, n, H, W, cfg["k"], cfg["s"], cfg["p"], cfg["pm"], Ho, Wo, cfg["rm"], w, out, bias=bias, act=cfg["act"],
                           slope=0.2, E=E if keep else None, row_img=ri)
    a, b = both(bes, fn, [torch.full((rows, 64), 7.0, device="cuda", dtype=torch.bfloat16), torch.full((rows, 64), 5.0, device="cuda", dtype=torch.bfloat16),
                          torch.zeros(rows, device="cuda", dtype=torch.int16)])
    assert torch.equal(a[2], b[2])
    if keep and not cfg.get("aff"):
        assert torch.equal(a[1], b[1])
    elif keep:
        # the per-channel affine is one fused multiply-add in the kernel and a multiply + add in the restatement: the fp32 values can
        # differ in the last bit, which now and then lands on the other side of a bf16 rounding boundary
        d = (a[1].float() - b[1].float()).abs()
        assert (d <= 2 ** -7 * b[1].float().abs() + 1e-6).all() and (d > 0).float().mean().item() < 1e-2
    # both sides round an fp32 value to bf16: agreement to one bf16 ulp of the largest magnitude
    err = (a[0].float() - b[0].float()).abs().max().item()
    assert err <= 2 ** -7 * b[0].float().abs().max().item() + 1e-6, err
    assert ((a[0].float() - b[0].float()).norm() / b[0].float().norm()).item() < 3e-3
    assert (a[0][a[2] < 0] == 0).all()


def test_conv_gemm_tap_mode_scale_accumulate(bes):
    """data gradient of a 3 x 3 convolution with 3 input planes (VGG conv1_1) through the horizontal-tap epilogue: += scale * (...)"""
    g = gen(17)
    n, H, W = 2, 20, 28
    hp, wp = H + 2, W + 2
    dz = torch.zeros(n, hp, wp, 64, device="cuda")
    dz[:, 1:-1, 1:-1] = torch.randn(n, H, W, 64, device="cuda", generator=g)
    dz = dz.view(-1, 64).bfloat16()
    wt = (torch.randn(32, 3 * 64, device="cuda", generator=g) * 0.1).bfloat16()
    wt[9:] = 0
    scale = torch.rand(3, device="cuda", generator=g) + 0.5
    out0 = torch.randn(n, 3, H, W, device="cuda", generator=g)
    dummy = torch.zeros(8, 32, device="cuda")
    fn = lambda be, o: be.conv_gemm(dz, 0, 64, [wp, 0, -wp], wt, 32, dummy,
                                    tap=dict(out=o, nshift=3, nco=3, H=H, W=W, hp=hp, wp=wp, oy=1, ox=1, act=0, scale=scale, accumulate=True))
    a, b = both(bes, fn, [out0])
    close(a[0], b[0], 2e-5, "tap-mode dgrad")


@pytest.mark.parametrize("size", [(2, 19, 23), (3, 64, 96), (1, 7, 40)])
def test_ssim_metric(bes, size):
    """skimage-style SSIM metric (irc:1208-1215) on the device: float64 sums, == the float64 restatement to 1e-10 per image"""
    g = gen(23)
    n, H, W = size
    gt = torch.rand(n, 3, H, W, device="cuda", generator=g)
    u8 = (torch.rand(n, H, W, 3, device="cuda", generator=g) * 255).to(torch.uint8)
    u8[0] = (gt[0].permute(1, 2, 0) * 255 + 3 * torch.randn(H, W, 3, device="cuda", generator=g)).clamp(0, 255).to(torch.uint8)
    a, b = both(bes, lambda be, s_: be.ssim_metric(u8, gt, s_), [torch.zeros(n, device="cuda", dtype=torch.float64)])
    cnt = 3 * (H - 6) * (W - 6)
    # torch's CUDA division by a scalar multiplies by the rounded reciprocal, so the restatement's u8 / 255 differs from the
    # kernel's (and numpy's, i.e. the reference's) IEEE division in the last float32 bit of some pixels: 1e-7 here, tight below
    assert ((a[0] - b[0]).abs() / cnt).max().item() < 1e-7, (a[0] / cnt, b[0] / cnt)
    import irc_oracle as O
    for i in range(n):
        want = O.skimage_ssim(gt[i].permute(1, 2, 0).cpu().numpy(), u8[i].cpu().numpy().astype("float32") / 255.0)
        assert abs(a[0][i].item() / cnt - want) < 1e-10, (i, a[0][i].item() / cnt, want)


@pytest.mark.parametrize("training", [True, False])
def test_batch_norm_on_the_instance_norm_kernels(bes, training):
    """nn.BatchNorm2d (irc:158-159): irc_bn_finalize (batch / running statistics + affine -> effective moments, running-statistics
    update), apply through irc_gather with eps < 0, backward = reduce + irc_bn_bwd_fix + apply; against torch's own batch_norm"""
    g = gen(31)
    n, C, H, W, grp = 4, 64, 12, 20, 2
    z = frame(n, H, W, 1, C, g)
    gam = torch.randn(C, device="cuda", generator=g) * 0.3 + 1.0; bet = torch.randn(C, device="cuda", generator=g) * 0.2
    rm0 = torch.randn(C, device="cuda", generator=g) * 0.1; rv0 = torch.rand(C, device="cuda", generator=g) + 0.5
    gup = frame(n, H, W, 0, C, g)

    def fn(be, st, eff, rm, rv, dst, dz, bsum, dg, db):
        be.in_stats(z.view(), C, n, H, W, st)
        be.bn_finalize(st, n, grp, C, float(H * W), gam, bet, rm, rv, eff, training=training, updates=2)
        be.gather(z.view(), View(dst, 0, H + 2, W + 2, 1, 1), C, n, H, W, 1, 0, stats=eff, cnt=grp * H * W, eps=-1.0, act=2, slope=0.2)
        be.in_bwd(z.view(), gup.view(), View(dz, 0, H + 2, W + 2, 1, 1), C, n, H, W, stats=eff, cnt=grp * H * W, eps=-1.0, act=2, slope=0.2, bsum=bsum,
                  bn=dict(group=grp, gamma=gam, beta=bet, dgamma=dg, dbeta=db))
    from irc_b200._native import View
    rows = n * (H + 2) * (W + 2)
    outs = [torch.zeros(n, C, 2, device="cuda"), torch.zeros(n, C, 2, device="cuda"), rm0.clone(), rv0.clone(),
            torch.zeros(rows, C, device="cuda", dtype=torch.bfloat16), torch.zeros(rows, C, device="cuda", dtype=torch.bfloat16),
            torch.zeros(n, C, 2, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")]
    a, b = both(bes, fn, outs)
    for i, (x, y, tol, what) in enumerate(zip(a, b, (1e-5, 1e-4, 1e-5, 1e-5, 1e-2, 2e-2, 1e-3, 2e-3, 2e-3),
                                            ("sums", "eff", "running_mean", "running_var", "y", "dx", "bsum", "dgamma", "dbeta"))):
        close(x, y, tol, what)
    # and against torch.nn.functional.batch_norm itself (group by group), in fp32 on the bf16-rounded input
    x = z.t.view(n, H + 2, W + 2, C)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().clone().requires_grad_(True)
    rm, rv = rm0.clone(), rv0.clone()
    ys = []
    for g0 in range(0, n, grp):
        for _ in range(2 if training else 1):
            yy = torch.nn.functional.batch_norm(x[g0:g0 + grp], rm, rv, gam, bet, training, 0.1, 1e-5)
        ys.append(torch.nn.functional.leaky_relu(yy, 0.2))
    y = torch.cat(ys)
    got = a[4].view(n, H + 2, W + 2, C)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float()
    assert ((got - y).norm() / y.norm()).item() < 5e-3
    if training:
        close(a[2], rm, 1e-5, "running_mean vs torch"); close(a[3], rv, 1e-5, "running_var vs torch")


def test_structured_weight_packing_equals_the_index_map():
    """irc_pack_std (shared-memory transposes of the standard layers, no index map) fills the forward and data-gradient operands with
    exactly the bits irc_pack_bf16 produces through the host-built index maps; irregular layouts keep the map"""
    import irc_b200  # noqa: F401
    from irc_b200 import layout as L
    from irc_b200._native import CudaBackend
    be = CudaBackend()
    shapes = {"a.weight": (128, 64, 3, 3), "b.weight": (256, 128, 3, 3), "c.weight": (512, 256, 4, 4), "d.weight": (64, 192, 3, 3),
              "e.weight": (64, 1, 7, 7), "f.weight": (3, 64, 7, 7), "g.weight": (128, 64, 4, 4)}
    arena = L.ParamArena(shapes, "cuda")
    arena.flat.normal_(generator=gen(41))
    P = L.Packer(arena)
    lays = [L.layout_std(P, arena, "a.weight", 128, 64, 3, 3), L.layout_std(P, arena, "b.weight", 256, 128, 3, 3),
            L.layout_im2col(P, arena, "e.weight", 64, 1, 7), L.layout_std(P, arena, "c.weight", 512, 256, 4, 4),
            L.layout_outc(P, arena, "f.weight", 3, 64, 7), L.layout_std(P, arena, "d.weight", 64, 192, 3, 3), L.layout_s2d(P, arena, "g.weight", 128, 64)]
    P.finish()
    assert P.n_jobs == 4 and P.n_mapped > 0
    P.refresh(be)
    torch.cuda.synchronize()
    got = P.packed.clone()
    for lay in lays:
        for op in (lay.w_f, lay.w_d):
            idx = torch.from_numpy(op.index).cuda()
            want = torch.where(idx >= 0, arena.flat[idx.clamp_min(0)], torch.zeros((), device="cuda")).bfloat16()
            assert torch.equal(op.t, want), (op.rows, op.cols, op.std)
    # everything through the map gives the same buffer
    P.n_jobs = 0
    P.packed.zero_()
    P.refresh(be)
    torch.cuda.synchronize()
    assert torch.equal(P.packed, got)


def test_taps(bes):
    g = gen(7)
    n, H, W, p = 2, 10, 12, 3
    hp, wp = H + 2 * p, W + 2 * p
    shifts = [(0, s - 3) for s in range(7)]
    P = torch.randn(n * hp * wp, 32, device="cuda", generator=g)
    bias = torch.randn(3, device="cuda", generator=g)
    a, b = both(bes, lambda be, o: be.tap_reduce(P, shifts, 3, n, H, W, hp, wp, p, p, bias, 3, o), [torch.zeros(n, 3, H, W, device="cuda")])
    close(a[0], b[0], 1e-5, "tap_reduce")
    gr = torch.randn(n, 3, H, W, device="cuda", generator=g); y = torch.tanh(torch.randn(n, 3, H, W, device="cuda", generator=g))
    fn = lambda be, E, db: be.tap_expand(gr, y, shifts, 3, n, H, W, hp, wp, p, p, E, dbias=db)
    a, b = both(bes, fn, [torch.full((n * hp * wp, 64), 3.0, device="cuda", dtype=torch.bfloat16), torch.zeros(3, device="cuda")])
    assert torch.equal(a[0], b[0]); close(a[1], b[1], 1e-5, "dbias")
    # live_cols_only: the 16-byte groups without tap columns (21 columns -> groups 3..7) are left as the caller filled them
    E2 = torch.full((n * hp * wp, 64), 3.0, device="cuda", dtype=torch.bfloat16); db2 = torch.zeros(3, device="cuda")
    bes[0].tap_expand(gr, y, shifts, 3, n, H, W, hp, wp, p, p, E2, dbias=db2, live_cols_only=True)
    assert torch.equal(E2[:, :24], a[0][:, :24]) and (E2[:, 32:] == 3.0).all() and torch.equal(db2, a[1])      # (columns 24..31 may be zero-filled: whole sectors)
    # 4x4 taps on a small top-left anchored frame (D model.11 at the test size: horizontal shifts reach wp/2)
    n, H, W, hp, wp = 3, 2, 2, 5, 5
    sh = [(r, s) for r in range(4) for s in range(4)]
    gr = torch.randn(n, 1, H, W, device="cuda", generator=g)
    fn = lambda be, E: be.tap_expand(gr, None, sh, 1, n, H, W, hp, wp, 0, 0, E)
    a, b = both(bes, fn, [torch.full((n * hp * wp, 64), 3.0, device="cuda", dtype=torch.bfloat16)])
    assert torch.equal(a[0], b[0])
    P = torch.randn(n * hp * wp, 32, device="cuda", generator=g)
    a, b = both(bes, lambda be, o: be.tap_reduce(P, sh, 1, n, H, W, hp, wp, 0, 0, None, 0, o), [torch.zeros(n, 1, H, W, device="cuda")])
    close(a[0], b[0], 1e-5, "tap_reduce 4x4")


def test_losses_vs_reference_golden(bes):
    """L1/TV/SSIM values and gradients against the reference's own autograd (golden fixture)."""
    from irc_b200.train_step import gaussian_window
    be = bes[0]
    gold = np.load(GOLD)
    a = torch.from_numpy(gold["loss_a"]).cuda(); b = torch.from_numpy(gold["loss_b"]).cuda()
    n, c, h, w = a.shape
    sums = torch.zeros(3, device="cuda"); d = torch.zeros_like(a)
    be.pixel_loss(a, None, 0.0, 1.0 / (n * c * (h - 1) * w), 1.0 / (n * c * h * (w - 1)), sums, d)
    tv = sums[1].item() / (n * c * (h - 1) * w) + sums[2].item() / (n * c * h * (w - 1))
    assert abs(tv - float(gold["tv"])) < 1e-6
    close(d, torch.from_numpy(gold["tv_grad"]).cuda(), 1e-5, "tv grad")
    win = gaussian_window()
    ss = torch.zeros(n, device="cuda"); ga, gb, gc = (torch.zeros_like(a) for _ in range(3))
    be.ssim_fwd(a, b, 1.0, 0.0, win, ss, ga, gb, gc)
    per = 1.0 - ss / (c * h * w)
    close(per, torch.from_numpy(gold["ssim_per_sample"]).cuda(), 2e-5, "ssim per sample")
    assert abs(1.0 - ss.sum().item() / a.numel() - float(gold["ssim"])) < 2e-5
    d.zero_()
    be.ssim_bwd(a, b, 1.0, 0.0, win, ga, gb, gc, -1.0 / a.numel(), d, False)
    close(d, torch.from_numpy(gold["ssim_grad"]).cuda(), 2e-3, "ssim grad")
    # L1 value + sign gradient
    sums.zero_()
    be.pixel_loss(a, b, 1.0, 0.0, 0.0, sums, d)
    assert abs(sums[0].item() - (a - b).abs().sum().item()) < 1e-2
    assert torch.equal(d, torch.sign(a - b))


@pytest.mark.parametrize("n,h,w", [(2, 45, 70), (1, 32, 64), (2, 256, 256), (1, 7, 5), (64, 128, 160)])
def test_ssim_loss_paths(bes, n, h, w):
    """ssim_loss_torch (irc:714-750) value and gradient against the oracle's autograd at sizes that do not divide the 64 x 32
    32 x 16 tiles, at the train-step plane size, below the window size, and on the streaming kernels (large batch)"""
    import irc_oracle as O
    from irc_b200.train_step import gaussian_window
    be = bes[0]
    g = gen(31)
    a = torch.tanh(torch.randn(n, 3, h, w, device="cuda", generator=g)); b = torch.rand(n, 3, h, w, device="cuda", generator=g) * 2 - 1
    x = a.cpu().clone().requires_grad_(True)
    loss = O.ssim_loss((x + 1) / 2, (b.cpu() + 1) / 2)
    loss.backward()
    win = gaussian_window()
    ss = torch.zeros(n, device="cuda"); ga, gb, gc = (torch.zeros_like(a) for _ in range(3))
    be.ssim_fwd(a, b, 0.5, 0.5, win, ss, ga, gb, gc)
    assert abs(1.0 - ss.sum().item() / a.numel() - loss.item()) < 2e-5
    d = torch.zeros_like(a)
    be.ssim_bwd(a, b, 0.5, 0.5, win, ga, gb, gc, -1.0 / a.numel(), d, False)
    close(d, x.grad.cuda(), 2e-3, "ssim grad")
    d2 = d.clone()
    be.ssim_bwd(a, b, 0.5, 0.5, win, ga, gb, gc, -1.0 / a.numel(), d2, True)
    close(d2, 2.0 * x.grad.cuda(), 2e-3, "ssim grad (accumulate)")


@pytest.mark.parametrize("h,w", [(20, 32), (17, 30), (64, 260), (70, 640), (5, 132), (1, 8)])
def test_pixel_loss_paths(bes, h, w):
    """fused L1 + TV (irc:686-694, :1664): the streaming kernel (W % 4 == 0: several column tiles, strips that do not divide H,
    a single row) and the scalar one against torch autograd; TV alone (no target) through the same kernel"""
    be = bes[0]
    g = gen(9)
    a = torch.randn(3, 3, h, w, device="cuda", generator=g); b = torch.randn(3, 3, h, w, device="cuda", generator=g)
    x = a.clone().requires_grad_(True)
    l1 = (x - b).abs().sum(); tvv = (x[:, :, 1:] - x[:, :, :-1]).abs().sum(); tvh = (x[:, :, :, 1:] - x[:, :, :, :-1]).abs().sum()
    (0.7 * l1 + 0.3 * tvv + 0.2 * tvh).backward()
    sums = torch.zeros(3, device="cuda"); d = torch.zeros_like(a)
    be.pixel_loss(a, b, 0.7, 0.3, 0.2, sums, d)
    for got, ref in zip(sums.tolist(), (l1.item(), tvv.item(), tvh.item())):
        assert abs(got - ref) <= 1e-4 * abs(ref)
    close(d, x.grad, 1e-6, "pixel_loss grad")
    x2 = a.clone().requires_grad_(True)
    (0.3 * (x2[:, :, 1:] - x2[:, :, :-1]).abs().sum() + 0.2 * (x2[:, :, :, 1:] - x2[:, :, :, :-1]).abs().sum()).backward()
    sums2 = torch.zeros(3, device="cuda"); d2 = torch.zeros_like(a)
    be.pixel_loss(a, None, 0.0, 0.3, 0.2, sums2, d2)
    assert sums2[0].item() == 0 and abs(sums2[1].item() - tvv.item()) <= 1e-4 * max(abs(tvv.item()), 1e-6)
    close(d2, x2.grad, 1e-6, "tv grad") if h > 1 or w > 1 else None


def test_hinge_featl1(bes):
    g = gen(8)
    pred = torch.randn(4 * 900, device="cuda", generator=g) * 2
    for mode in (0, 1):
        a, b = both(bes, lambda be, s, d: be.hinge(pred, 1800, mode, 0.25, 0.5, s, d), [torch.zeros(3, device="cuda"), torch.zeros_like(pred)])
        close(a[0], b[0], 1e-5, "hinge sums"); close(a[1], b[1], 1e-6, "hinge grad")
    feat = torch.relu(rnd((2 * 300, 256), g))
    a, b = both(bes, lambda be, s, d: be.feat_l1(feat, 300, 256, 0.125, s, d), [torch.zeros(1, device="cuda"), torch.zeros(300, 256, device="cuda", dtype=torch.bfloat16)])
    close(a[0], b[0], 1e-4, "feat sum"); assert torch.equal(a[1], b[1])


def test_quantize_metrics_bit_exact(bes):
    """irc:865-876 truncation and irc:1197-1205 metrics against the reference's own output."""
    be = bes[0]
    gold = np.load(GOLD)
    fake = torch.from_numpy(gold["fake"]).cuda()
    gt = torch.from_numpy(gold["metrics_gt"]).permute(2, 0, 1).unsqueeze(0).contiguous().cuda()
    u8 = torch.zeros(1, 32, 32, 3, device="cuda", dtype=torch.uint8); sums = torch.zeros(1, 2, device="cuda", dtype=torch.float64)
    be.quantize_metrics(fake[:1].contiguous(), gt, u8, sums)
    assert np.array_equal(u8[0].cpu().numpy(), gold["quant_u8"])
    mae, mse = (sums[0] / (3 * 32 * 32)).tolist()
    psnr = -10.0 * np.log10(mse + 1e-12)
    for got, ref in zip((mae, mse, psnr), gold["metrics"]):
        assert abs(got - ref) < 5e-7 * max(1, abs(ref)), (got, ref)
    # batched, odd pixel count (scalar kernel) and 4-pixel vector kernel against the oracle's numpy formulas
    import irc_oracle as O
    for (n, hh, ww) in ((3, 12, 20), (2, 7, 9)):
        f = torch.tanh(torch.randn(n, 3, hh, ww, device="cuda", generator=gen(4))) * 1.2
        t = torch.rand(n, 3, hh, ww, device="cuda", generator=gen(5))
        u = torch.zeros(n, hh, ww, 3, device="cuda", dtype=torch.uint8); sm = torch.zeros(n, 2, device="cuda", dtype=torch.float64)
        be.quantize_metrics(f, t, u, sm)
        for i in range(n):
            q = O.quantize_u8(f[i].cpu())
            assert np.array_equal(u[i].cpu().numpy(), q)
            m = O.compute_metrics(q.astype(np.float32) / 255.0, t[i].permute(1, 2, 0).cpu().numpy())
            assert abs(sm[i, 0].item() / (3 * hh * ww) - m[0]) < 1e-6 and abs(sm[i, 1].item() / (3 * hh * ww) - m[1]) < 1e-6
    probe = torch.full((1, 3, 4, 4), 254.87 / 255.0 * 2 - 1, device="cuda")
    u = torch.zeros(1, 4, 4, 3, device="cuda", dtype=torch.uint8)
    be.quantize_metrics(probe, None, u, None)
    assert (u == 254).all()


def test_adam_pack_gather(bes):
    g = gen(9)
    n = 10007
    p0 = torch.randn(n, device="cuda", generator=g); gr = torch.randn(n, device="cuda", generator=g) * 1e-3
    m0 = torch.randn(n, device="cuda", generator=g) * 1e-3; v0 = torch.rand(n, device="cuda", generator=g) * 1e-6
    hyper = torch.tensor([2e-4, 0.5, 0.999, 1e-8, 0.7, 0.5, 0.0, 0.0], device="cuda", dtype=torch.float64)
    a, b = both(bes, lambda be, p, m, v, t: be.adam(p, gr, m, v, hyper, t), [p0, m0, v0, torch.full((1,), 2, device="cuda", dtype=torch.int64)])
    for x, y in zip(a[:3], b[:3]):
        close(x, y, 2e-6, "adam")
    assert int(a[3].item()) == 3 and int(b[3].item()) == 3          # the device step counter advances inside the call
    # torch.optim.Adam itself
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=2e-4, betas=(0.5, 0.999))
    for _ in range(3):
        pt.grad = gr.clone(); opt.step()
    pc, mc, vc = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    h = torch.tensor([2e-4, 0.5, 0.999, 1e-8, 1.0, 1.0, 0.0, 0.0], device="cuda", dtype=torch.float64)
    tdev = torch.zeros(1, device="cuda", dtype=torch.int64)
    for _ in range(3):
        bes[0].adam(pc, gr, mc, vc, h, tdev)          # back to back, no host involvement between the steps
    assert int(tdev.item()) == 3
    close(pc, pt.detach(), 1e-6, "adam vs torch.optim.Adam")
    src = torch.randn(5000, device="cuda", generator=g)
    mp = torch.randint(-1, 5000, (7777,), device="cuda", generator=g, dtype=torch.int32)
    a, b = both(bes, lambda be, d: be.pack_bf16(src, mp, d), [torch.zeros(7777, device="cuda", dtype=torch.bfloat16)])
    assert torch.equal(a[0], b[0])
    part = torch.randn(3 * 6000, device="cuda", generator=g)
    mp2 = torch.randint(-1, 6000, (4000,), device="cuda", generator=g, dtype=torch.int32)
    a, b = both(bes, lambda be, d: be.gather_sum(part, mp2, 3, 6000, d), [torch.zeros(4000, device="cuda")])
    close(a[0], b[0], 1e-6, "gather_sum")
