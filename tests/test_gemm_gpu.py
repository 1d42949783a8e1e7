"""tcgen05 GEMM kernels through the C ABI vs a plain torch fp32 evaluation of the same sums."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _conv_ref(A, taps, W, cin, a_off):
    rows = A.shape[0]
    Af = A.float()
    out = torch.zeros(rows, W.shape[0], device=A.device)
    q = torch.arange(rows, device=A.device)
    for t, sh in enumerate(taps):
        idx = q + sh
        ok = (idx >= 0) & (idx < rows)
        g = Af[idx.clamp(0, rows - 1), a_off:a_off + cin] * ok[:, None]
        out += g @ W[:, t * cin:(t + 1) * cin].float().t()
    return out


@pytest.mark.parametrize("rows,ld,a_off,cin,taps,n_out,fp32", [
    (1000, 64, 0, 64, [0], 64, False),
    (128 * 5 + 17, 192, 64, 128, [-11, -10, -9, -1, 0, 1, 9, 10, 11], 256, False),
    (4096, 256, 0, 256, [-67, -66, -65, -1, 0, 1, 65, 66, 67], 256, False),
    (3000, 128, 0, 128, [0, 1, 50, 51], 128, False),
    (2000, 64, 0, 64, [-3, 0, 3], 32, True),
    (150 * 128 + 5, 64, 0, 64, [-1, 0, 1], 192, False),
    (2500, 512, 0, 512, [0, 1, 2, 3], 512, False),
    (148 * 256 * 2 + 77, 64, 0, 64, [-1, 0, 1], 128, False),
    (148 * 126 * 2 + 50, 64, 0, 64, [-259, -258, -257, -1, 0, 1, 257, 258, 259], 64, False),       # packed-taps shapes (64 outputs, 3 x 3)
    (5000, 256, 64, 192, [67, 66, 65, 1, 0, -1, -65, -66, -67], 64, False),                        # ... data-gradient tap order, channel offset
    (126 * 3, 64, 0, 64, [-1, 0, 1], 64, False),                                                   # ... rows == a whole number of 126-row tiles
])
@pytest.mark.parametrize("mt", [1, 2, 4])
@pytest.mark.parametrize("reuse", [0, 1, 2])
def test_conv_gemm(rows, ld, a_off, cin, taps, n_out, fp32, mt, reuse):
    """reuse = 1: the tap-run kernel (one staged A box per kernel row of taps, shifted descriptors) wherever taps form runs;
    reuse = 2: packed taps (the three taps of a kernel row stacked along N, shifted sum in the epilogue) where eligible"""
    if reuse == 2 and mt != 1:
        pytest.skip("the packed-taps kernel has no M sub-tiling")
    import irc_b200
    from irc_b200 import _native as nat
    nat.arch_check()
    if mt * min(n_out, 256) > 512:
        pytest.skip("mt * bn exceeds TMEM")
    g = torch.Generator(device="cuda").manual_seed(rows + n_out)
    A = torch.randn(rows, ld, device="cuda", generator=g).bfloat16()
    W = (torch.randn(n_out, len(taps) * cin, device="cuda", generator=g) * 0.05).bfloat16()
    out = torch.full((rows, n_out), float("nan"), device="cuda", dtype=torch.float32 if fp32 else torch.bfloat16)
    a = nat.ConvGemmArgs()
    a.a = A.data_ptr(); a.a_rows = rows; a.a_ld = ld; a.a_chan_off = a_off; a.cin = cin
    a.ntaps = len(taps)
    for i, t in enumerate(taps):
        a.taps[i] = t
    a.w = W.data_ptr(); a.n_out = n_out
    a.out = out.data_ptr(); a.out_ld = n_out; a.out_chan_off = 0; a.out_fp32 = int(fp32)
    a.bias = None; a.act = 0; a.slope = 0.0; a.row_img = None; a.mask = None; a.bn = 0; a.mt = mt; a.reuse = reuse
    nat.check(nat.lib().irc_conv_gemm(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    ref = _conv_ref(A, taps, W, cin, a_off)
    err = (out.float() - ref).norm() / ref.norm()
    assert torch.isfinite(out.float()).all()
    assert err < (1e-5 if fp32 else 4e-3), err


def test_conv_gemm_epilogue():
    from irc_b200 import _native as nat
    rows, cin, n_out = 1500, 64, 64
    taps = [-1, 0, 1]
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(rows, cin, device="cuda", generator=g).bfloat16()
    W = (torch.randn(n_out, 3 * cin, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(n_out, device="cuda", generator=g)
    mask = torch.randn(rows, n_out, device="cuda", generator=g).bfloat16()
    img = (torch.arange(rows, device="cuda") // 500).short()
    img[::7] = -1
    addend = torch.randn(rows, 128, device="cuda", generator=g).bfloat16()
    for mode, reuse in [(m, r) for m in ("bias_lrelu_rows", "mask", "addend") for r in (0, 1, 2)]:
        out = torch.full((rows, n_out), float("nan"), device="cuda", dtype=torch.bfloat16)
        a = nat.ConvGemmArgs()
        a.reuse = reuse
        a.a = A.data_ptr(); a.a_rows = rows; a.a_ld = cin; a.a_chan_off = 0; a.cin = cin; a.ntaps = 3
        for i, t in enumerate(taps):
            a.taps[i] = t
        a.w = W.data_ptr(); a.n_out = n_out; a.out = out.data_ptr(); a.out_ld = n_out
        ref = _conv_ref(A, taps, W, cin, 0)
        if mode == "bias_lrelu_rows":
            a.bias = bias.data_ptr(); a.act = 2; a.slope = 0.2; a.row_img = img.data_ptr()
            ref = torch.nn.functional.leaky_relu(ref + bias, 0.2) * (img >= 0)[:, None]
        elif mode == "mask":
            a.mask = mask.data_ptr(); a.mask_ld = n_out; a.mask_chan_off = 0; a.mask_slope = 0.2
            ref = ref * torch.where(mask.float() > 0, 1.0, 0.2)
        else:
            a.addend = addend.data_ptr(); a.addend_ld = 128; a.addend_chan_off = 64
            ref = ref + addend[:, 64:].float()
        nat.check(nat.lib().irc_conv_gemm(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        err = (out.float() - ref).norm() / ref.norm()
        assert err < 4e-3, (mode, err)
        if mode == "bias_lrelu_rows":
            assert (out[img < 0] == 0).all()


@pytest.mark.parametrize("rows,m,n,lda,ldb,a_off,b_off,shifts,splits", [
    (4096, 128, 64, 128, 64, 0, 0, [0], 1),
    (5000, 256, 256, 256, 256, 0, 0, [-71, -70, -69, -1, 0, 1, 69, 70, 71], 3),
    (3333, 64, 192, 64, 192, 0, 0, [-1, 0, 1], 2),
    (2048, 21, 64, 64, 64, 0, 0, [-10, 0, 10], 4),
    (3000, 128, 128, 384, 384, 256, 128, [0, 5], 2),
    (6000, 128, 64, 128, 64, 0, 0, [-71, -70, -69, -1, 0, 1, 69, 70, 71], 5),       # 3 taps per CTA
    (6000, 21, 64, 64, 64, 0, 0, [-210, -140, -70, 0, 70, 140, 210], 4),            # 7 taps in one ragged group of 8
    (5000, 128, 64, 128, 64, 0, 0, [-35, -34, -33, -1, 1, 33, 34, 35], 3),           # 64-wide tiles, 4 taps per CTA: one N = 256 MMA per step
    (5000, 100, 50, 128, 64, 0, 0, [-36, -35, -34, 0, 34, 35, 36], 2),              # 7 taps: groups of 4 and 3 (N = 256 and N = 192), ragged m / n
    (4000, 512, 256, 512, 256, 0, 0, [r * 34 + s for r in range(4) for s in range(4)], 2),   # 2 taps per CTA, 4 M tiles
    (7000, 64, 192, 64, 192, 0, 0, [-259, -258, -257, -1, 0, 1, 257, 258, 259], 3),          # paired taps (m <= 64): 5 pairs, last one half empty
    (2500, 64, 128, 192, 320, 128, 64, [-3, 4], 1),                                          # paired taps, channel offsets, one pair
])
def test_tn_gemm(rows, m, n, lda, ldb, a_off, b_off, shifts, splits):
    from irc_b200 import _native as nat
    g = torch.Generator(device="cuda").manual_seed(rows + m)
    A = torch.randn(rows, lda, device="cuda", generator=g).bfloat16()
    B = torch.randn(rows, ldb, device="cuda", generator=g).bfloat16()
    nt = len(shifts)
    out = torch.full((splits, m, nt, n), float("nan"), device="cuda")
    a = nat.TnGemmArgs()
    a.a = A.data_ptr(); a.a_rows = rows; a.a_ld = lda; a.a_chan_off = a_off; a.m = m
    a.b = B.data_ptr(); a.b_rows = rows; a.b_ld = ldb; a.b_chan_off = b_off; a.n = n
    a.k_rows = rows; a.ntaps = nt
    for i, s in enumerate(shifts):
        a.a_shift[i] = 0; a.b_shift[i] = s
    a.out = out.data_ptr(); a.out_tap_stride = n; a.out_m_stride = nt * n; a.out_n_stride = 1
    a.out_split_stride = m * nt * n; a.splits = splits; a.bn = 0
    nat.check(nat.lib().irc_tn_gemm(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    got = out.sum(0)
    Af = A.float()[:, a_off:a_off + m]; Bf = B.float()[:, b_off:b_off + n]
    q = torch.arange(rows, device="cuda")
    for t, s in enumerate(shifts):
        idx = q + s
        ok = ((idx >= 0) & (idx < rows)).float()[:, None]
        ref = Af.t() @ (Bf[idx.clamp(0, rows - 1)] * ok)
        err = (got[:, t, :] - ref).norm() / ref.norm()
        assert err < 1e-4, (t, err)


@pytest.mark.parametrize("n_img,hp,wp,cin,n_out,taps", [
    (5, 34, 34, 64, 64, [-35, -34, -33, -1, 0, 1, 33, 34, 35]),      # 1156 rows per image: images start inside sub-tiles
    (4, 32, 32, 128, 128, [-1, 0, 1]),                               # 1024 rows per image: boundaries on sub-tile edges
    (3, 66, 66, 64, 256, [-67, -66, -65, -1, 0, 1, 65, 66, 67]),
    (2, 20, 13, 64, 192, [0, 1, 13, 14]),                            # 260 rows per image, three 64-wide column chunks
])
@pytest.mark.parametrize("mt", [0, 1, 2, 4])
def test_conv_gemm_epilogue_instance_norm_statistics(n_img, hp, wp, cin, n_out, taps, mt):
    """north star (1): InstanceNorm statistics from the GEMM epilogue == sums over the live rows of the stored bf16 output,
    per image and channel, for every tiling; bit-reproducible across launches"""
    from irc_b200._native import CudaBackend
    if mt * min(n_out, 256) > 512:
        pytest.skip("mt * bn exceeds TMEM")
    be = CudaBackend(); be.conv_mt = mt
    g = torch.Generator(device="cuda").manual_seed(n_img * 100 + n_out)
    rows = n_img * hp * wp
    A = torch.randn(rows, cin, device="cuda", generator=g).bfloat16()
    W = (torch.randn(n_out, len(taps) * cin, device="cuda", generator=g) * 0.05).bfloat16()
    ri = torch.zeros(rows, device="cuda", dtype=torch.int16)
    be.row_index(ri, n_img, hp, wp, 1, hp - 1, 1, wp - 1)
    out = torch.full((rows, n_out), float("nan"), device="cuda", dtype=torch.bfloat16)
    st = [torch.full((n_img, n_out, 2), float("nan"), device="cuda") for _ in range(2)]
    for s_ in st:
        be.conv_gemm(A, 0, cin, taps, W, n_out, out, row_img=ri, in_stats=(s_, n_img, hp * wp))
    torch.cuda.synchronize()
    assert torch.equal(st[0], st[1])
    ref = _conv_ref(A, taps, W, cin, 0) * (ri >= 0)[:, None]
    assert (out.float() - ref).norm() / ref.norm() < 4e-3
    o = out.float().view(n_img, hp * wp, n_out)
    want = torch.stack([o.sum(1), (o * o).sum(1)], -1)
    err = (st[0] - want).abs().max() / want.abs().max()
    assert err < 1e-5, err


@pytest.mark.parametrize("n_img,H,W", [(3, 20, 24), (1, 64, 40), (2, 37, 129)])
def test_conv_gemm_fused_horizontal_taps_bias_tanh(n_img, H, W):
    """output head (irc:527-531): GEMM over the 7 kernel rows with the 7-tap horizontal reduction, bias and tanh in the
    epilogue (overlapping tiles) == per-tap partial products + separate reduction"""
    from irc_b200._native import CudaBackend
    be = CudaBackend()
    g = torch.Generator(device="cuda").manual_seed(H * W)
    hp, wp = H + 6, W + 6
    rows = n_img * hp * wp
    A = torch.randn(rows, 64, device="cuda", generator=g).bfloat16()
    Wt = torch.zeros(32, 7 * 64, device="cuda")
    Wt[:21] = torch.randn(21, 7 * 64, device="cuda", generator=g) * 0.03
    Wt = Wt.bfloat16()
    bias = torch.randn(3, device="cuda", generator=g) * 0.1
    taps = [(r - 3) * wp for r in range(7)]
    P = torch.zeros(rows, 32, device="cuda")
    be.conv_gemm(A, 0, 64, taps, Wt, 32, P)
    want = torch.zeros(n_img, 3, H, W, device="cuda")
    be.tap_reduce(P, [(0, s - 3) for s in range(7)], 3, n_img, H, W, hp, wp, 3, 3, bias, 3, want)
    got = torch.full((n_img, 3, H, W), float("nan"), device="cuda")
    be.conv_gemm(A, 0, 64, taps, Wt, 32, torch.zeros(8, 32, device="cuda"), bias=bias,
                 tap=dict(out=got, nshift=7, nco=3, H=H, W=W, hp=hp, wp=wp, oy=3, ox=3, act=3))
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    assert (got - want).abs().max() < 2e-6, (got - want).abs().max()
