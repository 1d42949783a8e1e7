"""Teacher-forced per-layer parity at the REAL layer shapes (256 x 256 input, B = 1): every kernel of the generator's
forward and backward plan is fed the ORACLE's activations / gradients (rounded to the bf16 the frames store) and its
output is compared with the oracle's fp32 result on the same inputs.  Tolerance: the north star's per-layer figure for
bf16, rel-L2 <= 1e-2 (measured: 2e-3 .. 5e-3, i.e. bf16 rounding of the operands and of the stored output).

Why teacher-forced: one D+G iteration chains 24 convolutions with ReLU / sign() decisions in between; a whole-step
gradient comparison mixes every kernel's error with mask flips inherited from upstream (tests/test_step_gpu.py bounds
those loosely).  Here each kernel sees exactly the values the oracle saw, so a wrong tap, a missing fold on a border
or a mis-indexed weight shows up as an O(1) error on that layer alone.

The oracle side is local fp32 autograd of the reference's own operator sequence (oracle/irc_oracle.py primitives) on the
CPU.  All CUDA work goes through the C ABI via the engine's launch methods."""
import torch
import torch.nn.functional as F


class Cfg:
    """set by the test module that drives the checks"""
    TOL = 1e-2
    B, H, W = 1, 256, 256
    dev = "cuda"
    round_bf16 = True



def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape and b.norm() > 0, (a.shape, b.shape)          # a comparison of two empty / all-zero tensors proves nothing
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def r16(t):
    """the value a frame stores (bf16 on the GPU; the CPU restatement keeps float32 frames)"""
    return t.detach().to(torch.bfloat16).float() if Cfg.round_bf16 else t.detach().clone()


def put(fr, x, chan_off=0, ring="zero"):
    """NCHW fp32 -> interior of a frame (channels chan_off..), ring zero / reflected / untouched"""
    n, c, h, w = x.shape
    p = fr.p
    v = fr.t.view(fr.N, fr.hp, fr.wp, fr.C)
    if ring == "reflect" and p:
        x = F.pad(x, (p, p, p, p), mode="reflect")
        v[:, :, :, chan_off:chan_off + c].copy_(x.permute(0, 2, 3, 1))
    else:
        if ring == "zero" and p:
            v[:, :, :, chan_off:chan_off + c].zero_()
        v[:, p:p + h, p:p + w, chan_off:chan_off + c].copy_(x.permute(0, 2, 3, 1))


def put_full(fr, xp, chan_off=0):
    """NCHW fp32 of the PADDED size -> whole frame"""
    c = xp.shape[1]
    fr.t.view(fr.N, fr.hp, fr.wp, fr.C)[:, :, :, chan_off:chan_off + c].copy_(xp.permute(0, 2, 3, 1))


def get(fr, c, chan_off=0, full=False):
    v = fr.t.view(fr.N, fr.hp, fr.wp, fr.C)
    if not full:
        v = v[:, fr.p:fr.p + fr.H, fr.p:fr.p + fr.W]
    return v[:, :, :, chan_off:chan_off + c].permute(0, 3, 1, 2).float().cpu()


class Ctx:
    pass


def make_ctx(be):
    """oracle forward + backward of the full generator loss at 256 x 256 with every tap's gradient retained, and a
    generator engine holding the same weights"""
    import irc_oracle as O
    import irc_b200  # noqa: F401
    from irc_b200 import engine as E
    B, H, W = Cfg.B, Cfg.H, Cfg.W
    torch.manual_seed(0)
    pG = O.seeded_params(O.generator_shapes(), 1234, bias_std=0.02)
    pD = O.seeded_params(O.discriminator_shapes(), 1235, bias_std=0.02)
    pV = O.seeded_params(O.vgg_shapes(), 1236, kaiming=True, bias_std=0.05)
    ir, rgb = O.synthetic_pair(B, H, W)
    leaves = {k: v.clone().requires_grad_(True) for k, v in pG.items()}
    taps = {}
    fake = O.generator_forward(leaves, ir, taps=taps)
    for t in taps.values():
        t.retain_grad()
    lam = O.LAMBDAS
    total = (lam["gan"] * -O.discriminator_forward(pD, torch.cat([ir, fake], 1)).mean() + lam["L1"] * (fake - rgb).abs().mean()
             + lam["perc"] * (O.vgg_forward(pV, fake) - O.vgg_forward(pV, rgb)).abs().mean() + lam["tv"] * O.tv_loss(fake)
             + lam["ssim"] * O.ssim_loss((fake + 1) / 2, (rgb + 1) / 2))
    total.backward()
    c = Ctx()
    c.O, c.E = O, E
    c.p = pG
    c.ir = ir
    c.act = {k: v.detach() for k, v in taps.items()}
    c.grad = {k: v.grad.detach() for k, v in taps.items()}
    c.wgrad = {k: v.grad.detach() for k, v in leaves.items()}
    c.be = be
    c.eng = E.GeneratorEngine(c.be, B, H, W, Cfg.dev)
    c.eng.arena.load(pG)
    c.eng.refresh_weights()
    return c


def conv_local(x, w, dz, pad=0, reflect=0):
    """fp32 CPU: z = conv(x), and the vector-Jacobian products of dz"""
    x = x.clone().requires_grad_(True); w = w.clone().requires_grad_(True)
    xin = F.pad(x, (reflect,) * 4, mode="reflect") if reflect else x
    z = F.conv2d(xin, w, None, padding=pad)
    gx, gw = torch.autograd.grad(z, (x, w), dz)
    return z.detach(), gx, gw


def norm_act_local(z, g, O, act="relu", res=None):
    """a = act(IN(z)) (+res);  dz = vjp(g)"""
    z = z.clone().requires_grad_(True)
    a = O.instance_norm(z)
    if act == "relu":
        a = torch.relu(a)
    (dz,) = torch.autograd.grad(a, z, g)
    out = a.detach() + (res if res is not None else 0)
    return out, dz


def scaled(g):
    """unit-RMS copy of a gradient tensor (kernels are linear in it)"""
    return g / g.pow(2).mean().sqrt().clamp_min(1e-30)


def zero_wgrad(eng):
    eng.arena.grad.zero_()


# ----------------------------------------------------------------------------------------------------------
def check_inc_7x7_reflect(ctx):
    """inc.1: ReflectionPad2d(3) + Conv2d(1, 64, 7) (irc:458-463) forward and weight gradient; IN + ReLU apply"""
    B, H, W, TOL = Cfg.B, Cfg.H, Cfg.W, Cfg.TOL
    O, eng, be = ctx.O, ctx.eng, ctx.be
    w = ctx.p["inc.1.weight"]
    ir = ctx.ir
    z_ref, _, _ = conv_local(ir, r16(w), torch.zeros(B, 64, H, W), reflect=3)
    be.im2col(ir.to(Cfg.dev), None, None, None, B, H, W, 7, 1, 3, 1, H, W, 0, eng.E_in)
    eng.inc.fwd(eng.E_in, 0, eng.Z0.t)
    e = rel(get(eng.Z0, 64), z_ref)
    print("inc fwd", e); assert e < TOL
    # IN + ReLU into cat2[128:192) with a zero ring
    z = r16(z_ref)
    put(eng.Z0, z)
    be.in_stats(eng.Z0.view(), 64, B, H, W, eng.st0)
    be.gather(eng.Z0.view(), eng.cat2.view(128), 64, B, H, W, 1, 0, stats=eng.st0, cnt=H * W, eps=1e-5, act=1)
    a_ref, _ = norm_act_local(z, torch.zeros_like(z), O)
    e = rel(get(eng.cat2, 64, 128), a_ref)
    print("inc IN+ReLU", e); assert e < TOL
    ring = get(eng.cat2, 64, 128, full=True)
    assert ring[:, :, 0].abs().max() == 0 and ring[:, :, :, -1].abs().max() == 0
    # backward: IN+ReLU^T with two gradient sources (skip connection + down1), then the weight gradient
    g = scaled(ctx.grad["x0"])
    ga, gb = r16(g * 0.25), r16(g * 0.75)
    put(eng.Gcat2, ga, 128); put(eng.Gx0, gb)
    be.in_bwd(eng.Z0.view(), eng.Gcat2.view(128), eng.dZ0.view(), 64, B, H, W, stats=eng.st0, cnt=H * W, eps=1e-5, act=1,
              g2=eng.Gx0.view(), bsum=eng.bsum)
    _, dz_ref = norm_act_local(z, ga + gb, O)
    e = rel(get(eng.dZ0, 64), dz_ref)
    print("inc IN+ReLU bwd (2 sources)", e); assert e < TOL
    dz = r16(dz_ref)
    put(eng.dZ0, dz)
    zero_wgrad(eng)
    eng.inc.wgrad(eng.dZ0.t, eng.E_in, 0, eng.dZ0.rows); be.flush_sums()
    _, _, gw = conv_local(r16(ir), w, dz, reflect=3)
    e = rel(eng.arena.view("inc.1.weight", eng.arena.grad), gw)
    print("inc wgrad", e); assert e < TOL


def check_3x3_zero_pad_layers(ctx, name):
    """down1.0 / down2.0 / up1_conv.0 / up2_conv.0 (irc:469-524): conv forward, IN+ReLU fused with the following
    Downsample / UpsampleAA stencil (or the reflect-3 ring for up2), their transposes, IN backward, dgrad, wgrad"""
    B, H, W, TOL = Cfg.B, Cfg.H, Cfg.W, Cfg.TOL
    O, eng, be = ctx.O, ctx.eng, ctx.be
    A, G = ctx.act, ctx.grad
    H2, W2, H4, W4 = H // 2, W // 2, H // 4, W // 4
    cfg = {
        # conv op, input frame, channel offset, input tensor, Z frame, stats, output channels, (h, w) of the conv
        "down1": (eng.down1, eng.cat2, 128, A["x0"], eng.Z1, eng.st1, 128, (H, W), "down1.0.weight"),
        "down2": (eng.down2, eng.cat1, 256, A["x1"], eng.Z2, eng.st2, 256, (H2, W2), "down2.0.weight"),
        "up1": (eng.up1, eng.cat1, 0, torch.cat([A["up1_up"], A["x1"]], 1), eng.Z3, eng.st3, 128, (H2, W2), "up1_conv.0.weight"),
        "up2": (eng.up2, eng.cat2, 0, torch.cat([A["up2_up"], A["x0"]], 1), eng.Z4, eng.st4, 64, (H, W), "up2_conv.0.weight"),
    }[name]
    op, fin, off, x, Zf, st, co, (h, w_), wname = cfg
    wt = ctx.p[wname]
    x = r16(x)
    ci = x.shape[1]
    fin.t.zero_()
    put(fin, x, off)
    op.fwd(fin.t, off, Zf.t)
    z_ref, _, _ = conv_local(x, r16(wt), torch.zeros(B, co, h, w_), pad=1)
    e = rel(get(Zf, co), z_ref)
    print(name, "fwd", e); assert e < TOL
    # ---- IN + ReLU + what follows it in the plan
    z = r16(z_ref)
    put(Zf, z)
    be.in_stats(Zf.view(), co, B, h, w_, st)
    if name == "down1":
        be.gather(Zf.view(), eng.cat1.view(256), co, B, H2, W2, 1, 0, tables=eng.t_down1, stats=st, cnt=h * w_, eps=1e-5, act=1)
        got = get(eng.cat1, co, 256); post = O.blur_down
    elif name == "down2":
        be.gather(Zf.view(), eng.X[0].view(), co, B, H4, W4, 1, 1, tables=eng.t_down2, stats=st, cnt=h * w_, eps=1e-5, act=1)
        got = get(eng.X[0], co); post = O.blur_down
    elif name == "up1":
        be.gather(Zf.view(), eng.cat2.view(0), co, B, H, W, 1, 0, tables=eng.t_up2, stats=st, cnt=h * w_, eps=1e-5, act=1)
        got = get(eng.cat2, co, 0); post = O.upsample_aa
    else:
        be.gather(Zf.view(), eng.y4.view(), co, B, H, W, 3, 1, stats=st, cnt=h * w_, eps=1e-5, act=1)
        got = get(eng.y4, co); post = lambda t: t
    zz = z.clone().requires_grad_(True)
    a = torch.relu(O.instance_norm(zz))
    out_ref = post(a)
    e = rel(got, out_ref)
    print(name, "IN+ReLU+stencil fwd", e); assert e < TOL
    if name == "down2":       # reflect ring of the bottleneck input
        full = get(eng.X[0], co, full=True)
        assert rel(full, F.pad(got, (1, 1, 1, 1), mode="reflect")) < 1e-6
    if name == "up2":
        full = get(eng.y4, co, full=True)
        assert rel(full, F.pad(got, (3, 3, 3, 3), mode="reflect")) < 1e-6
    # ---- backward of the same group: stencil^T (or reflection fold), IN+ReLU^T
    key = {"down1": "x1", "down2": "x2", "up1": "up2_up", "up2": "up2"}[name]
    g_out = r16(scaled(G[key]))
    if name == "down1":
        ga, gb = r16(g_out * 0.5), r16(g_out * 0.5)
        put(eng.Gcat1, ga, 256); put(eng.Gx1, gb)
        be.gather(eng.Gcat1.view(256), eng.g1.view(), co, B, H, W, 0, 0, tables=eng.t_down1_T, src2=eng.Gx1.view())
        gfr = eng.g1; g_out = ga + gb
    elif name == "down2":
        gfr = eng.g2
        cur = eng.dOut[0]
        put(cur, g_out)
        be.gather(cur.view(), gfr.view(), co, B, H2, W2, 0, 0, tables=eng.t_down2_T)
    elif name == "up1":
        put(eng.Gcat2, g_out, 0)
        gfr = eng.g3
        be.gather(eng.Gcat2.view(0), gfr.view(), co, B, H2, W2, 0, 0, tables=eng.t_up2_T)
    else:
        # gradient w.r.t. the reflect-padded y4 arrives over the whole frame; ReflectionPad2d(3)^T folds it
        gp = r16(F.pad(g_out, (3, 3, 3, 3)) + 0.3 * scaled(torch.randn(B, co, H + 6, W + 6, generator=torch.Generator().manual_seed(5))))
        put_full(eng.G4, gp)
        be.fold_inplace(eng.G4.t, 0, co, B, H, W, 3)
        gfr = eng.G4
        y = torch.zeros(B, co, H, W, requires_grad=True)
        (g_out,) = torch.autograd.grad(F.pad(y, (3, 3, 3, 3), mode="reflect"), y, gp)
    (g_ref,) = torch.autograd.grad(out_ref, a, g_out, retain_graph=True) if name != "up2" else (g_out,)
    e = rel(get(gfr, co), g_ref)
    print(name, "stencil^T / fold", e); assert e < TOL
    g_in = r16(get(gfr, co))              # what the plan hands to the InstanceNorm backward
    dZ = {"down1": eng.dZ1, "down2": eng.dZ2, "up1": eng.dZ3, "up2": eng.dZ4}[name]
    be.in_bwd(Zf.view(), gfr.view(), dZ.view(), co, B, h, w_, stats=st, cnt=h * w_, eps=1e-5, act=1, bsum=eng.bsum)
    (dz_ref,) = torch.autograd.grad(a, zz, g_in)
    e = rel(get(dZ, co), dz_ref)
    print(name, "IN+ReLU bwd", e); assert e < TOL
    # ---- conv data / weight gradients
    dz = r16(dz_ref)
    put(dZ, dz)
    zero_wgrad(eng)
    op.wgrad(dZ.t, fin.t, off, fin.rows); be.flush_sums()
    Gout = {"down1": eng.Gx0, "down2": eng.Gx1, "up1": eng.Gcat1, "up2": eng.Gcat2}[name]
    op.dgrad(dZ.t, Gout.t)
    _, gx, gw = conv_local(x, wt, dz, pad=1)
    e = rel(eng.arena.view(wname, eng.arena.grad), gw)
    print(name, "wgrad", e); assert e < TOL
    _, gx16, _ = conv_local(x, r16(wt), dz, pad=1)
    e = rel(get(Gout, ci), gx16)
    print(name, "dgrad", e); assert e < TOL


def check_resnet_block(ctx, b):
    """ResnetBlock (irc:362-418) at 256 ch, 64 x 64: both convs forward, cluster InstanceNorm forward (stats + apply +
    residual + reflect ring), and the whole backward chain with the reflection folds deferred to the consumer"""
    B, H, W, TOL = Cfg.B, Cfg.H, Cfg.W, Cfg.TOL
    O, eng, be = ctx.O, ctx.eng, ctx.be
    A, G = ctx.act, ctx.grad
    H4, W4 = H // 4, W // 4
    xin = r16(A["x2"] if b == 0 else A[f"res{b - 1}"])
    w1, w2 = ctx.p[f"resblocks.{b}.conv_block.1.weight"], ctx.p[f"resblocks.{b}.conv_block.5.weight"]
    c1, c2 = eng.res[b]
    put(eng.X[b], xin, ring="reflect")
    c1.fwd(eng.X[b].t, 0, eng.Za[b].t)
    za_ref, _, _ = conv_local(xin, r16(w1), torch.zeros(B, 256, H4, W4), reflect=1)
    e = rel(get(eng.Za[b], 256), za_ref); print("res conv1 fwd", e); assert e < TOL
    za = r16(za_ref); put(eng.Za[b], za)
    be.in_apply(eng.Za[b].view(), eng.Hh[b].view(), 256, B, H4, W4, 1, 1, eng.sta[b], eps=1e-5, act=1)
    zza = za.clone().requires_grad_(True)
    hh = torch.relu(O.instance_norm(zza))
    e = rel(get(eng.Hh[b], 256, full=True), F.pad(hh.detach(), (1, 1, 1, 1), mode="reflect")); print("res IN+ReLU (ring incl.)", e); assert e < TOL
    h16 = r16(hh); put(eng.Hh[b], h16, ring="reflect")
    c2.fwd(eng.Hh[b].t, 0, eng.Zb[b].t)
    zb_ref, _, _ = conv_local(h16, r16(w2), torch.zeros(B, 256, H4, W4), reflect=1)
    e = rel(get(eng.Zb[b], 256), zb_ref); print("res conv2 fwd", e); assert e < TOL
    zb = r16(zb_ref); put(eng.Zb[b], zb)
    be.in_apply(eng.Zb[b].view(), eng.X[b + 1].view(), 256, B, H4, W4, 1, 1, eng.stb[b], eps=1e-5, act=0, res=eng.X[b].view())
    zzb = zb.clone().requires_grad_(True)
    nb_ = O.instance_norm(zzb)
    e = rel(get(eng.X[b + 1], 256), xin + nb_.detach()); print("res IN + residual", e); assert e < TOL
    # ---- backward: gradient w.r.t. the padded block output arrives with its ring (unfolded)
    gen = torch.Generator().manual_seed(40 + b)
    g_int = scaled(G[f"res{b}"])
    gp = r16(F.pad(g_int, (1, 1, 1, 1)) + 0.2 * torch.randn(B, 256, H4 + 2, W4 + 2, generator=gen))
    cur, nxt = eng.dOut[0], eng.dOut[1]
    put_full(cur, gp)
    y = torch.zeros(B, 256, H4, W4, requires_grad=True)
    (g_fold,) = torch.autograd.grad(F.pad(y, (1, 1, 1, 1), mode="reflect"), y, gp)
    g_fold = r16(g_fold)          # the fused kernel rounds the folded gradient to bf16 like the stand-alone fold does
    be.in_bwd(eng.Zb[b].view(), cur.view(), eng.dZb.view(), 256, B, H4, W4, stats=eng.stb[b], cnt=H4 * W4, eps=1e-5, act=0, bsum=eng.bsum, fold_pad=1)
    (dzb_ref,) = torch.autograd.grad(nb_, zzb, g_fold)
    e = rel(get(eng.dZb, 256), dzb_ref); print("res IN bwd + fold", e); assert e < TOL
    dzb = r16(dzb_ref); put(eng.dZb, dzb)
    zero_wgrad(eng)
    c2.wgrad(eng.dZb.t, eng.Hh[b].t, 0, eng.dZb.rows)
    c2.dgrad(eng.dZb.t, eng.Gh.t)
    hp = F.pad(h16, (1, 1, 1, 1), mode="reflect").requires_grad_(True)
    wv = w2.clone().requires_grad_(True)
    zloc = F.conv2d(hp, wv)
    (gw2,) = torch.autograd.grad(zloc, wv, dzb, retain_graph=True)
    (ghp,) = torch.autograd.grad(F.conv2d(hp, r16(w2)), hp, dzb)
    e = rel(get(eng.Gh, 256, full=True), ghp); print("res conv2 dgrad (frame)", e); assert e < TOL
    gh16 = r16(get(eng.Gh, 256, full=True))
    (gh_fold,) = torch.autograd.grad(F.pad(y, (1, 1, 1, 1), mode="reflect"), y, gh16)
    gh_fold = r16(gh_fold)
    be.in_bwd(eng.Za[b].view(), eng.Gh.view(), eng.dZa.view(), 256, B, H4, W4, stats=eng.sta[b], cnt=H4 * W4, eps=1e-5, act=1, bsum=eng.bsum, fold_pad=1)
    (dza_ref,) = torch.autograd.grad(hh, zza, gh_fold)
    e = rel(get(eng.dZa, 256), dza_ref); print("res IN+ReLU bwd + fold", e); assert e < TOL
    dza = r16(dza_ref); put(eng.dZa, dza)
    c1.wgrad(eng.dZa.t, eng.X[b].t, 0, eng.dZa.rows)
    from irc_b200._native import View
    c1.dgrad(eng.dZa.t, nxt.t, addend=View(cur.t, 0, 0, 0))
    be.flush_sums()
    xp = F.pad(xin, (1, 1, 1, 1), mode="reflect").requires_grad_(True)
    w1v = w1.clone().requires_grad_(True)
    (gw1,) = torch.autograd.grad(F.conv2d(xp, w1v), w1v, dza)
    (gxp,) = torch.autograd.grad(F.conv2d(xp, r16(w1)), xp, dza)
    e = rel(eng.arena.view(f"resblocks.{b}.conv_block.5.weight", eng.arena.grad), gw2); print("res conv2 wgrad", e); assert e < TOL
    e = rel(eng.arena.view(f"resblocks.{b}.conv_block.1.weight", eng.arena.grad), gw1); print("res conv1 wgrad", e); assert e < TOL
    # the residual-stream gradient may be added folded or unfolded (fold(a + b) = fold(a) + fold(b); the single-pass cluster
    # kernel folds on load and leaves `cur` untouched, the two-pass path folds it in place): compare after the fold
    refold = lambda t: torch.autograd.grad(F.pad(y, (1, 1, 1, 1), mode="reflect"), y, t)[0]
    e = rel(refold(get(nxt, 256, full=True)), refold(gxp + gp)); print("res conv1 dgrad + residual gradient (folded)", e); assert e < TOL
    # leaving the blocks: one in-place fold of the residual-stream gradient
    nx16 = r16(get(nxt, 256, full=True))
    be.fold_inplace(nxt.t, 0, 256, B, H4, W4, 1)
    (fold_ref,) = torch.autograd.grad(F.pad(y, (1, 1, 1, 1), mode="reflect"), y, nx16)
    e = rel(get(nxt, 256), fold_ref); print("fold_inplace(1)", e); assert e < TOL


def check_upsample_into_concat_and_transpose(ctx):
    """up1_up (irc:500-501): UpsampleAA of the bottleneck output into cat1[0:256), and its transpose"""
    B, H, W, TOL = Cfg.B, Cfg.H, Cfg.W, Cfg.TOL
    O, eng, be = ctx.O, ctx.eng, ctx.be
    H2, W2, H4, W4 = H // 2, W // 2, H // 4, W // 4
    x = r16(ctx.act["res8"])
    put(eng.X[9], x, ring="reflect")
    be.gather(eng.X[9].view(), eng.cat1.view(0), 256, B, H2, W2, 1, 0, tables=eng.t_up1)
    xx = x.clone().requires_grad_(True)
    up = O.upsample_aa(xx)
    e = rel(get(eng.cat1, 256, 0), up); print("UpsampleAA 64->128", e); assert e < TOL
    g = r16(scaled(ctx.grad["up1_up"]))
    put(eng.Gcat1, g, 0)
    be.gather(eng.Gcat1.view(0), eng.dOut[0].view(), 256, B, H4, W4, 1, 0, tables=eng.t_up1_T)
    (gr,) = torch.autograd.grad(up, xx, g)
    e = rel(get(eng.dOut[0], 256), gr); print("UpsampleAA^T", e); assert e < TOL


def check_outc_7x7_tanh(ctx):
    """outc.1 + tanh (irc:527-531): forward, tanh' + bias gradient, weight and data gradients"""
    B, H, W, TOL = Cfg.B, Cfg.H, Cfg.W, Cfg.TOL
    O, eng, be = ctx.O, ctx.eng, ctx.be
    w, bias = ctx.p["outc.1.weight"], ctx.p["outc.1.bias"]
    x = r16(ctx.act["up2"])
    put(eng.y4, x, ring="reflect")
    eng.output_head()
    xx = x.clone().requires_grad_(True)
    wv = r16(w).requires_grad_(True)
    bv = bias.clone().requires_grad_(True)
    out = torch.tanh(F.conv2d(F.pad(xx, (3, 3, 3, 3), mode="reflect"), wv, bv))
    e = rel(eng.fake, out); print("outc fwd + tanh", e); assert e < 2e-3          # fp32 accumulate, fp32 output
    g = scaled(ctx.grad["out"]).contiguous()
    zero_wgrad(eng)
    # tanh' uses the engine's own output (what the plan does); feed it the oracle's so that both sides see one value
    eng.fake.copy_(out.detach().to(Cfg.dev))
    be.tap_expand(g.to(Cfg.dev), eng.fake, eng.outc_shifts, 3, B, H, W, eng.y4.hp, eng.y4.wp, 3, 3, eng.E_out,
                  dbias=eng.arena.view("outc.1.bias", eng.arena.grad))
    eng.outc.wgrad(eng.E_out, eng.y4.t, 0, eng.y4.rows)
    eng.outc.dgrad(eng.E_out, eng.G4.t, k_live=21)       # as in the plan: only the 21 tap columns of E_out are live
    be.flush_sums()
    xp = F.pad(x, (3, 3, 3, 3), mode="reflect").requires_grad_(True)
    w32 = w.clone().requires_grad_(True)
    o2 = torch.tanh(F.conv2d(xp, w32, bv))
    gw, gb = torch.autograd.grad(o2, (w32, bv), g, retain_graph=True)
    (gxp,) = torch.autograd.grad(torch.tanh(F.conv2d(xp, r16(w), bias)), xp, g)
    e = rel(eng.arena.view("outc.1.bias", eng.arena.grad), gb); print("outc bias grad", e); assert e < TOL
    e = rel(eng.arena.view("outc.1.weight", eng.arena.grad), gw); print("outc wgrad", e); assert e < TOL
    e = rel(get(eng.G4, 64, full=True), gxp); print("outc dgrad (frame)", e); assert e < TOL


# ==========================================================================================================
# PatchGAN discriminator and VGG trunk: the same per-layer teacher forcing (every launch of
# DiscriminatorEngine.forward / .backward and VggEngine.forward / .backward, fed the oracle's tensors)
# ==========================================================================================================
def putv(v, x):
    """NCHW fp32 -> pixels (0..h, 0..w) of a View (plain or space-to-depth), other elements untouched"""
    from ref_backend import RefBackend
    n, c, h, w = x.shape
    dev = v.t.device
    RefBackend()._write(v, n, torch.arange(h, device=dev), torch.arange(w, device=dev), c, x.permute(0, 2, 3, 1).to(dev))


def getv(v, n, c, h, w):
    from ref_backend import RefBackend
    dev = v.t.device
    return RefBackend()._read(v, n, torch.arange(h, device=dev), torch.arange(w, device=dev), c).permute(0, 3, 1, 2).float().cpu()


def conv_s(x, w, dz, stride=1, pad=0, bias=None):
    """fp32 CPU strided conv and its vector-Jacobian products (dz may be None: forward only)"""
    x = x.clone().requires_grad_(True); w = w.clone().requires_grad_(True)
    z = F.conv2d(x, w, bias, stride=stride, padding=pad)
    if dz is None:
        return z.detach(), None, None
    gx, gw = torch.autograd.grad(z, (x, w), dz)
    return z.detach(), gx, gw


def in_lrelu_local(z, g, O):
    z = z.clone().requires_grad_(True)
    a = F.leaky_relu(O.instance_norm(z), 0.2)
    dz = torch.autograd.grad(a, z, g)[0] if g is not None else None
    return a.detach(), dz


def make_dv_ctx(be):
    """oracle forward / backward of the PatchGAN on (ir, rgb) + (ir, fake) and of the VGG trunk on (fake, rgb), every tap's
    gradient retained; engines with the same weights"""
    import irc_oracle as O
    import irc_b200  # noqa: F401
    from irc_b200 import engine as E
    B, H, W = Cfg.B, Cfg.H, Cfg.W
    pD = O.seeded_params(O.discriminator_shapes(), 1235, bias_std=0.02)
    pV = O.seeded_params(O.vgg_shapes(), 1236, kaiming=True, bias_std=0.05)
    ir, rgb = O.synthetic_pair(B, H, W)
    _, fake = O.synthetic_pair(B, H, W, rank=3)
    fake = (0.6 * fake + 0.4 * rgb).contiguous()
    c = Ctx()
    c.O, c.E, c.be = O, E, be
    c.pD, c.pV, c.ir, c.rgb, c.fake = pD, pV, ir, rgb, fake
    # ---- discriminator
    xd = torch.cat([torch.cat([ir, rgb], 1), torch.cat([ir, fake], 1)], 0).requires_grad_(True)
    taps = {}
    pred = O.discriminator_forward(pD, xd, taps=taps)
    for t in taps.values():
        t.retain_grad()
    c.g_pred = scaled(torch.randn(pred.shape, generator=torch.Generator().manual_seed(11)))
    pred.backward(c.g_pred)
    c.xd = xd.detach()
    c.dact = {k: v.detach() for k, v in taps.items()}
    c.dgrad = {k: v.grad.detach() for k, v in taps.items()}
    c.deng = E.DiscriminatorEngine(be, 2 * B, H, W, Cfg.dev)
    c.deng.arena.load(pD)
    c.deng.refresh_weights()
    # ---- VGG
    c.veng = E.VggEngine(be, 2 * B, B, H, W, Cfg.dev)
    c.veng.arena.load(pV)
    c.veng.refresh_weights()
    return c


def check_discriminator_layers(ctx):
    """NLayerDiscriminator (irc:598-635): model.0 (direct 4x4 s2 conv + LeakyReLU), model.2 / .5 (stride-2 convs over
    space-to-depth blocks) and model.8 with InstanceNorm + LeakyReLU, model.11 (per-tap partial products + shifted reduction);
    backward: tap expansion, every data / weight / bias gradient, the three InstanceNorm + LeakyReLU transposes, the
    LeakyReLU mask of model.0 in the GEMM epilogue, col2im into the image gradient"""
    B, H, W, TOL = Cfg.B, Cfg.H, Cfg.W, Cfg.TOL
    O, E, be, eng, p = ctx.O, ctx.E, ctx.be, ctx.deng, ctx.pD
    from irc_b200._native import View
    n = 2 * B
    A, G = ctx.dact, ctx.dgrad
    dev = Cfg.dev
    LR = dict(eps=E.EPS, act=E.ACT_LRELU, slope=0.2)
    H1, W1, H2, W2, H3, W3, H8, W8, Ho, Wo = eng.H1, eng.W1, eng.H2, eng.W2, eng.H3, eng.W3, eng.H8o, eng.W8o, eng.Ho, eng.Wo
    vS0 = View(eng.S0v, 0, eng.hb0, eng.wb0, 1, 1, 64)
    vS2 = View(eng.S2, 0, eng.hb2, eng.wb2, 1, 1, 128)
    vdS2 = View(eng.dS2, 0, eng.hb2, eng.wb2, 1, 1, 128)
    vdZ0 = View(eng.dZ0v, 0, eng.hb0, eng.wb0, 1, 1, 64)
    wgt = lambda k: p[f"model.{k}.weight"]
    garena = lambda name: eng.arena.view(name, eng.arena.grad)
    ir, rgb, fake = ctx.ir.to(dev), ctx.rgb.to(dev), ctx.fake.to(dev)
    # ---- model.0 through the engine's own first launches (its input IS the oracle's input); fills E0 / row_img0
    pred = eng.forward(ir, rgb, ir, fake)
    x16 = r16(ctx.xd)
    d0_ref = F.leaky_relu(F.conv2d(x16, r16(wgt(0)), p["model.0.bias"], stride=2, padding=1), 0.2)
    e = rel(getv(vS0, n, 64, H1, W1), d0_ref); print("D.0 fwd + LeakyReLU (s2d rows)", e); assert e < TOL
    e = rel(pred, A["d11"]); print("D whole forward (5 layers, not teacher-forced)", e); assert e < 5 * TOL
    # ---- model.2: conv over the space-to-depth blocks, statistics (epilogue or pass), IN + LeakyReLU into the next s2d operand
    d0 = r16(A["d0"])
    eng.S0.zero_(); putv(vS0, d0)
    eng.c2.fwd_stats(eng.S0v, 0, eng.Z2, eng.st2, eng.ri2, n, eng.hb0 * eng.wb0, eng._vZ2(eng.Z2), 128, H2, W2)
    z2_ref, _, _ = conv_s(d0, r16(wgt(2)), None, 2, 1)
    e = rel(getv(eng._vZ2(eng.Z2), n, 128, H2, W2), z2_ref); print("D.2 fwd", e); assert e < TOL
    st_ref = torch.stack([z2_ref.sum((2, 3)), z2_ref.pow(2).sum((2, 3))], -1)
    e = rel(eng.st2[..., 1], st_ref[..., 1]); print("D.2 InstanceNorm sum of squares", e); assert e < TOL
    e = ((eng.st2[..., 0].cpu() - st_ref[..., 0]).abs().max() / st_ref[..., 1].sqrt().mean()).item(); print("D.2 InstanceNorm sums", e); assert e < TOL
    z2 = r16(z2_ref); putv(eng._vZ2(eng.Z2), z2)
    be.in_stats(eng._vZ2(eng.Z2), 128, n, H2, W2, eng.st2)
    be.gather(eng._vZ2(eng.Z2), View(eng.S2, 0, eng.hb2, eng.wb2), 128, n, H2, W2, 1, 0, dst_s2d=1, **eng._na(eng.st2, H2 * W2))
    a2_ref, _ = in_lrelu_local(z2, None, O)
    got = getv(vS2, n, 128, H2, W2)
    e = rel(got, a2_ref); print("D.2 IN + LeakyReLU -> s2d", e); assert e < TOL
    assert abs(eng.S2.float().abs().sum().item() - got.abs().sum().item()) <= 1e-3 * got.abs().sum().item()      # zero ring
    # ---- model.5
    d2 = r16(A["d2"])
    eng.S2.zero_(); putv(vS2, d2)
    eng.c5.fwd(eng.S2, 0, eng.Z5)
    z5_ref, _, _ = conv_s(d2, r16(wgt(5)), None, 2, 1)
    e = rel(getv(eng._vZ5(eng.Z5), n, 256, H3, W3), z5_ref); print("D.5 fwd", e); assert e < TOL
    z5 = r16(z5_ref); putv(eng._vZ5(eng.Z5), z5)
    be.in_apply(eng._vZ5(eng.Z5), eng.X8.view(), 256, n, H3, W3, 1, 0, eng.st5, **LR)
    a5_ref, _ = in_lrelu_local(z5, None, O)
    e = rel(get(eng.X8, 256, full=True), F.pad(a5_ref, (1, 1, 1, 1))); print("D.5 IN + LeakyReLU (zero ring incl.)", e); assert e < TOL
    # ---- model.8 (stride 1, k4 p1: the output shrinks by one)
    d5 = r16(A["d5"])
    put(eng.X8, d5)
    eng.c8.fwd(eng.X8.t, 0, eng.Z8)
    z8_ref, _, _ = conv_s(d5, r16(wgt(8)), None, 1, 1)
    e = rel(getv(eng._vZ8(eng.Z8), n, 512, H8, W8), z8_ref); print("D.8 fwd", e); assert e < TOL
    z8 = r16(z8_ref); putv(eng._vZ8(eng.Z8), z8)
    be.in_apply(eng._vZ8(eng.Z8), eng.X11.view(), 512, n, H8, W8, 1, 0, eng.st8, **LR)
    a8_ref, _ = in_lrelu_local(z8, None, O)
    e = rel(get(eng.X11, 512, full=True), F.pad(a8_ref, (1, 1, 1, 1))); print("D.8 IN + LeakyReLU (zero ring incl.)", e); assert e < TOL
    # ---- model.11: 512 -> 1
    d8 = r16(A["d8"])
    put(eng.X11, d8)
    eng.c11.fwd(eng.X11.t, 0, eng.P11)
    be.tap_reduce(eng.P11, eng.shifts11, 1, n, Ho, Wo, eng.X11.hp, eng.X11.wp, 0, 0, eng.c11.bias(), E.ACT_NONE, eng.pred)
    b11 = p["model.11.bias"].clone().requires_grad_(True)
    w11 = wgt(11).clone().requires_grad_(True)
    x11 = d8.clone().requires_grad_(True)
    e = rel(eng.pred, F.conv2d(d8, r16(wgt(11)), b11, padding=1)); print("D.11 fwd (tap partials + shifted sum)", e); assert e < TOL
    # ======== backward
    g = ctx.g_pred
    zero_wgrad(eng)
    be.tap_expand(g.to(dev).contiguous(), None, eng.shifts11, 1, n, Ho, Wo, eng.X11.hp, eng.X11.wp, 0, 0, eng.E11,
                  dbias=garena("model.11.bias"), live_cols_only=True)
    eng.c11.wgrad(eng.E11, eng.X11.t, 0, eng.X11.rows)
    eng.c11.dgrad(eng.E11, eng.G11.t)
    be.flush_sums()
    gw, gb = torch.autograd.grad(F.conv2d(d8, w11, b11, padding=1), (w11, b11), g)
    (gx,) = torch.autograd.grad(F.conv2d(x11, r16(wgt(11)), None, padding=1), x11, g)
    e = rel(garena("model.11.bias"), gb); print("D.11 bias grad", e); assert e < TOL
    e = rel(garena("model.11.weight"), gw); print("D.11 wgrad", e); assert e < TOL
    e = rel(get(eng.G11, 512), gx); print("D.11 dgrad", e); assert e < TOL
    # ---- model.8: IN + LeakyReLU transpose, weight / data gradients
    g8 = r16(scaled(G["d8"]))
    put(eng.G11, g8)
    be.in_bwd(eng._vZ8(eng.Z8), eng.G11.view(), eng._vZ8(eng.dZ8), 512, n, H8, W8, bsum=eng.bsum, **eng._nb(eng.st8, H8 * W8, True))
    _, dz8_ref = in_lrelu_local(z8, g8, O)
    e = rel(getv(eng._vZ8(eng.dZ8), n, 512, H8, W8), dz8_ref); print("D.8 IN + LeakyReLU bwd", e); assert e < TOL
    dz8 = r16(dz8_ref); eng.dZ8.zero_(); putv(eng._vZ8(eng.dZ8), dz8)
    zero_wgrad(eng)
    eng.c8.wgrad(eng.dZ8, eng.X8.t, 0, eng.X8.rows)
    eng.c8.dgrad(eng.dZ8, eng.G8.t)
    be.flush_sums()
    _, _, gw = conv_s(d5, wgt(8), dz8, 1, 1)
    _, gx, _ = conv_s(d5, r16(wgt(8)), dz8, 1, 1)
    e = rel(garena("model.8.weight"), gw); print("D.8 wgrad", e); assert e < TOL
    e = rel(get(eng.G8, 256), gx); print("D.8 dgrad", e); assert e < TOL
    # ---- model.5
    g5 = r16(scaled(G["d5"]))
    put(eng.G8, g5)
    be.in_bwd(eng._vZ5(eng.Z5), eng.G8.view(), eng._vZ5(eng.dZ5), 256, n, H3, W3, bsum=eng.bsum, **eng._nb(eng.st5, H3 * W3, True))
    _, dz5_ref = in_lrelu_local(z5, g5, O)
    e = rel(getv(eng._vZ5(eng.dZ5), n, 256, H3, W3), dz5_ref); print("D.5 IN + LeakyReLU bwd", e); assert e < TOL
    dz5 = r16(dz5_ref); eng.dZ5.zero_(); putv(eng._vZ5(eng.dZ5), dz5)
    zero_wgrad(eng)
    eng.c5.wgrad(eng.dZ5, eng.S2, 0, eng.S2.shape[0])
    eng.c5.dgrad(eng.dZ5, eng.dS2)
    be.flush_sums()
    _, _, gw = conv_s(d2, wgt(5), dz5, 2, 1)
    _, gx, _ = conv_s(d2, r16(wgt(5)), dz5, 2, 1)
    e = rel(garena("model.5.weight"), gw); print("D.5 wgrad", e); assert e < TOL
    e = rel(getv(vdS2, n, 128, H2, W2), gx); print("D.5 dgrad (s2d)", e); assert e < TOL
    # ---- model.2
    g2 = r16(scaled(G["d2"]))
    putv(vdS2, g2)
    be.in_bwd(eng._vZ2(eng.Z2), vdS2, eng._vZ2(eng.dZ2), 128, n, H2, W2, bsum=eng.bsum, **eng._nb(eng.st2, H2 * W2, True))
    _, dz2_ref = in_lrelu_local(z2, g2, O)
    e = rel(getv(eng._vZ2(eng.dZ2), n, 128, H2, W2), dz2_ref); print("D.2 IN + LeakyReLU bwd (s2d gradient)", e); assert e < TOL
    dz2 = r16(dz2_ref); eng.dZ2.zero_(); putv(eng._vZ2(eng.dZ2), dz2)
    zero_wgrad(eng)
    eng.c2.wgrad(eng.dZ2, eng.S0v, 0, eng.S0v.shape[0])
    eng.c2.dgrad(eng.dZ2, eng.dZ0v, mask=View(eng.S0v, 0, 0, 0), mask_slope=0.2)
    be.flush_sums()
    _, _, gw = conv_s(d0, wgt(2), dz2, 2, 1)
    _, gx, _ = conv_s(d0, r16(wgt(2)), dz2, 2, 1)
    e = rel(garena("model.2.weight"), gw); print("D.2 wgrad", e); assert e < TOL
    dz0_ref = gx * torch.where(d0 > 0, torch.ones_like(d0), torch.full_like(d0, 0.2))
    e = rel(getv(vdZ0, n, 64, H1, W1), dz0_ref); print("D.2 dgrad + LeakyReLU mask of model.0 (s2d)", e); assert e < TOL
    # ---- model.0: bias / weight gradient from the saved im2col operand, image gradient through col2im
    dz0 = r16(dz0_ref); eng.dZ0.zero_(); putv(vdZ0, dz0)
    zero_wgrad(eng)
    be.colsum(eng.dZ0, 0, 64, garena("model.0.bias"), row_img=eng.row_img0)
    eng.c0.wgrad(eng.dZ0, eng.E0, 0, eng.rows0)
    be.flush_sums()
    dimg = torch.zeros(n, 3, H, W, device=dev)
    eng.c0.dgrad(eng.dZ0, eng.dE0)
    be.col2im(eng.dE0, 4, 1, 3, n, H, W, 4, 2, 1, H1, W1, 2, None, dimg, False)
    _, _, gw = conv_s(x16, wgt(0), dz0, 2, 1)
    _, gx, _ = conv_s(x16, r16(wgt(0)), dz0, 2, 1)
    e = rel(garena("model.0.bias"), dz0.sum((0, 2, 3))); print("D.0 bias grad", e); assert e < TOL
    e = rel(garena("model.0.weight"), gw); print("D.0 wgrad", e); assert e < TOL
    e = rel(dimg, gx[:, 1:4]); print("D.0 dgrad -> image gradient", e); assert e < TOL


def check_vgg_layers(ctx):
    """VGGPerceptual trunk features[:16] (irc:642-683): conv1_1 straight from the fp32 image with the ImageNet normalisation
    folded in, six 3x3 convolutions with bias + ReLU in the epilogue, two max-pools; backward (data gradients of the first
    n_bwd images only): data-gradient GEMMs with the ReLU mask in the epilogue, max-pool transposes, conv1_1's gradient
    summed into d(fake)"""
    B, H, W, TOL = Cfg.B, Cfg.H, Cfg.W, Cfg.TOL
    O, E, be, eng, p = ctx.O, ctx.E, ctx.be, ctx.veng, ctx.pV
    from irc_b200._native import View
    dev = Cfg.dev
    n = 2 * B
    x = torch.cat([ctx.fake, ctx.rgb], 0)
    eng.forward(ctx.fake.to(dev), ctx.rgb.to(dev))
    mean = torch.tensor(O.IMAGENET_MEAN).view(1, 3, 1, 1); std = torch.tensor(O.IMAGENET_STD).view(1, 3, 1, 1)
    h_in = ((x + 1) / 2 - mean) / std
    gen = torch.Generator().manual_seed(21)
    acts = []            # r16 of every conv's ReLU output (pre-pool), all n images
    for i, (idx, ci, co) in enumerate(E.VGG_CFG):
        conv, fr = eng.convs[i], eng.act[i]
        w, b = p[f"features.{idx}.weight"], p[f"features.{idx}.bias"]
        hi = r16(h_in)
        a_ref = torch.relu(F.conv2d(hi, r16(w), b, padding=1))
        if i > 0:
            src = eng.pool[i - 1] if (i - 1) in eng.pool else eng.act[i - 1]
            put(src, hi)
            conv.fwd(src.t, 0, fr.t, bias=conv.bias(), act=E.ACT_RELU, row_img=None if i in eng.pool else eng._ri(i))
        if i in eng.pool:      # feeds the max-pool only: the ring of this frame is never read, the plan leaves it unmasked
            e = rel(get(fr, co), a_ref); print(f"V.{idx} fwd + bias + ReLU (interior)", e); assert e < TOL
        else:
            e = rel(get(fr, co, full=True), F.pad(a_ref, (1, 1, 1, 1))); print(f"V.{idx} fwd + bias + ReLU (zero ring incl.)", e); assert e < TOL
        a16 = r16(a_ref)
        acts.append(a16)
        put(fr, a16)
        if i in eng.pool:
            h, w_ = eng.res[i + 1]
            be.maxpool2(fr.view(), eng.pool[i].view(), fr.C, n, h, w_)
            h_in = F.max_pool2d(a16, 2, 2)
            e = rel(get(eng.pool[i], co, full=True), F.pad(h_in, (1, 1, 1, 1))); print(f"V.{idx} max-pool", e); assert e < 1e-6
        else:
            h_in = a16
    # ======== backward on the first B images
    nb = B
    dfake = torch.zeros(nb, 3, H, W, device=dev)
    for i in range(len(eng.convs) - 1, -1, -1):
        idx, ci, co = E.VGG_CFG[i]
        conv = eng.convs[i]
        w = p[f"features.{idx}.weight"]
        a_i = acts[i][:nb]
        dz = r16(scaled(torch.randn(a_i.shape, generator=gen)) * (a_i > 0))        # gradient w.r.t. the pre-ReLU output of conv i
        put(eng.dz[i], dz)
        if i == 0:
            hi = r16(((x[:nb] + 1) / 2 - mean) / std)
            _, gx, _ = conv_s(hi, r16(w), dz, 1, 1)
            ref = gx * (0.5 / std)
            if eng.fused_dgrad0:
                d0 = eng.dz[0]
                be.conv_gemm(d0.t, 0, 64, [d0.wp, 0, -d0.wp], eng.w0_dg.t, 32, eng.P0,
                             tap=dict(out=dfake, nshift=3, nco=3, H=H, W=W, hp=d0.hp, wp=d0.wp, oy=1, ox=1, act=E.ACT_NONE, scale=eng.scale, accumulate=True))
            else:
                conv.dgrad(eng.dz[0].t, eng.dE)
                be.col2im(eng.dE, 3, 0, 3, nb, H, W, 3, 1, 1, H, W, 1, eng.scale, dfake, True)
            e = rel(dfake, ref); print("V.0 dgrad -> d(fake)", e); assert e < TOL
            break
        prev = eng.act[i - 1]
        a_prev = acts[i - 1][:nb]
        if (i - 1) in eng.pool:
            pooled = a_prev.clone().requires_grad_(True)
            hin = F.max_pool2d(pooled, 2, 2)
            _, gx, _ = conv_s(hin.detach(), r16(w), dz, 1, 1)
            conv.dgrad(eng.dz[i].t, eng.dpool[i - 1].t)
            e = rel(get(eng.dpool[i - 1], ci), gx); print(f"V.{idx} dgrad (to the pooled map)", e); assert e < TOL
            gx16 = r16(gx); put(eng.dpool[i - 1], gx16)
            h, w_ = eng.res[i]
            be.maxpool2_bwd(View(prev.t[:prev.rows_of(nb)], 0, prev.hp, prev.wp, 1, 1), eng.dpool[i - 1].view(), eng.dz[i - 1].view(), prev.C, nb, h, w_)
            (ref,) = torch.autograd.grad(hin, pooled, gx16)
            ref = ref * (a_prev > 0)
            e = rel(get(eng.dz[i - 1], ci), ref); print(f"V.{idx} max-pool^T + ReLU mask", e); assert e < 1e-6
        else:
            _, gx, _ = conv_s(a_prev, r16(w), dz, 1, 1)
            conv.dgrad(eng.dz[i].t, eng.dz[i - 1].t, mask=View(prev.t[:prev.rows_of(nb)], 0, 0, 0), mask_slope=0.0)
            e = rel(get(eng.dz[i - 1], ci), gx * (a_prev > 0)); print(f"V.{idx} dgrad + ReLU mask", e); assert e < TOL
