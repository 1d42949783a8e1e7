"""Non-default generator graphs: no_antialias_up=True - nn.ConvTranspose2d up-sampling (irc:495-499, :512-516) - and
no_antialias=True - stride-2 down-sampling convolutions without the blur modules (irc:468, :474, :482) - and both together,
against golden vectors produced by the unmodified reference (tests/golden/ref_variants.npz, oracle/make_golden_variants.py).
CPU: the oracle and the product's plan (torch restatement of the primitives, float32 frames).  GPU: the CUDA path, incl. the
transposed convolution through its C-ABI entry point irc_convT2d_fwd."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import irc_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_variants.npz"))
B, H, W = 2, 32, 32


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm()).item()


def sample(t, n=256):
    f = t.detach().float().cpu().reshape(-1)
    return f[torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()].numpy()


def _params():
    return O.seeded_params(O.generator_shapes(no_antialias_up=True), 4321, bias_std=0.02)


def test_oracle_variant_forward_and_grads():
    pG = _params()
    ir, _ = O.synthetic_pair(B, H, W)
    leaves = {k: v.clone().requires_grad_(True) for k, v in pG.items()}
    fake = O.generator_forward(leaves, ir)
    assert np.abs(fake.detach().numpy() - GOLD["fake"]).max() < 5e-5
    fake.backward(torch.from_numpy(GOLD["upstream"]))
    for k in ("up1_up.weight", "up2_up.weight", "up1_up.bias", "up2_up.bias", "outc.1.weight", "up1_conv.0.weight"):
        assert abs(leaves[k].grad.norm().item() - float(GOLD["grad_norm/" + k])) < 2e-3 * float(GOLD["grad_norm/" + k]), k


# variant tag -> (fixture prefix, no_antialias_up, no_antialias, parameter seed of make_golden_variants.py)
VARIANTS = {"up": ("", True, False, 4321), "na": ("na/", False, True, 777), "nab": ("nab/", True, True, 778), "odd": ("odd/", False, False, 999),
            "nn": ("nn/", False, False, 555), "bn": ("bn/", False, False, 333)}
SIZES = {"odd": (29, 38)}          # default graph at a size that is not a multiple of 4: bilinear fix-up of irc:555-556, :562-563
NORM = {"nn": "none", "bn": "batch"}   # norm='none': Identity layers, bias-free convolutions; 'batch': nn.BatchNorm2d (irc:158-163)
BIAS_STD = {"bn": 0.05}


def _variant_params(tag):
    pre, up, na, seed = VARIANTS[tag]
    return O.seeded_params(O.generator_shapes(no_antialias_up=up, norm=NORM.get(tag, "instance")), seed, bias_std=BIAS_STD.get(tag, 0.02))


@pytest.mark.parametrize("tag", ["na", "nab"])
def test_oracle_no_antialias_forward_and_grads(tag):
    pre, up, na, _ = VARIANTS[tag]
    ir, _ = O.synthetic_pair(B, H, W)
    leaves = {k: v.clone().requires_grad_(True) for k, v in _variant_params(tag).items()}
    fake = O.generator_forward(leaves, ir, no_antialias=True)
    assert np.abs(fake.detach().numpy() - GOLD[pre + "fake"]).max() < 5e-5
    fake.backward(torch.from_numpy(GOLD[pre + "upstream"]))
    for k in ("down1.0.weight", "down2.0.weight", "inc.1.weight", "outc.1.weight", "up1_conv.0.weight"):
        assert abs(leaves[k].grad.norm().item() - float(GOLD[pre + "grad_norm/" + k])) < 2e-3 * float(GOLD[pre + "grad_norm/" + k]), k


def _check_engine(be, dev, tol_fwd, tol_dec, tol_enc, tag="up"):
    import irc_b200  # noqa: F401
    from irc_b200 import engine as E
    pre, up, na, _ = VARIANTS[tag]
    pG = _variant_params(tag)
    h, w = SIZES.get(tag, (H, W))
    ir, _ = O.synthetic_pair(B, h, w)
    eng = E.GeneratorEngine(be, B, h, w, dev, no_antialias_up=up, no_antialias=na, norm=NORM.get(tag, "instance"))
    eng.arena.load(pG); eng.refresh_weights()
    fake = eng.forward(ir.to(dev))
    e = rel(fake, torch.from_numpy(GOLD[pre + "fake"]))
    print("variant", tag, "fake rel", e)
    assert e < tol_fwd
    eng.arena.grad.zero_()
    eng.backward(torch.from_numpy(GOLD[pre + "upstream"]).to(dev).contiguous())
    worst = {}
    for k in pG:
        conv_bias_in_front_of_instance_norm = tag not in NORM and k.endswith("bias") and not k.startswith(("outc", "up1_up", "up2_up"))
        if float(GOLD[pre + "grad_absmax/" + k]) <= 1e-4 or conv_bias_in_front_of_instance_norm:
            continue
        got = sample(eng.arena.view(k, eng.arena.grad)); want = GOLD[pre + "grad_sample/" + k]
        worst[k] = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    print({k: round(v, 4) for k, v in worst.items() if not k.startswith("resblocks") or k.startswith("resblocks.8")})
    for k, v in worst.items():
        decoder = k.startswith(("outc", "up2", "up1"))
        assert v < (tol_dec if decoder else tol_enc), (k, v)


def test_plan_with_transposed_conv_upsampling_matches_reference():
    """host logic of the variant (weight layout of the 4-phase GEMM, depth-to-space reads, zero ring, gradient routing) in
    float32; the encoder-side gradients pass through 18 ReLU masks evaluated in a different summation order (2e-3 fp32 noise)"""
    import irc_b200  # noqa: F401
    from irc_b200 import layout as L
    from ref_backend import RefBackend
    old = L.ACT_DTYPE
    L.ACT_DTYPE = torch.float32
    try:
        _check_engine(RefBackend(), "cpu", 1e-5, 1e-4, 1e-2)
    finally:
        L.ACT_DTYPE = old


@pytest.mark.parametrize("tag", ["na", "nab"])
def test_plan_with_strided_downsampling_matches_reference(tag):
    """host logic of no_antialias=True in float32: 3x3 stride-2 convolutions as 2x2 convolutions over space-to-depth blocks (weight
    layout with structural zeros), InstanceNorm on the block grid, gradients returning in space-to-depth order into the two-source
    InstanceNorm backward of the layer below"""
    import irc_b200  # noqa: F401
    from irc_b200 import layout as L
    from ref_backend import RefBackend
    old = L.ACT_DTYPE
    L.ACT_DTYPE = torch.float32
    try:
        _check_engine(RefBackend(), "cpu", 1e-5, 1e-4, 1e-2, tag=tag)
    finally:
        L.ACT_DTYPE = old


def test_oracle_odd_size_forward():
    ir, _ = O.synthetic_pair(B, *SIZES["odd"])
    fake = O.generator_forward(_variant_params("odd"), ir)
    assert np.abs(fake.numpy() - GOLD["odd/fake"]).max() < 5e-5


def test_plan_at_a_size_that_is_not_a_multiple_of_4_matches_reference():
    """29 x 38: Downsample rounds up, the decoder maps are resized onto the skip-connection grids (irc:555-556, :562-563); the
    resize is composed with UpsampleAA into one separable operator per axis (layout.resize_matrix @ layout.up_matrix)"""
    import irc_b200  # noqa: F401
    from irc_b200 import layout as L
    from ref_backend import RefBackend
    old = L.ACT_DTYPE
    L.ACT_DTYPE = torch.float32
    try:
        _check_engine(RefBackend(), "cpu", 1e-5, 1e-4, 1e-2, tag="odd")
    finally:
        L.ACT_DTYPE = old


def test_oracle_norm_none_forward():
    ir, _ = O.synthetic_pair(B, H, W)
    fake = O.generator_forward(_variant_params("nn"), ir)
    assert np.abs(fake.numpy() - GOLD["nn/fake"]).max() < 5e-5


def test_plan_without_normalisation_matches_reference():
    """norm='none' in float32: no statistics, ReLU in the GEMM epilogues, plain copies into the next frame (ring / stencil /
    residual), mask-only backward passes; every parameter gradient against the reference (no InstanceNorm -> no ReLU-mask noise
    amplification, hence the tight encoder bound)"""
    import irc_b200  # noqa: F401
    from irc_b200 import layout as L
    from ref_backend import RefBackend
    old = L.ACT_DTYPE
    L.ACT_DTYPE = torch.float32
    try:
        _check_engine(RefBackend(), "cpu", 1e-5, 1e-4, 1e-3, tag="nn")
    finally:
        L.ACT_DTYPE = old


def _check_discriminator_none(be, dev, tol_fwd, tol_grad, norm="none", pre="nnD/", seed=556, bias_std=0.02):
    import irc_b200  # noqa: F401
    from irc_b200 import engine as E
    pD = O.seeded_params(O.discriminator_shapes(norm=norm), seed, bias_std=bias_std)
    ir, rgb = O.synthetic_pair(B, H, W)
    eng = E.DiscriminatorEngine(be, B, H, W, dev, norm=norm)
    eng.arena.load(pD); eng.refresh_weights()
    pred = eng.forward(ir.to(dev).contiguous(), rgb.to(dev).contiguous())
    assert rel(pred, torch.from_numpy(GOLD[pre + "pred"])) < tol_fwd
    eng.arena.grad.zero_()
    dx = torch.zeros(B, 4, H, W, device=dev)
    eng.backward(torch.from_numpy(GOLD[pre + "upstream"]).to(dev).contiguous(), True, dx, c_first=0, accumulate=False)
    assert rel(dx, torch.from_numpy(GOLD[pre + "dx"])) < tol_grad
    for k in pD:
        got = sample(eng.arena.view(k, eng.arena.grad), 512); want = GOLD[pre + "grad/" + k]
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < tol_grad, k
    return eng


def test_discriminator_plan_without_normalisation_matches_reference():
    import irc_b200  # noqa: F401
    from irc_b200 import layout as L
    from ref_backend import RefBackend
    old = L.ACT_DTYPE
    L.ACT_DTYPE = torch.float32
    try:
        _check_discriminator_none(RefBackend(), "cpu", 1e-5, 1e-4)
    finally:
        L.ACT_DTYPE = old


def test_plan_with_batch_norm_matches_reference():
    """norm='batch' in float32: batch statistics folded with the affine parameters into effective moments (irc_bn_finalize), the
    InstanceNorm kernels doing the rest, parameter gradients and the input-gradient correction of irc_bn_bwd_fix, running statistics;
    then eval(): running statistics instead of batch statistics"""
    import irc_b200  # noqa: F401
    from irc_b200 import engine as E, layout as L
    from ref_backend import RefBackend
    old = L.ACT_DTYPE
    L.ACT_DTYPE = torch.float32
    try:
        _check_engine(RefBackend(), "cpu", 1e-5, 2e-4, 2e-3, tag="bn")
        # running statistics after one training forward, and the eval-mode output
        eng = E.GeneratorEngine(RefBackend(), B, H, W, "cpu", norm="batch", training=False)
        eng.arena.load(_variant_params("bn")); eng.refresh_weights()
        ir, _ = O.synthetic_pair(B, H, W)
        eng.forward(ir)
        for k in ("inc.2", "down2.1", "resblocks.4.conv_block.6", "up2_conv.1"):
            assert np.abs(eng.bn_state[k][0].numpy() - GOLD["bn/rm/" + k]).max() < 1e-5 and np.abs(eng.bn_state[k][1].numpy() - GOLD["bn/rv/" + k]).max() < 1e-5, k
        eng.bn_training = False
        assert rel(eng.forward(ir), torch.from_numpy(GOLD["bn/fake_eval"])) < 1e-5
        d = _check_discriminator_none(RefBackend(), "cpu", 1e-5, 2e-4, norm="batch", pre="bnD/", seed=334, bias_std=0.05)
        assert np.abs(d.bn_state["model.6"][0].numpy() - GOLD["bnD/rm/model.6"]).max() < 1e-5
        assert np.abs(d.bn_state["model.6"][1].numpy() - GOLD["bnD/rv/model.6"]).max() < 1e-5
    finally:
        L.ACT_DTYPE = old


def test_full_train_step_with_batch_norm_matches_reference():
    """one fused D + G iteration with norm='batch' against the reference's own iteration (irc:1636-1681, fixture bnstep/*): the
    discriminator normalises the real and the fake half of its batched pass separately (two calls in the reference), the single
    generator forward stands for the reference's two (two running-statistics updates), three discriminator updates; losses,
    post-Adam parameters (BatchNorm affine parameters included) and running statistics"""
    import irc_b200  # noqa: F401
    from irc_b200 import layout as L
    from irc_b200.train_step import TrainStep
    from ref_backend import RefBackend
    old = L.ACT_DTYPE
    L.ACT_DTYPE = torch.float32
    try:
        pG = _variant_params("bn")
        pD = O.seeded_params(O.discriminator_shapes(norm="batch"), 334, bias_std=0.05)
        pV = O.seeded_params(O.vgg_shapes(), 13, kaiming=True, bias_std=0.05)
        ir, rgb = O.synthetic_pair(B, H, W)
        ts = TrainStep(RefBackend(), B, H, W, "cpu", norm="batch")
        ts.load(pG, pD, pV)
        ts.step(ir, rgb)
        los = ts.losses()
        for k in ("D", "G", "GAN", "L1", "perc", "SSIM"):
            ref = float(GOLD["bnstep/loss_" + k])
            assert abs(los[k] - ref) < 5e-5 * max(1.0, abs(ref)), (k, los[k], ref)
        for k in ("model.3", "model.6", "model.9"):
            assert np.abs(ts.D2.bn_state[k][0].numpy() - GOLD["bnstep/rmD/" + k]).max() < 2e-5, k
            assert np.abs(ts.D2.bn_state[k][1].numpy() - GOLD["bnstep/rvD/" + k]).max() < 2e-5, k
        for k in ("inc.2", "resblocks.0.conv_block.2", "up1_conv.1"):
            assert np.abs(ts.G.bn_state[k][0].numpy() - GOLD["bnstep/rmG/" + k]).max() < 2e-5, k
            assert np.abs(ts.G.bn_state[k][1].numpy() - GOLD["bnstep/rvG/" + k]).max() < 2e-5, k
        # first Adam step moves every element by +-lr (sign of the gradient): elements whose gradient is ~eps are rounding-sensitive
        for k in pD:
            got = sample(ts.D2.arena.view(k)); want = GOLD["bnstep/pD_after/" + k]
            assert (np.abs(got - want) > 5e-5).mean() < 0.02, k
        for k in ("outc.1.weight", "up2_conv.1.weight", "up2_conv.1.bias", "inc.2.weight"):
            got = sample(ts.G.arena.view(k)); want = GOLD["bnstep/pG_after/" + k]
            assert (np.abs(got - want) > 5e-5).mean() < 0.02, k
    finally:
        L.ACT_DTYPE = old


def test_module_surface_with_batch_norm():
    """get_norm_layer('batch') (irc:158-159): nn.BatchNorm2d keys in the state_dict (affine parameters + running statistics +
    num_batches_tracked), train() / eval() semantics through the module, N(1, gain) initialisation of the norm weights (irc:190-194)"""
    import irc_b200 as R
    from irc_b200 import layout as L, modules as M
    from ref_backend import RefBackend
    old, oldbe = L.ACT_DTYPE, M._BACKEND
    L.ACT_DTYPE = torch.float32; M.set_backend(RefBackend())
    try:
        cfg = R.Config(); cfg.device = "cpu"; cfg.norm = "batch"
        m = R.IRColorizationModel(cfg)
        sd = m.netG.state_dict()
        assert {"inc.2.weight", "inc.2.running_var", "down1.1.num_batches_tracked", "resblocks.8.conv_block.6.bias", "up2_conv.1.running_mean"} <= set(sd)
        assert "inc.1.bias" not in sd and abs(sd["down2.1.weight"].mean().item() - 1.0) < 0.02 and sd["inc.2.bias"].abs().max() == 0
        m.netG.load_state_dict(_variant_params("bn"), strict=False)
        ir, _ = O.synthetic_pair(B, H, W)
        m.train()
        with torch.no_grad():
            assert rel(m(ir), torch.from_numpy(GOLD["bn/fake"])) < 1e-5
        assert int(m.netG.state_dict()["inc.2.num_batches_tracked"]) == 1
        assert np.abs(m.netG.state_dict()["inc.2.running_mean"].numpy() - GOLD["bn/rm/inc.2"]).max() < 1e-5
        m.eval()
        with torch.no_grad():
            assert rel(m(ir), torch.from_numpy(GOLD["bn/fake_eval"])) < 1e-5
        netD = R.NLayerDiscriminator(4, 64, 3, R.get_norm_layer("batch"))
        assert {"model.3.weight", "model.6.running_mean", "model.9.num_batches_tracked"} <= set(netD.state_dict()) and "model.2.bias" not in netD.state_dict()
    finally:
        L.ACT_DTYPE = old; M.set_backend(oldbe)


def test_module_surface_without_normalisation():
    """get_norm_layer('none') (irc:154-165): Identity factory; the generator / discriminator state_dicts lose the biases of the
    convolutions that sit in front of a norm layer (use_bias False, irc:452-455, :590-593); 'batch' says that it is not built"""
    import irc_b200 as R
    from irc_b200 import layout as L, modules as M
    from ref_backend import RefBackend
    old, oldbe = L.ACT_DTYPE, M._BACKEND
    L.ACT_DTYPE = torch.float32; M.set_backend(RefBackend())
    try:
        assert isinstance(R.get_norm_layer("none")(64), R.Identity) and isinstance(R.get_norm_layer(None)(8), R.Identity)
        with pytest.raises(NotImplementedError):
            R.get_norm_layer("layer")
        cfg = R.Config(); cfg.device = "cpu"; cfg.norm = "none"
        m = R.IRColorizationModel(cfg)
        sd = m.netG.state_dict()
        assert set(k for k in sd if not k.endswith("filt")) == set(O.generator_shapes(norm="none")) and "outc.1.bias" in sd
        m.netG.load_state_dict(_variant_params("nn"), strict=False)
        ir, _ = O.synthetic_pair(B, H, W)
        with torch.no_grad():
            assert rel(m(ir), torch.from_numpy(GOLD["nn/fake"])) < 1e-5
        netD = R.NLayerDiscriminator(4, 64, 3, R.get_norm_layer("none"))
        assert set(netD.state_dict()) == set(O.discriminator_shapes(norm="none"))
    finally:
        L.ACT_DTYPE = old; M.set_backend(oldbe)


def test_module_surface_of_the_strided_variant():
    """no_antialias=True: the reference's state_dict has no down{1,2}_down.filt buffers (the modules are None, irc:474, :482)"""
    import irc_b200 as R
    from irc_b200 import layout as L, modules as M
    from ref_backend import RefBackend
    old, oldbe = L.ACT_DTYPE, M._BACKEND
    L.ACT_DTYPE = torch.float32; M.set_backend(RefBackend())
    try:
        cfg = R.Config(); cfg.device = "cpu"; cfg.no_antialias = True
        m = R.IRColorizationModel(cfg)
        sd = m.netG.state_dict()
        assert "down1_down.filt" not in sd and "down2_down.filt" not in sd and "up1_up.filt" in sd and len(sd) == 50
        assert m.netG.down1_down is None and m.netG.down2_down is None
        m.netG.load_state_dict(_variant_params("na"), strict=False)
        ir, _ = O.synthetic_pair(B, H, W)
        with torch.no_grad():
            assert rel(m(ir), torch.from_numpy(GOLD["na/fake"])) < 1e-5
    finally:
        L.ACT_DTYPE = old; M.set_backend(oldbe)


def test_module_surface_of_the_variant():
    """state_dict keys of the reference with no_antialias_up=True: ConvTranspose2d parameters, no up*_up.filt buffers"""
    import irc_b200 as R
    from irc_b200 import layout as L, modules as M
    from ref_backend import RefBackend
    old, oldbe = L.ACT_DTYPE, M._BACKEND
    L.ACT_DTYPE = torch.float32; M.set_backend(RefBackend())
    try:
        cfg = R.Config(); cfg.device = "cpu"; cfg.no_antialias_up = True
        m = R.IRColorizationModel(cfg)
        sd = m.netG.state_dict()
        assert tuple(sd["up1_up.weight"].shape) == (256, 256, 3, 3) and tuple(sd["up2_up.bias"].shape) == (128,)
        assert "up1_up.filt" not in sd and "down1_down.filt" in sd
        m.netG.load_state_dict(_params(), strict=False)
        ir, _ = O.synthetic_pair(B, H, W)
        with torch.no_grad():
            assert rel(m(ir), torch.from_numpy(GOLD["fake"])) < 1e-5
    finally:
        L.ACT_DTYPE = old; M.set_backend(oldbe)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_generator_with_transposed_conv_upsampling():
    """whole-network bf16 run: forward 1e-2; gradients inherit the ReLU-mask flips of the bf16 forward exactly like the default
    graph (tests/test_step_gpu.py bounds them at 0.35); the tight 1e-2 bound of the new layer is the C-ABI test below"""
    from irc_b200._native import CudaBackend
    _check_engine(CudaBackend(), "cuda", 3e-2, 0.25, 0.35)


@pytest.mark.gpu
@pytest.mark.gpu
def test_gpu_full_train_step_with_batch_norm():
    """norm='batch' through the fused CUDA iteration (bf16): losses and running statistics against the reference's iteration, and
    CUDA-graph replay == eager, bit for bit, over two iterations (the BatchNorm helper launches are captured like everything else)"""
    from irc_b200._native import CudaBackend
    from irc_b200.train_step import TrainStep
    pG = _variant_params("bn")
    pD = O.seeded_params(O.discriminator_shapes(norm="batch"), 334, bias_std=0.05)
    pV = O.seeded_params(O.vgg_shapes(), 13, kaiming=True, bias_std=0.05)
    ir, rgb = O.synthetic_pair(B, H, W)
    res = []
    for graph in (False, True):
        ts = TrainStep(CudaBackend(), B, H, W, "cuda", norm="batch", use_graph=graph)
        ts.load(pG, pD, pV)
        ts.step(ir.cuda(), rgb.cuda())
        if not graph:
            los = ts.losses()
            for k, tol in (("D", 2e-2), ("G", 1e-2), ("L1", 1e-2), ("perc", 3e-2), ("SSIM", 1e-2)):
                ref = float(GOLD["bnstep/loss_" + k])
                assert abs(los[k] - ref) < tol * max(1.0, abs(ref)), (k, los[k], ref)
            for k in ("model.3", "model.6", "model.9"):
                assert rel(ts.D2.bn_state[k][1], torch.from_numpy(GOLD["bnstep/rvD/" + k])) < 3e-2, k
            for k in ("inc.2", "resblocks.0.conv_block.2", "up1_conv.1"):
                assert rel(ts.G.bn_state[k][0], torch.from_numpy(GOLD["bnstep/rmG/" + k])) < 5e-2, k
                assert rel(ts.G.bn_state[k][1], torch.from_numpy(GOLD["bnstep/rvG/" + k])) < 3e-2, k
        ts.step(ir.cuda(), rgb.cuda())
        torch.cuda.synchronize()
        res.append((ts.G.arena.flat.clone(), ts.D2.arena.flat.clone(), ts.G.bn_state["up2_conv.1"][1].clone(), ts.D2.bn_state["model.9"][0].clone()))
    for a, b in zip(*res):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_gpu_networks_without_normalisation():
    """norm='none' on the CUDA path (bf16): without InstanceNorm the forward error stays at the bf16 level and the gradients do not
    see amplified ReLU-mask flips"""
    from irc_b200._native import CudaBackend
    _check_engine(CudaBackend(), "cuda", 2e-2, 0.15, 0.2, tag="nn")
    _check_engine(CudaBackend(), "cuda", 3e-2, 0.25, 0.35, tag="bn")          # nn.BatchNorm2d: same whole-network bounds as InstanceNorm
    _check_discriminator_none(CudaBackend(), "cuda", 3e-2, 0.2, norm="batch", pre="bnD/", seed=334, bias_std=0.05)
    _check_discriminator_none(CudaBackend(), "cuda", 2e-2, 0.12)      # four LeakyReLU masks evaluated on bf16 pre-activations


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["na", "nab", "odd"])
def test_gpu_generator_with_strided_downsampling(tag):
    """no_antialias=True, and the default graph at 29 x 38 (odd-size bilinear fix-up), on the CUDA path (bf16): same bounds as the
    other whole-network runs"""
    from irc_b200._native import CudaBackend
    _check_engine(CudaBackend(), "cuda", 3e-2, 0.25, 0.35, tag=tag)


@pytest.mark.gpu
def test_gpu_convT2d_fwd_through_the_c_abi():
    """irc_convT2d_fwd + the data / weight gradients through irc_conv_gemm / irc_tn_gemm against nn.ConvTranspose2d (reference
    run: ct_* fixtures), on a zero-ringed frame"""
    import irc_b200  # noqa: F401
    from irc_b200 import _native as nat, layout as L, engine as E
    from irc_b200._native import CudaBackend, View
    be = CudaBackend()
    x = torch.from_numpy(GOLD["ct_x"]); w = torch.from_numpy(GOLD["ct_w"]); b = torch.from_numpy(GOLD["ct_b"])
    n, c, h, wd = x.shape
    arena = L.ParamArena({"up.weight": (c, c, 3, 3), "up.bias": (c,)}, "cuda")
    arena.load({"up.weight": w, "up.bias": b})
    packer = L.Packer(arena)
    src = L.Frame(n, h, wd, 1, c, "cuda")
    up = E.TransposedUp(be, packer, arena, "up", c, src, "cuda", "ct")
    packer.finish(); packer.refresh(be); up.refresh()
    src.t.view(n, src.hp, src.wp, c)[:, 1:-1, 1:-1].copy_(x.permute(0, 2, 3, 1))
    # forward through the C entry point itself
    out = torch.zeros(src.rows, 4 * c, device="cuda", dtype=torch.bfloat16)
    nat.check(nat.lib().irc_convT2d_fwd(C.c_void_p(src.t.data_ptr()), C.c_longlong(src.rows), c, 0, c, src.wp, C.c_void_p(up.op.lay.w_f.t.data_ptr()), c,
                                        C.c_void_p(up.bias4.data_ptr()), C.c_void_p(out.data_ptr()), C.c_longlong(4 * c), 0,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    dst = L.Frame(n, 2 * h, 2 * wd, 1, c, "cuda")
    be.gather(View(out, 0, src.hp, src.wp, 2, 2, c), dst.view(0), c, n, 2 * h, 2 * wd, 1, 0)
    y = dst.t.view(n, dst.hp, dst.wp, c)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().cpu()
    e = rel(y, torch.from_numpy(GOLD["ct_y"]))
    print("convT fwd rel", e)
    assert e < 1e-2
    # the engine path gives the same bits
    up.forward(dst, 0)
    assert torch.equal(dst.t.view(n, dst.hp, dst.wp, c)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().cpu(), y)
    # backward
    g = L.Frame(n, 2 * h, 2 * wd, 1, c, "cuda")
    g.t.view(n, g.hp, g.wp, c)[:, 1:-1, 1:-1].copy_(torch.from_numpy(GOLD["ct_gy"]).permute(0, 2, 3, 1))
    dx = torch.zeros(src.rows, c, device="cuda", dtype=torch.bfloat16)
    arena.grad.zero_()
    up.backward(g, 0, dx); be.flush_sums()
    gx = dx.view(n, src.hp, src.wp, c)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
    assert rel(gx, torch.from_numpy(GOLD["ct_gx"])) < 1e-2
    assert rel(arena.view("up.weight", arena.grad), torch.from_numpy(GOLD["ct_gw"])) < 1e-2
    assert rel(arena.view("up.bias", arena.grad), torch.from_numpy(GOLD["ct_gb"])) < 1e-2
    ring = dx.view(n, src.hp, src.wp, c)
    assert ring[:, 0].abs().max() == 0 and ring[:, :, -1].abs().max() == 0
