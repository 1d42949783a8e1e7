"""BASELINE.json configs at their own shapes: the non-square KAIST resolution 512x640 (config 4) and batched
test-mode inference + metrics (config 5).  Small non-square cases are compared with the oracle run on the spot; the
full-size cases through size-independent properties (the losses the step reports == the reference's formulas evaluated on
the step's own output, determinism, sample independence) because the CPU oracle takes minutes at that size."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm()).item()


def _params():
    import irc_oracle as O
    return (O.seeded_params(O.generator_shapes(), 1234, bias_std=0.02), O.seeded_params(O.discriminator_shapes(), 1235, bias_std=0.02),
            O.seeded_params(O.vgg_shapes(), 1236, kaiming=True, bias_std=0.05))


def test_nonsquare_step_vs_oracle():
    """H != W exercises every row/column stride of the frame layout (48 x 80: the 512 x 640 aspect ratio)"""
    import irc_oracle as O
    from irc_b200._native import CudaBackend
    from irc_b200.train_step import TrainStep
    B, H, W = 2, 48, 80
    pG, pD, pV = _params()
    ir, rgb = O.synthetic_pair(B, H, W)
    ts = TrainStep(CudaBackend(), B, H, W, "cuda")
    ts.load(pG, pD, pV)
    ts.step(ir.cuda(), rgb.cuda())
    torch.cuda.synchronize()
    oG = {k: v.clone() for k, v in pG.items()}; oD = {k: v.clone() for k, v in pD.items()}
    fake_ref = O.generator_forward(oG, ir)
    assert rel(ts.G.fake, fake_ref) < 4e-2                      # bf16 operands, 24 chained convs
    losses, gG, gD = O.train_step(oG, oD, pV, O.AdamState(oG), O.AdamState(oD), ir, rgb)
    los = ts.losses()
    for k, tol in (("D", 1e-2), ("G", 3e-3), ("L1", 2e-3), ("perc", 1e-2), ("TV", 2e-2), ("SSIM", 3e-3)):
        assert abs(los[k] - float(losses[k])) <= tol * max(1.0, abs(float(losses[k]))), (k, los[k], float(losses[k]))
    gg = ts.G.arena.grads()
    assert rel(gg["outc.1.weight"], gG["outc.1.weight"]) < 0.1 and rel(gg["outc.1.bias"], gG["outc.1.bias"]) < 0.05


@pytest.fixture(scope="module")
def full_step():
    import irc_oracle as O
    from irc_b200._native import CudaBackend
    from irc_b200.train_step import TrainStep
    B, H, W = 2, 512, 640
    pG, pD, pV = _params()
    ir, rgb = O.synthetic_pair(B, H, W)
    ts = TrainStep(CudaBackend(), B, H, W, "cuda", use_graph=True)
    ts.load(pG, pD, pV)
    ts.step(ir.cuda(), rgb.cuda())
    torch.cuda.synchronize()
    return dict(ts=ts, ir=ir, rgb=rgb, p=(pG, pD, pV))


def test_full_resolution_step_losses_are_the_reference_formulas(full_step):
    """config 4: one D+G iteration at 512 x 640 with anti-aliased Downsample / UpsampleAA.  The pixel-space losses the step
    reports must equal the reference's formulas (oracle, CPU fp32) evaluated on the step's own `fake` (fp32 kernels: 1e-4)."""
    import irc_oracle as O
    ts, rgb = full_step["ts"], full_step["rgb"]
    los = ts.losses()
    assert all(np.isfinite(v) for v in los.values()), los
    fake = ts.G.fake.detach().cpu()
    assert fake.shape == (2, 3, 512, 640) and fake.abs().max() <= 1.0
    l1 = (fake - rgb).abs().mean().item()
    tv = O.tv_loss(fake).item()
    ssim = O.ssim_loss((fake + 1) / 2, (rgb + 1) / 2).item()
    lam = O.LAMBDAS
    assert abs(los["L1"] - lam["L1"] * l1) < 1e-4 * lam["L1"] * l1
    assert abs(los["TV"] - lam["tv"] * tv) < 1e-4 * lam["tv"] * tv
    assert abs(los["SSIM"] - lam["ssim"] * ssim) < 2e-4 * max(1.0, lam["ssim"] * ssim)
    # PatchGAN output geometry at this resolution (62 x 78)
    assert tuple(ts.D1.pred.shape[-2:]) == (62, 78)
    for arena in (ts.G.arena, ts.D2.arena):
        assert torch.isfinite(arena.flat).all() and torch.isfinite(arena.grad).all()


def test_full_resolution_generator_vs_oracle_one_image(full_step):
    """generator forward at 512 x 640 against the oracle (one image: ~10 s of CPU)"""
    import irc_oracle as O
    from irc_b200._native import CudaBackend
    from irc_b200.engine import GeneratorEngine
    pG = full_step["p"][0]
    ir = full_step["ir"][:1].contiguous()
    eng = GeneratorEngine(CudaBackend(), 1, 512, 640, "cuda", training=False)
    eng.arena.load(pG); eng.refresh_weights()
    fake = eng.forward(ir.cuda())
    with torch.no_grad():
        ref = O.generator_forward(pG, ir)
    assert rel(fake, ref) < 4e-2


def test_full_resolution_step_is_deterministic_and_sample_independent(full_step):
    """same state + same batch -> same bits; and InstanceNorm makes samples independent: the generator output of sample 0
    does not depend on what sample 1 is (the property data parallelism over the batch rests on)"""
    import irc_oracle as O
    from irc_b200._native import CudaBackend
    from irc_b200.engine import GeneratorEngine
    pG = full_step["p"][0]
    ir = full_step["ir"]
    eng = GeneratorEngine(CudaBackend(), 2, 512, 640, "cuda", training=False)
    eng.arena.load(pG); eng.refresh_weights()
    a = eng.forward(ir.cuda()).clone()
    b = eng.forward(ir.cuda()).clone()
    assert torch.equal(a, b)
    ir2 = ir.clone(); ir2[1] = -ir2[1]
    c = eng.forward(ir2.cuda())
    assert torch.equal(a[0], c[0]) and not torch.equal(a[1], c[1])


def test_batched_inference_metrics_512x640():
    """config 5 per GPU: 8 images of 512 x 640, generator forward + truncating uint8 quantisation (irc:865-876) + per-image
    MAE / MSE / PSNR (irc:1197-1205) in one device pass == the reference's numpy formulas on the same prediction."""
    import irc_b200 as R
    from irc_b200.train import batch_metrics
    B, H, W = 8, 512, 640
    cfg = R.Config(); cfg.device = "cuda"
    model = R.IRColorizationModel(cfg).eval()
    g = torch.Generator().manual_seed(5)
    ir = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    gt = torch.rand(B, 3, H, W, generator=g)
    with torch.no_grad():
        fake = model(ir.cuda())
    u8, mae, mse, psnr = batch_metrics(fake, gt.cuda())
    f = fake.float().cpu().numpy()
    ref_u8 = (np.clip((f + 1.0) / 2.0, 0.0, 1.0) * 255.0).astype(np.uint8).transpose(0, 2, 3, 1)      # irc:869-875 (truncation)
    assert np.array_equal(u8.cpu().numpy(), ref_u8)
    for i in range(B):
        m = R.compute_metrics(ref_u8[i].astype(np.float32) / 255.0, gt[i].permute(1, 2, 0).numpy())
        assert abs(m[0] - mae[i]) < 1e-6 and abs(m[1] - mse[i]) < 1e-6 and abs(m[2] - psnr[i]) < 1e-4, (i, m, mae[i], mse[i], psnr[i])
