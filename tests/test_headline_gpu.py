"""The CUDA path at the HEADLINE shape (256 x 256; BASELINE.json configs[0] and [1]) against golden vectors produced by
the UNMODIFIED reference (tests/golden/ref_256.npz, oracle/make_golden_256.py): per-layer taps of the generator, the
output image, the test-mode MAE / MSE / PSNR, and one full D+G iteration at B = 2 (losses, dL/dfake, every parameter
gradient, post-Adam parameters).

Stated tolerances (bf16 operands, fp32 accumulation):
  per-layer taps      rel-L2 of the 256-point sample grows with depth: 1e-2 after the first conv .. 3e-2 after 20
  output image        rel-L2 <= 3e-2 over all 196 608 values
  test-mode metrics   MAE and MSE to 3 decimals (|d| < 5e-4), PSNR to 2 decimals (|d| < 5e-3 dB)
  losses              D 1e-2, G / L1 / SSIM 3e-3, perc 1e-2 relative
  gradients           whole-step gradients inherit ReLU-mask / sign() flips from the bf16 forward (the reference's own fp32
                      restatement already differs from it by 3e-3 at this size, fixture key
                      fp32_oracle_vs_reference_grad_rel_max): cosine similarity >= 0.95 and norm ratio within 10 % per tensor;
                      the tight per-kernel bound (1e-2, teacher-forced) is tests/test_teacher_forced_gpu.py"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_256.npz"))
H = W = 256


def sample(t, n=256):
    f = t.detach().float().cpu().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].numpy()


def rel_np(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def interior(fr, c, off=0):
    v = fr.t.view(fr.N, fr.hp, fr.wp, fr.C)[:, fr.p:fr.p + fr.H, fr.p:fr.p + fr.W, off:off + c]
    return v.permute(0, 3, 1, 2).float().contiguous()


def params():
    import irc_oracle as O
    return (O.seeded_params(O.generator_shapes(), 1234, bias_std=0.02), O.seeded_params(O.discriminator_shapes(), 1235, bias_std=0.02),
            O.seeded_params(O.vgg_shapes(), 1236, kaiming=True, bias_std=0.05))


def test_generator_256_per_layer_taps_and_output_vs_reference():
    """config 1 on the GPU: B = 1, eval-mode generator forward"""
    import irc_oracle as O
    from irc_b200 import engine as E
    from irc_b200._native import CudaBackend
    pG, _, _ = params()
    ir, _ = O.synthetic_pair(1, H, W)
    eng = E.GeneratorEngine(CudaBackend(), 1, H, W, "cuda", training=False)
    eng.arena.load(pG); eng.refresh_weights()
    fake = eng.forward(ir.cuda())
    torch.cuda.synchronize()
    taps = {"x0": (interior(eng.cat2, 64, 128), 1e-2), "x1": (interior(eng.cat1, 128, 256), 1e-2), "x2": (interior(eng.X[0], 256), 1.2e-2),
            "up1_up": (interior(eng.cat1, 256, 0), 3e-2), "up2_up": (interior(eng.cat2, 128, 0), 3e-2), "up2": (interior(eng.y4, 64), 3e-2)}
    for b in range(9):
        taps[f"res{b}"] = (interior(eng.X[b + 1], 256), 1.2e-2 + 2e-3 * (b + 1))
    worst = {}
    for k, (t, tol) in taps.items():
        e = rel_np(sample(t), GOLD[f"G1_{k}_sample"])
        am = abs(t.abs().mean().item() - float(GOLD[f"G1_{k}_absmean"])) / float(GOLD[f"G1_{k}_absmean"])
        nr = abs(t.norm().item() - float(GOLD[f"G1_{k}_norm"])) / float(GOLD[f"G1_{k}_norm"])
        worst[k] = round(e, 4)
        assert e < tol, (k, e, tol)
        assert am < 5e-3 and nr < 5e-3, (k, am, nr)
    print("per-layer rel-L2 (256-point samples):", worst)
    e = rel_np(fake.cpu().numpy(), GOLD["fake_b1"])
    print("fake rel-L2", e)
    assert e < 3e-2


def test_test_mode_metrics_agree_with_the_reference_to_3_decimals():
    """north star: test-mode MAE / MSE / PSNR of the GPU generator vs the reference generator (irc:1381-1389, :865-876,
    :1184-1205) on the same input and ground truth, through the module surface"""
    import irc_oracle as O
    import irc_b200 as R
    from irc_b200.train import batch_metrics
    pG, _, _ = params()
    cfg = R.Config(); cfg.device = "cuda"
    model = R.IRColorizationModel(cfg)
    model.netG.load_state_dict(pG, strict=False)
    model.eval()
    ir, _ = O.synthetic_pair(1, H, W)
    gt = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(99))
    with torch.no_grad():
        fake = model(ir.cuda())
    u8, mae, mse, psnr = batch_metrics(fake, gt.permute(2, 0, 1)[None].cuda())
    want = GOLD["metrics_b1"]
    print("metrics", (mae[0], mse[0], psnr[0]), "reference", tuple(want))
    assert abs(mae[0] - want[0]) < 5e-4 and abs(mse[0] - want[1]) < 5e-4
    assert abs(psnr[0] - want[2]) < 5e-3
    # quantised bytes: the reference truncates (irc:874); a bf16 forward error of ~2 % of the range moves a byte by a few counts
    q = sample(u8[0].int(), 1024)
    d = np.abs(q.astype(np.int32) - GOLD["quant_b1_sample"].astype(np.int32))
    print("uint8 |diff| mean", d.mean(), "max", d.max())
    assert d.mean() < 2.0 and d.max() <= 24


@pytest.fixture(scope="module")
def step():
    import irc_oracle as O
    from irc_b200._native import CudaBackend
    from irc_b200.train_step import TrainStep
    pG, pD, pV = params()
    ir, rgb = O.synthetic_pair(2, H, W)
    ts = TrainStep(CudaBackend(), 2, H, W, "cuda")
    ts.load(pG, pD, pV)
    ts.step(ir.cuda(), rgb.cuda())
    torch.cuda.synchronize()
    return ts


def test_full_step_256_losses_vs_reference(step):
    los = step.losses()
    for k, tol in (("D", 1e-2), ("G", 3e-3), ("L1", 3e-3), ("perc", 1e-2), ("TV", 2e-2), ("SSIM", 3e-3)):
        ref = float(GOLD["loss_" + k])
        print(k, los[k], ref)
        assert abs(los[k] - ref) <= tol * max(1.0, abs(ref)), (k, los[k], ref)
    e = rel_np(sample(step.G.fake, 4096), GOLD["fake_b2_sample"])
    print("fake (B=2) sample rel-L2", e)
    assert e < 3e-2


def _cos(a, b):
    return float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))


def test_full_step_256_gradients_vs_reference(step):
    e = rel_np(sample(step.dfake, 4096), GOLD["dfake_b2_sample"])
    c = _cos(sample(step.dfake, 4096), GOLD["dfake_b2_sample"])
    print("dfake sample rel-L2", e, "cos", c)
    assert c > 0.9          # d|f - r|/df = sign(f - r) flips wherever the bf16 forward error crosses the target
    worst_c, worst_n = {}, {}
    for net, arena, tag, floor in (("G", step.G.arena, "gG", 1e-4), ("D", step.D2.arena, "gD", 1e-5)):
        for k, g in arena.grads().items():
            if float(GOLD[f"{tag}_absmax/{k}"]) <= floor:
                if net == "G" and k.endswith("bias"):
                    assert g.abs().max().item() == 0     # bias in front of a non-affine InstanceNorm: exactly zero here, rounding noise in the reference
                continue
            c = _cos(sample(g), GOLD[f"{tag}_sample/{k}"])
            n = g.norm().item() / float(GOLD[f"{tag}_norm/{k}"])
            worst_c[f"{net}/{k}"] = c; worst_n[f"{net}/{k}"] = n
    lo = sorted(worst_c.items(), key=lambda kv: kv[1])[:5]
    print("lowest cosine similarities:", lo)
    print("norm ratios: min", min(worst_n.values()), "max", max(worst_n.values()))
    for k, c in worst_c.items():
        assert c > (0.95 if k.startswith("G/") else 0.9), (k, c)
    for k, n in worst_n.items():
        assert 0.9 < n < 1.1, (k, n)


def test_full_step_256_post_adam_parameters_vs_reference(step):
    """Adam's first step moves every weight by +-lr * g / (|g| + eps): the updated parameters agree wherever the gradient
    sign does"""
    lr = 2e-4
    for k in ("model.0.weight", "model.2.weight", "model.5.weight", "model.8.weight", "model.11.weight"):
        got = sample(step.D2.arena.view(k)); want = GOLD["pD_after_sample/" + k]
        bad = (np.abs(got - want) > 0.5 * lr).mean()
        print("post-Adam D", k, "fraction off by more than lr/2:", bad)
        assert bad < 0.15, (k, bad)
    for k in ("outc.1.weight", "outc.1.bias", "up2_conv.0.weight", "resblocks.4.conv_block.1.weight", "inc.1.weight"):
        got = sample(step.G.arena.view(k)); want = GOLD["pG_after_sample/" + k]
        bad = (np.abs(got - want) > 0.5 * lr).mean()
        print("post-Adam G", k, "fraction off by more than lr/2:", bad)
        assert bad < 0.2, (k, bad)
