"""Host logic of the product (frame geometry, tap tables, weight packing maps, gradient routing, optimizer
plumbing, the module surface) driven with the torch restatement of the primitives (tests/ref_backend.py) in
float32 storage, compared with the oracle and the reference's golden vectors.  No GPU, no CUDA library calls."""
import os

import numpy as np
import pytest
import torch

import irc_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_small.npz"))
B, H, W = 2, 32, 32


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture()
def fp32_frames():
    import irc_b200  # noqa: F401
    from irc_b200 import layout as L, modules as M
    from ref_backend import RefBackend
    old, oldbe = L.ACT_DTYPE, M._BACKEND
    L.ACT_DTYPE = torch.float32
    M.set_backend(RefBackend())
    yield M._BACKEND
    L.ACT_DTYPE = old
    M.set_backend(oldbe)


def test_stencil_tables_match_reference_modules():
    from irc_b200 import layout as L
    for name in ("even", "odd"):
        x = torch.from_numpy(GOLD[f"stencil_{name}_in"]).double()
        h, w = x.shape[-2:]
        D = lambda n: torch.from_numpy(L.down_matrix(n)); U = lambda n: torch.from_numpy(L.up_matrix(n))
        assert (D(h) @ x @ D(w).T - torch.from_numpy(GOLD[f"down_{name}"])).abs().max() < 1e-6
        assert (U(h) @ x @ U(w).T - torch.from_numpy(GOLD[f"up_{name}"])).abs().max() < 2e-6
    f = L.fold_matrix(6, 3)
    assert f.shape == (6, 12) and f.sum() == 12 and f[3, 0] == 1 and f[1, 2] == 1 and f[3, 10] == 1 and f[4, 7] == 1


def test_train_step_plan_matches_reference(fp32_frames):
    from irc_b200.train_step import TrainStep
    pG = O.seeded_params(O.generator_shapes(), 1234, bias_std=0.02)
    pD = O.seeded_params(O.discriminator_shapes(), 1235, bias_std=0.02)
    pV = O.seeded_params(O.vgg_shapes(), 1236, kaiming=True, bias_std=0.05)
    ir, rgb = O.synthetic_pair(B, H, W)
    ts = TrainStep(fp32_frames, B, H, W, "cpu")
    ts.load(pG, pD, pV)
    ts.step(ir, rgb)
    los = ts.losses()
    assert ts.losses_async().get() == los          # the handle form returns the same terms (on the CPU it is a plain copy)
    for k in ("D", "G", "GAN", "L1", "perc", "TV", "SSIM"):
        ref = float(GOLD["loss_" + k])
        assert abs(los[k] - ref) < 3e-5 * max(1.0, abs(ref)), (k, los[k], ref)
    assert np.abs(ts.G.fake.numpy() - GOLD["fake"]).max() < 5e-5
    assert rel(ts.dfake, torch.from_numpy(GOLD["dfake"])) < 1e-4
    oG = {k: v.clone() for k, v in pG.items()}; oD = {k: v.clone() for k, v in pD.items()}
    _, gG, gD = O.train_step(oG, oD, pV, O.AdamState(oG), O.AdamState(oD), ir, rgb)
    for k, g in ts.D2.arena.grads().items():
        if gD[k].abs().max() > 1e-5:
            assert rel(g, gD[k]) < 1e-4, k
    for k, g in ts.G.arena.grads().items():
        if gG[k].abs().max() > 1e-4:
            assert rel(g, gG[k]) < 1e-4, k
        elif k.endswith("bias"):
            assert g.abs().max() == 0      # bias in front of a non-affine InstanceNorm: exactly zero
    for k, v in ts.D2.arena.state_dict().items():
        if gD[k].abs().max() > 1e-4:
            assert ((v - oD[k]).abs() > 2e-5).float().mean() < 1e-3, k


def test_module_surface_and_state_dict(fp32_frames):
    import irc_b200 as R
    cfg = R.Config(); cfg.device = "cpu"
    model = R.IRColorizationModel(cfg)
    sd = model.netG.state_dict()
    want = set(O.generator_shapes()) | {"down1_down.filt", "down2_down.filt", "up1_up.filt", "up2_up.filt"}
    assert set(sd) == want and len(sd) == 52
    assert tuple(sd["down1_down.filt"].shape) == (128, 1, 3, 3) and abs(sd["up2_up.filt"][0, 0].sum().item() - 1) < 1e-6
    netD = R.NLayerDiscriminator(4, 64, 3, R.get_norm_layer("instance"))
    assert set(netD.state_dict()) == set(O.discriminator_shapes())
    lam = R.get_lr_lambda(cfg)
    assert np.allclose([lam(e) for e in range(50)], GOLD["lr_factor"])
    for bad in (dict(ngf=32), dict(use_dropout=True), dict(padding_type="zero")):      # variants that are still not built say so
        with pytest.raises(NotImplementedError):
            R.ResnetUNetGenerator(1, 3, **bad)


def test_modules_autograd_matches_reference(fp32_frames):
    import irc_b200 as R
    pG = O.seeded_params(O.generator_shapes(), 1234, bias_std=0.02)
    pD = O.seeded_params(O.discriminator_shapes(), 1235, bias_std=0.02)
    ir, rgb = O.synthetic_pair(B, H, W)
    cfg = R.Config(); cfg.device = "cpu"
    model = R.IRColorizationModel(cfg)
    model.netG.load_state_dict(pG, strict=False)
    netD = R.NLayerDiscriminator(4, 64, 3, R.get_norm_layer("instance"))
    netD.load_state_dict(pD)
    # reference loop shape: D called twice before one backward (irc:1642-1650)
    with torch.no_grad():
        fake_d = model(ir)
    assert np.abs(fake_d.numpy() - GOLD["fake"]).max() < 5e-5
    pred_real = netD(torch.cat([ir, rgb], 1)); pred_fake = netD(torch.cat([ir, fake_d], 1))
    loss_D = 0.5 * (torch.relu(1.0 - pred_real).mean() + torch.relu(1.0 + pred_fake).mean())
    assert abs(loss_D.item() - float(GOLD["loss_D"])) < 1e-5
    loss_D.backward()
    _, gD = O.d_loss_and_grads(pD, ir, rgb, fake_d)
    for k, p in netD.named_parameters():
        if gD[k].abs().max() > 1e-5:
            assert rel(p.grad, gD[k]) < 1e-4, k
    # generator through the module + torch losses
    fake = model(ir)
    loss = 30.0 * (fake - rgb).abs().mean() + 1e-4 * R.tv_loss(fake) + 2.0 * R.ssim_loss_torch((fake + 1) / 2, (rgb + 1) / 2)
    loss.backward()
    leaves = {k: v.clone().requires_grad_(True) for k, v in pG.items()}
    f2 = O.generator_forward(leaves, ir)
    (30.0 * (f2 - rgb).abs().mean() + 1e-4 * O.tv_loss(f2) + 2.0 * O.ssim_loss((f2 + 1) / 2, (rgb + 1) / 2)).backward()
    for k, p in model.netG.named_parameters():
        if leaves[k].grad.abs().max() > 1e-4:
            assert rel(p.grad, leaves[k].grad) < 2e-4, k


def test_small_modules_match_golden(fp32_frames):
    import irc_b200 as R
    for name in ("even", "odd"):
        x = torch.from_numpy(GOLD[f"stencil_{name}_in"]).requires_grad_(True)
        C = x.shape[1]
        d = R.Downsample(C)(x); u = R.UpsampleAA(C)(x)
        assert np.abs(d.detach().numpy() - GOLD[f"down_{name}"]).max() < 1e-6
        assert np.abs(u.detach().numpy() - GOLD[f"up_{name}"]).max() < 2e-6
        gd = torch.autograd.grad(d.sum() + (u * u).sum(), x)[0]
        x2 = x.detach().clone().requires_grad_(True)
        ref = torch.autograd.grad(O.blur_down(x2).sum() + (O.upsample_aa(x2) ** 2).sum(), x2)[0]
        assert (gd - ref).abs().max() < 1e-4
    a = torch.from_numpy(GOLD["loss_a"]).requires_grad_(True); b = torch.from_numpy(GOLD["loss_b"])
    s, t = R.ssim_loss_torch(a, b), R.tv_loss(a)
    assert abs(s.item() - float(GOLD["ssim"])) < 1e-6 and abs(t.item() - float(GOLD["tv"])) < 1e-6
    (gs,) = torch.autograd.grad(s, a, retain_graph=True); (gt,) = torch.autograd.grad(t, a)
    assert np.abs(gs.numpy() - GOLD["ssim_grad"]).max() < 1e-6 and np.abs(gt.numpy() - GOLD["tv_grad"]).max() < 1e-7
    per = R.ssim_loss_torch(a, b, size_average=False)
    assert np.abs(per.detach().numpy() - GOLD["ssim_per_sample"]).max() < 1e-6
    # ResnetBlock with the generator's first block weights
    pG = O.seeded_params(O.generator_shapes(), 1234, bias_std=0.02)
    blk = R.ResnetBlock(256, "reflect", torch.nn.InstanceNorm2d, False, True)
    blk.load_state_dict({k[len("resblocks.0."):]: v for k, v in pG.items() if k.startswith("resblocks.0.")})
    x2 = torch.from_numpy(GOLD["x2"]).requires_grad_(True)
    y = blk(x2)
    assert np.abs(y.detach().numpy() - GOLD["res0_out"]).max() < 5e-5
    y.sum().backward()
    xr = torch.from_numpy(GOLD["x2"]).requires_grad_(True)
    O.resnet_block(pG, "resblocks.0.", xr).sum().backward()
    assert rel(x2.grad, xr.grad) < 1e-4


def test_vgg_module_and_test_mode(fp32_frames):
    import irc_b200 as R
    import warnings
    pV = O.seeded_params(O.vgg_shapes(), 1236, kaiming=True, bias_std=0.05)
    _, rgb = O.synthetic_pair(B, H, W)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vgg = R.VGGPerceptual("cpu")
    assert set(k for k in vgg.state_dict()) == set(O.vgg_shapes()) | {"mean", "std"}
    vgg.load_state_dict(pV, strict=False)
    x = rgb.clone().requires_grad_(True)
    f = vgg(x)
    xr = rgb.clone().requires_grad_(True)
    fr = O.vgg_forward(pV, xr)
    assert rel(f.detach(), fr.detach()) < 1e-5
    f.abs().mean().backward(); fr.abs().mean().backward()
    assert rel(x.grad, xr.grad) < 1e-4
    # test-mode core against the reference's own output
    fake = torch.from_numpy(GOLD["fake"])
    assert np.array_equal(R.tensor_to_rgb_image(fake), GOLD["quant_u8"])
    gt = torch.from_numpy(GOLD["metrics_gt"]).permute(2, 0, 1).unsqueeze(0)
    u8, mae, mse, psnr = R.batch_metrics(fake[:1], gt)
    assert np.allclose([mae[0], mse[0], psnr[0]], GOLD["metrics"], rtol=1e-6)
    m = R.compute_metrics(GOLD["quant_u8"].astype(np.float32) / 255.0, GOLD["metrics_gt"])
    assert np.allclose(m[:3], GOLD["metrics"], rtol=0, atol=0) and m[3] is None


def test_ssim_metric_restatement_matches_the_scipy_oracle():
    """the SSIM column (irc:1208-1215): the batched restatement (avg_pool2d over the valid region, float64) against the per-image
    scipy.ndimage.uniform_filter oracle of skimage's documented defaults (parity unpinned: scikit-image is not in this image)"""
    from ref_backend import RefBackend
    g = torch.Generator().manual_seed(21)
    n, H, W = 2, 19, 23
    gt = torch.rand(n, 3, H, W, generator=g)
    u8 = (torch.rand(n, H, W, 3, generator=g) * 255).to(torch.uint8)
    u8[0] = (gt[0].permute(1, 2, 0) * 255).to(torch.uint8)          # one nearly identical pair
    sums = torch.zeros(n, dtype=torch.float64)
    RefBackend().ssim_metric(u8, gt, sums)
    for i in range(n):
        want = O.skimage_ssim(gt[i].permute(1, 2, 0).numpy(), u8[i].numpy().astype(np.float32) / 255.0)
        assert abs(sums[i].item() / (3 * (H - 6) * (W - 6)) - want) < 1e-10
    assert sums[0].item() / (3 * (H - 6) * (W - 6)) > 0.99


def test_stream_window_classifies_the_stencil_tables():
    """host-side gate of the streaming stencil kernels: the anti-aliased down/up-sampling operators and their transposes have
    a non-decreasing last source row and a bounded window; the reflection-fold operator does not (it falls back)"""
    from irc_b200 import layout as L
    for n in (8, 12, 64):
        cases = {"down": (L.down_matrix(n), 3), "up": (L.up_matrix(n), 3), "downT": (L.down_matrix(n).T, 2), "upT": (L.up_matrix(n).T, 6)}
        for name, (m, k) in cases.items():
            t = L.make_tables(m, m, "cpu")
            assert t.stream_window() == k, (name, n, t.stream_window())
            assert t.stream_window(8) == k
        fold = L.make_tables(L.fold_matrix(n, 1), L.fold_matrix(n, 1), "cpu")
        assert fold.stream_window() == 0
        # identity along one axis: not a streaming case
        assert L.make_tables(L.down_matrix(n), None, "cpu").stream_window() == 0


def test_epoch_means_average_every_step_and_adam_counter_lives_on_the_device(fp32_frames):
    """irc:1683-1697 averages the D and G losses of EVERY iteration; the step accumulates them on the device.  The Adam
    step counter is a device tensor advanced by the optimizer launch itself (no per-step host write)."""
    from irc_b200.train_step import TrainStep
    pG = O.seeded_params(O.generator_shapes(), 1234, bias_std=0.02)
    pD = O.seeded_params(O.discriminator_shapes(), 1235, bias_std=0.02)
    pV = O.seeded_params(O.vgg_shapes(), 1236, kaiming=True, bias_std=0.05)
    ts = TrainStep(fp32_frames, 1, H, W, "cpu")
    ts.load(pG, pD, pV)
    ts.reset_epoch_sums()
    ds, gs = [], []
    for r in range(3):
        ir, rgb = O.synthetic_pair(1, H, W, rank=r)
        ts.step(ir, rgb, lr_scale=1.0 if r < 2 else 0.5)
        l = ts.losses(); ds.append(l["D"]); gs.append(l["G"])
    d, g, n = ts.epoch_means()
    assert n == 3 and abs(d - sum(ds) / 3) < 1e-5 and abs(g - sum(gs) / 3) < 1e-4
    assert int(ts.optG.step_dev.item()) == 3 and int(ts.optD.step_dev.item()) == 3 and ts.optG.t == 3
    assert ts.optG.dev.dtype == torch.float64 and abs(ts.optG.dev[4].item() - 0.5) < 1e-15     # LambdaLR factor of the last step
    # against the oracle's Adam over the same three batches
    oG = {k: v.clone() for k, v in pG.items()}; oD = {k: v.clone() for k, v in pD.items()}
    aG, aD = O.AdamState(oG), O.AdamState(oD)
    for r in range(3):
        ir, rgb = O.synthetic_pair(1, H, W, rank=r)
        O.train_step(oG, oD, pV, aG, aD, ir, rgb, lr_scale=1.0 if r < 2 else 0.5)
    k = "outc.1.weight"
    assert ((ts.G.arena.view(k) - oG[k]).abs() > 1e-4).float().mean() < 0.02


def test_topk_ranking_csv_matches_the_reference_file(tmp_path):
    """irc:1220-1278: the ranking CSV written by the reference's save_best_k_outputs for the same rows (golden file
    tests/golden/top_5_ranking.csv, produced by oracle/make_golden_256.py)"""
    import irc_b200 as R
    from irc_b200.train import write_topk_ranking
    g = torch.Generator().manual_seed(11)
    rows = []
    for i in range(9):
        mse = float(torch.rand(1, generator=g)) * 0.05 + 1e-3
        rows.append({"file": os.path.join(f"set0{i % 2}", f"V00{i % 3}", f"I{i:05d}.jpg"), "mae": float(torch.rand(1, generator=g)) * 0.2,
                     "mse": mse, "psnr": -10.0 * float(np.log10(mse + 1e-12)), "ssim": None})
    rows[4]["psnr"] = rows[2]["psnr"]
    rows[6]["psnr"] = float("inf")
    cfg = R.Config(); cfg.output_dir = str(tmp_path); cfg.topk = 5
    path = write_topk_ranking(cfg, rows)
    assert os.path.basename(path) == "top_5_ranking.csv" and os.path.basename(os.path.dirname(path)) == cfg.best50_dirname
    want = open(os.path.join(os.path.dirname(__file__), "golden", "top_5_ranking.csv"), encoding="utf-8").read()
    assert open(path, encoding="utf-8").read() == want
