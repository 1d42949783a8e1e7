"""Drop-in module surface on the GPU (every op through libirc_sm100.so) vs the reference's golden vectors."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_small.npz"))


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm()).item()


def test_fp32_stencils_and_losses_match_reference():
    """fp32 kernels: tolerance 1e-5 (north star: 1e-4 in fp32)"""
    import irc_b200 as R
    for name in ("even", "odd"):
        x = torch.from_numpy(GOLD[f"stencil_{name}_in"]).cuda().requires_grad_(True)
        C = x.shape[1]
        d = R.Downsample(C).cuda()(x); u = R.UpsampleAA(C).cuda()(x)
        assert np.abs(d.detach().cpu().numpy() - GOLD[f"down_{name}"]).max() < 1e-5
        assert np.abs(u.detach().cpu().numpy() - GOLD[f"up_{name}"]).max() < 1e-5
        (g,) = torch.autograd.grad(d.sum() + (u * u).sum(), x)
        import irc_oracle as O
        x2 = torch.from_numpy(GOLD[f"stencil_{name}_in"]).requires_grad_(True)
        (gr,) = torch.autograd.grad(O.blur_down(x2).sum() + (O.upsample_aa(x2) ** 2).sum(), x2)
        assert (g.cpu() - gr).abs().max() < 1e-4
    a = torch.from_numpy(GOLD["loss_a"]).cuda().requires_grad_(True); b = torch.from_numpy(GOLD["loss_b"]).cuda()
    s, t = R.ssim_loss_torch(a, b), R.tv_loss(a)
    assert abs(s.item() - float(GOLD["ssim"])) < 2e-5 and abs(t.item() - float(GOLD["tv"])) < 1e-6
    (gs,) = torch.autograd.grad(s, a, retain_graph=True); (gt,) = torch.autograd.grad(t, a)
    assert rel(gs, torch.from_numpy(GOLD["ssim_grad"])) < 1e-3 and np.abs(gt.cpu().numpy() - GOLD["tv_grad"]).max() < 1e-7


def test_resnet_block_single_layer_tolerance():
    """one block, reference input: the per-layer bf16 tolerance of the north star (rel 1e-2)"""
    import irc_oracle as O
    import irc_b200 as R
    pG = O.seeded_params(O.generator_shapes(), 1234, bias_std=0.02)
    blk = R.ResnetBlock(256, "reflect", torch.nn.InstanceNorm2d, False, True).cuda()
    blk.load_state_dict({k[len("resblocks.0."):]: v for k, v in pG.items() if k.startswith("resblocks.0.")})
    x = torch.from_numpy(GOLD["x2"]).cuda().requires_grad_(True)
    y = blk(x)
    e = rel(y, torch.from_numpy(GOLD["res0_out"]))
    print("resblock rel", e)
    assert e < 1e-2
    g = torch.randn(y.shape, generator=torch.Generator().manual_seed(0))
    y.backward(g.cuda())
    xr = torch.from_numpy(GOLD["x2"]).requires_grad_(True)
    leaves = {k: v.clone().requires_grad_(True) for k, v in pG.items() if k.startswith("resblocks.0.")}
    O.resnet_block(leaves, "resblocks.0.", xr).backward(g)
    ex = rel(x.grad, xr.grad)
    ew = max(rel(p.grad, leaves["resblocks.0." + k].grad) for k, p in blk.named_parameters() if k.endswith("weight"))
    print("resblock dgrad rel", ex, "wgrad rel", ew)
    # gradients inherit ReLU-mask flips: a forward error of ~5e-3 moves ~0.4% of the pre-activations across zero, and each
    # flipped mask changes its gradient contribution by 100% (error ~ sqrt(flip fraction)); see DESIGN.md "precision"
    assert ex < 0.12 and ew < 0.12


def test_generator_and_discriminator_modules():
    import irc_oracle as O
    import irc_b200 as R
    pG = O.seeded_params(O.generator_shapes(), 1234, bias_std=0.02)
    pD = O.seeded_params(O.discriminator_shapes(), 1235, bias_std=0.02)
    ir, rgb = O.synthetic_pair(2, 32, 32)
    cfg = R.Config(); cfg.device = "cuda"
    model = R.IRColorizationModel(cfg)
    model.netG.load_state_dict(pG, strict=False)
    assert len(model.netG.state_dict()) == 52 and next(model.netG.parameters()).is_cuda
    netD = R.init_net(R.NLayerDiscriminator(4, 64, 3, R.get_norm_layer("instance")), device="cuda")
    netD.load_state_dict(pD)
    irc, rgbc = ir.cuda(), rgb.cuda()
    with torch.no_grad():
        fake_d = model(irc)
    assert rel(fake_d, torch.from_numpy(GOLD["fake"])) < 4e-2
    model.eval()
    with torch.no_grad():
        assert torch.equal(model(irc), fake_d)          # eval == train for this generator; also deterministic
    model.train()
    pred_real = netD(torch.cat([irc, rgbc], 1)); pred_fake = netD(torch.cat([irc, fake_d], 1))
    assert rel(pred_real, torch.from_numpy(GOLD["d_real"])) < 3e-2
    loss_D = 0.5 * (torch.relu(1.0 - pred_real).mean() + torch.relu(1.0 + pred_fake).mean())
    loss_D.backward()
    optD = torch.optim.Adam(netD.parameters(), lr=2e-4, betas=(0.5, 0.999)); optD.step()     # torch optimizers work on the arena views
    fake = model(irc)
    loss = 30.0 * torch.nn.L1Loss()(fake, rgbc) + 1e-4 * R.tv_loss(fake) + 2.0 * R.ssim_loss_torch((fake + 1) / 2, (rgbc + 1) / 2) \
        - 0.1 * netD(torch.cat([irc, fake], 1)).mean()
    loss.backward()
    g = dict(model.netG.named_parameters())["outc.1.weight"].grad
    assert g is not None and torch.isfinite(g).all() and g.abs().max() > 0
    torch.save(model.netG.state_dict(), "/tmp/netG_test.pth")
    m2 = R.IRColorizationModel(cfg); m2.load_weights("/tmp/netG_test.pth")
    with torch.no_grad():
        assert torch.equal(m2(irc), model(irc))


def test_drivers_run_on_synthetic_pairs(tmp_path):
    import irc_b200 as R
    cfg = R.Config()
    cfg.device = "cuda"; cfg.img_size = 32; cfg.batch_size = 2; cfg.epochs = 2; cfg.synthetic_steps = 3
    cfg.save_dir = str(tmp_path / "ckpt"); cfg.output_dir = str(tmp_path / "res"); cfg.save_every = 1; cfg.lr_decay_start_epoch = 1
    cfg.mode = "train"
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hist = R.main(cfg)
    assert len(hist) == 2 and all(np.isfinite(h[2]) for h in hist)
    assert os.path.isfile(os.path.join(cfg.save_dir, "netG_best.pth")) and os.path.isfile(os.path.join(cfg.save_dir, "netG_epoch_002.pth"))
    # resume from the full-state checkpoint: one more epoch on top of the two
    cfg.save_full_state = True; cfg.epochs = 3
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        R.main(cfg)                                                     # writes train_state_latest.pth (epoch 3)
        cfg.resume_from = os.path.join(cfg.save_dir, "train_state_latest.pth"); cfg.epochs = 4
        hist2 = R.main(cfg)
    assert len(hist2) == 1 and np.isfinite(hist2[0][2])                 # only epoch 4 ran
    cfg.resume_from = None
    cfg.mode = "test"; cfg.test_G_weights = os.path.join(cfg.save_dir, "netG_best.pth")
    summary, rows, preds = R.main(cfg)
    assert summary["count"] == 6 and len(rows) == 6 and preds[0].dtype == torch.uint8
    txt = open(os.path.join(cfg.output_dir, "metrics_test.csv")).read()
    assert txt.startswith("file,mae,mse,psnr,ssim\n") and "# mean_psnr," in txt
    # the CSV numbers equal the reference's compute_metrics on the same prediction
    m = R.compute_metrics(preds[0][0].numpy().astype(np.float32) / 255.0,
                          ((next(iter(R.train.SyntheticPairs(3, 2, 32, 32, seed=9)))["rgb"][0] + 1) * 0.5).permute(1, 2, 0).numpy())
    assert abs(m[0] - rows[0]["mae"]) < 1e-6 and abs(m[1] - rows[0]["mse"]) < 1e-6 and abs(m[2] - rows[0]["psnr"]) < 1e-4


def test_inference_reuses_its_engine():
    """under torch.no_grad() the generator module must hand its execution plan back to the pool (an engine rebuild costs
    0.3 s of host work: test-mode throughput, BASELINE configs 1 and 5, depends on it) and build it without backward buffers"""
    import irc_b200 as R
    cfg = R.Config(); cfg.device = "cuda"
    model = R.IRColorizationModel(cfg).eval()
    ir = torch.rand(2, 1, 32, 48, device="cuda") * 2 - 1
    with torch.no_grad():
        a = model(ir)
        pool = model.netG._free[(2, 32, 48, False)]
        assert len(pool) == 1
        eng = pool[0]
        assert not eng.training and not hasattr(eng, "dZ0")
        b = model(ir)
        assert model.netG._free[(2, 32, 48, False)] == [eng]
    assert torch.equal(a, b)
    out = model(ir)                      # grad mode: a training engine, kept alive by the autograd graph until backward
    assert (2, 32, 48, True) not in model.netG._free or model.netG._free[(2, 32, 48, True)] == []
    out.sum().backward()
    assert len(model.netG._free[(2, 32, 48, True)]) == 1
