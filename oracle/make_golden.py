"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference/Code) on seeded inputs, and assert the oracle restatement matches it.

Run in the build container only (``python oracle/make_golden.py``): /root/reference does
not exist on the GPU box.  Test infrastructure — never imported by the product.

Weights come from ``irc_oracle.seeded_params`` (not from the reference's init order) and
are pushed into the reference modules with ``load_state_dict`` so that oracle, reference
and CUDA path all see identical values.  VGG-16 uses seeded stand-in weights because
the ImageNet file cannot be downloaded here.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/Code")

import irc_oracle as O  # noqa: E402
import ir_colorization as R  # noqa: E402
import torchvision  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
torch.set_num_threads(8)


def ref_generator(pG):
    cfg = R.Config(); cfg.device = "cpu"
    m = R.IRColorizationModel(cfg)
    missing = m.netG.load_state_dict(pG, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("filt") for k in missing.missing_keys)
    return m


def ref_discriminator(pD):
    d = R.NLayerDiscriminator(4, 64, 3, R.get_norm_layer("instance"))
    d.load_state_dict(pD)
    return d


def ref_vgg(pV):
    R.models = types.SimpleNamespace(
        vgg16=lambda **kw: torchvision.models.vgg16(weights=None),
        VGG16_Weights=torchvision.models.VGG16_Weights)
    v = R.VGGPerceptual(torch.device("cpu"))
    v.features.load_state_dict({k[len("features."):]: t for k, t in pV.items()})
    return v


def close(a, b, tol, what):
    err = (a - b).abs().max().item()
    scale = max(b.abs().max().item(), 1e-12)
    assert err <= tol * max(scale, 1.0), f"{what}: max err {err} (scale {scale})"
    return err


def sample(t, n=256):
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].numpy()


def main():
    os.makedirs(OUT, exist_ok=True)
    B, H, W = 2, 32, 32
    pG = O.seeded_params(O.generator_shapes(), 1234, bias_std=0.02)
    pD = O.seeded_params(O.discriminator_shapes(), 1235, bias_std=0.02)
    pV = O.seeded_params(O.vgg_shapes(), 1236, kaiming=True, bias_std=0.05)
    ir, rgb = O.synthetic_pair(B, H, W)
    gold = {}

    # ---- schedule (irc:212-233) -------------------------------------------------
    lam = R.get_lr_lambda(R.Config())
    gold["lr_factor"] = np.array([lam(e) for e in range(50)])
    assert np.allclose(gold["lr_factor"], [O.lr_factor(e) for e in range(50)])

    # ---- stencils ----------------------------------------------------------------
    g = torch.Generator().manual_seed(3)
    xs = torch.randn(2, 5, 12, 10, generator=g)
    xo = torch.randn(1, 3, 9, 7, generator=g)  # odd sizes
    for name, x in (("even", xs), ("odd", xo)):
        C = x.shape[1]
        rd = R.Downsample(C)(x); ru = R.UpsampleAA(C)(x)
        close(O.blur_down(x), rd, 1e-6, "blur_down " + name)
        close(O.upsample_aa(x), ru, 1e-6, "upsample_aa " + name)
        gold[f"stencil_{name}_in"] = x.numpy(); gold[f"down_{name}"] = rd.numpy(); gold[f"up_{name}"] = ru.numpy()

    # ---- losses --------------------------------------------------------------------
    a = torch.rand(2, 3, 24, 20, generator=g); b = torch.rand(2, 3, 24, 20, generator=g)
    a.requires_grad_(True)
    rs = R.ssim_loss_torch(a, b); rt = R.tv_loss(a)
    (gs,) = torch.autograd.grad(rs, a, retain_graph=True); (gt,) = torch.autograd.grad(rt, a)
    a2 = a.detach().clone().requires_grad_(True)
    os_, ot = O.ssim_loss(a2, b), O.tv_loss(a2)
    (ogs,) = torch.autograd.grad(os_, a2, retain_graph=True); (ogt,) = torch.autograd.grad(ot, a2)
    close(os_, rs, 1e-6, "ssim"); close(ot, rt, 1e-6, "tv"); close(ogs, gs, 1e-5, "ssim grad"); close(ogt, gt, 1e-6, "tv grad")
    close(O.ssim_loss(a2, b, size_average=False), R.ssim_loss_torch(a, b, size_average=False), 1e-6, "ssim per-sample")
    gold.update(loss_a=a.detach().numpy(), loss_b=b.numpy(), ssim=rs.item(), tv=rt.item(),
                ssim_grad=gs.numpy(), tv_grad=gt.numpy(),
                ssim_per_sample=R.ssim_loss_torch(a, b, size_average=False).detach().numpy())

    # ---- networks forward ----------------------------------------------------------
    mG = ref_generator(pG); mD = ref_discriminator(pD); mV = ref_vgg(pV)
    with torch.no_grad():
        taps = {}
        fake_ref = mG(ir)
        fake_or = O.generator_forward(pG, ir, taps=taps)
        close(fake_or, fake_ref, 2e-5, "generator fwd")
        # hooks on the reference for per-layer parity
        ref_taps = {}
        x0 = mG.netG.inc(ir); ref_taps["x0"] = x0
        x1 = mG.netG.down1_down(mG.netG.down1(x0)); ref_taps["x1"] = x1
        x2 = mG.netG.down2_down(mG.netG.down2(x1)); ref_taps["x2"] = x2
        h = x2
        for i, blk in enumerate(mG.netG.resblocks):
            h = blk(h); ref_taps[f"res{i}"] = h
        ref_taps["up1_up"] = mG.netG.up1_up(h)
        for k, v in ref_taps.items():
            close(taps[k], v, 2e-5, "G tap " + k)
            gold["G_" + k + "_sample"] = sample(v); gold["G_" + k + "_absmean"] = v.abs().mean().item()
        # eval == train for this generator
        mG.eval(); assert torch.equal(mG(ir), fake_ref); mG.train()
        din = torch.cat([ir, rgb], 1)
        dref = mD(din); close(O.discriminator_forward(pD, din), dref, 2e-5, "D fwd")
        vref = mV(rgb); close(O.vgg_forward(pV, rgb), vref, 2e-5 * vref.abs().max().item(), "VGG fwd")
        # standalone ResnetBlock
        blk = mG.netG.resblocks[0]
        close(O.resnet_block(pG, "resblocks.0.", x2), blk(x2), 2e-5, "resblock")
    gold.update(fake=fake_ref.numpy(), d_real=dref.numpy(), vgg_rgb_sample=sample(vref),
                vgg_rgb_absmean=vref.abs().mean().item(), res0_out=ref_taps["res0"].numpy(), x2=ref_taps["x2"].numpy())

    # ---- one train step with the reference modules + torch.optim.Adam (irc:1636-1681) ----
    cfg = R.Config()
    optG = torch.optim.Adam(mG.netG.parameters(), lr=cfg.lr_G, betas=(cfg.beta1, cfg.beta2))
    optD = torch.optim.Adam(mD.parameters(), lr=cfg.lr_D, betas=(cfg.beta1, cfg.beta2))
    optD.zero_grad()
    with torch.no_grad():
        fake_det = mG(ir)
    pred_real = mD(torch.cat([ir, rgb], 1)); pred_fake = mD(torch.cat([ir, fake_det], 1))
    loss_D = 0.5 * (torch.relu(1.0 - pred_real).mean() + torch.relu(1.0 + pred_fake).mean())
    loss_D.backward()
    gD_ref = {k: p.grad.clone() for k, p in mD.named_parameters()}
    optD.step()
    optG.zero_grad()
    fake = mG(ir); fake.retain_grad()
    gan = -mD(torch.cat([ir, fake], 1)).mean()
    l1 = torch.nn.L1Loss()(fake, rgb) * cfg.lambda_L1
    perc = torch.nn.functional.l1_loss(mV(fake), mV(rgb)) * cfg.lambda_perc
    tv = R.tv_loss(fake) * cfg.lambda_tv
    ssim = R.ssim_loss_torch((fake + 1.0) / 2.0, (rgb + 1.0) / 2.0) * cfg.lambda_ssim
    loss_G = cfg.lambda_gan * gan + l1 + perc + tv + ssim
    loss_G.backward()
    gG_ref = {k: p.grad.clone() for k, p in mG.netG.named_parameters()}
    optG.step()

    # oracle step on copies
    oG = {k: v.clone() for k, v in pG.items()}; oD = {k: v.clone() for k, v in pD.items()}
    aG, aD = O.AdamState(oG), O.AdamState(oD)
    losses, gG_or, gD_or = O.train_step(oG, oD, pV, aG, aD, ir, rgb)
    close(losses["D"], loss_D.detach(), 1e-5, "loss_D")
    for k, r in (("G", loss_G), ("GAN", gan), ("L1", l1), ("perc", perc), ("TV", tv), ("SSIM", ssim)):
        close(losses[k], r.detach(), 2e-5, "loss " + k)
    for k in gD_ref:
        close(gD_or[k], gD_ref[k], 1e-4 * max(1.0, 1.0 / max(gD_ref[k].abs().max().item(), 1e-3)), "gD " + k)
    big = [k for k in gG_ref if gG_ref[k].abs().max() > 1e-4]
    for k in big:
        rel = ((gG_or[k] - gG_ref[k]).norm() / gG_ref[k].norm()).item()
        assert rel < 1e-3, f"gG {k}: rel {rel}"
    # post-Adam weights (only tensors with real gradients: noise-gradient biases excluded)
    for k, p_ in mD.named_parameters():
        if gD_ref[k].abs().max() > 1e-4:
            # first Adam step is +-lr * g/(|g|+eps): elements with |g| ~ eps are rounding-sensitive
            frac_bad = ((oD[k] - p_.detach()).abs() > 2e-5).float().mean().item()
            assert frac_bad < 1e-3, f"post-Adam D {k}: {frac_bad}"
    for k, p_ in mG.netG.named_parameters():
        if k.endswith("weight"):
            frac_bad = ((oG[k] - p_.detach()).abs() > 1e-4).float().mean().item()
            assert frac_bad < 1e-3, f"post-Adam G {k}: {frac_bad}"

    gold.update(loss_D=loss_D.item(), loss_G=loss_G.item(), loss_GAN=gan.item(), loss_L1=l1.item(),
                loss_perc=perc.item(), loss_TV=tv.item(), loss_SSIM=ssim.item(),
                dfake=fake.grad.numpy())
    for k, v in gD_ref.items():
        gold["gD_norm/" + k] = v.norm().item(); gold["gD_sample/" + k] = sample(v)
    for k, v in gG_ref.items():
        gold["gG_norm/" + k] = v.norm().item(); gold["gG_sample/" + k] = sample(v)
    for k, p_ in mD.named_parameters():
        gold["pD_after_sample/" + k] = sample(p_)
    gold["pG_after_sample/outc.1.weight"] = sample(dict(mG.netG.named_parameters())["outc.1.weight"])
    gold["pG_after_sample/outc.1.bias"] = sample(dict(mG.netG.named_parameters())["outc.1.bias"])

    # ---- test-mode core (irc:865-876, irc:1184-1205) -------------------------------
    with torch.no_grad():
        img = R.tensor_to_rgb_image(fake_ref)
    assert np.array_equal(img, O.quantize_u8(fake_ref[0]))
    gt = torch.rand(H, W, 3, generator=g).numpy().astype(np.float32)
    mae, mse, psnr, ssim_v = R.compute_metrics(img.astype(np.float32) / 255.0, gt)
    assert ssim_v is None, "skimage unexpectedly present: pin skimage_ssim against it!"
    assert (mae, mse, psnr) == O.compute_metrics(img.astype(np.float32) / 255.0, gt)
    gold.update(quant_u8=img, metrics_gt=gt, metrics=np.array([mae, mse, psnr]))
    # truncation probe from the survey: 254.87 -> 254
    probe = torch.full((1, 3, 1, 1), 254.87 / 255.0 * 2 - 1)
    assert R.tensor_to_rgb_image(probe)[0, 0, 0] == 254 == O.quantize_u8(probe[0])[0, 0, 0]

    np.savez_compressed(os.path.join(OUT, "ref_small.npz"), **gold)
    print("wrote", os.path.join(OUT, "ref_small.npz"), "keys:", len(gold))
    print({k: round(float(gold[k]), 6) for k in ("loss_D", "loss_G", "loss_GAN", "loss_L1", "loss_perc", "loss_TV", "loss_SSIM")})


if __name__ == "__main__":
    main()
