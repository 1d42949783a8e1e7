"""Golden vectors of the NON-DEFAULT generator graph the build supports: no_antialias_up=True, i.e.
nn.ConvTranspose2d(C, C, 3, stride=2, padding=1, output_padding=1) instead of UpsampleAA (irc:495-499, :512-516), from the
UNMODIFIED reference.  Build container only.  Writes tests/golden/ref_variants.npz; asserts the oracle restatement on the way."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402

R, O = MG.R, MG.O


def main():
    B, H, W = 2, 32, 32
    gold = {}
    pG = O.seeded_params(O.generator_shapes(no_antialias_up=True), 4321, bias_std=0.02)
    ir, rgb = O.synthetic_pair(B, H, W)
    cfg = R.Config(); cfg.device = "cpu"; cfg.no_antialias_up = True
    m = R.IRColorizationModel(cfg)
    keys = set(m.netG.state_dict().keys())
    assert "up1_up.weight" in keys and "up2_up.bias" in keys and "up1_up.filt" not in keys
    assert tuple(m.netG.state_dict()["up1_up.weight"].shape) == (256, 256, 3, 3)
    missing = m.netG.load_state_dict(pG, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("filt") for k in missing.missing_keys), missing
    fake = m(ir)
    g = torch.randn(fake.shape, generator=torch.Generator().manual_seed(9))
    fake.backward(g)
    leaves = {k: v.clone().requires_grad_(True) for k, v in pG.items()}
    fo = O.generator_forward(leaves, ir)
    MG.close(fo, fake.detach(), 2e-5, "variant forward")
    fo.backward(g)
    for k, p_ in m.netG.named_parameters():
        noise = k.endswith("bias") and not (k.startswith("outc") or k.startswith("up1_up") or k.startswith("up2_up"))   # bias in front of a non-affine InstanceNorm
        if p_.grad.abs().max() > 1e-4 and not noise:
            rel = ((leaves[k].grad - p_.grad).norm() / p_.grad.norm()).item()
            assert rel < 5e-3, (k, rel)
        gold["grad_norm/" + k] = p_.grad.norm().item(); gold["grad_sample/" + k] = MG.sample(p_.grad); gold["grad_absmax/" + k] = p_.grad.abs().max().item()
    gold["fake"] = fake.detach().numpy(); gold["upstream"] = g.numpy()
    # the transposed convolution alone (irc:495-499): input, weight, bias, output, gradients
    ct = torch.nn.ConvTranspose2d(64, 64, 3, stride=2, padding=1, output_padding=1)
    gen = torch.Generator().manual_seed(3)
    with torch.no_grad():
        ct.weight.copy_(torch.randn(64, 64, 3, 3, generator=gen) * 0.05); ct.bias.copy_(torch.randn(64, generator=gen) * 0.1)
    x = torch.randn(2, 64, 6, 10, generator=gen, requires_grad=True)
    y = ct(x)
    gy = torch.randn(y.shape, generator=gen)
    y.backward(gy)
    gold.update(ct_x=x.detach().numpy(), ct_w=ct.weight.detach().numpy(), ct_b=ct.bias.detach().numpy(), ct_y=y.detach().numpy(), ct_gy=gy.numpy(),
                ct_gx=x.grad.numpy(), ct_gw=ct.weight.grad.numpy(), ct_gb=ct.bias.grad.numpy())
    # no_antialias=True (irc:468, :474, :482): stride-2 down-sampling convolutions, no blur modules; alone and together with
    # the transposed-convolution up-sampling ("na/" and "nab/" fixtures)
    for tag, up in (("na", False), ("nab", True)):
        pV = O.seeded_params(O.generator_shapes(no_antialias_up=up), 777 + int(up), bias_std=0.02)
        cfg = R.Config(); cfg.device = "cpu"; cfg.no_antialias = True; cfg.no_antialias_up = up
        m = R.IRColorizationModel(cfg)
        keys = set(m.netG.state_dict().keys())
        assert "down1_down.filt" not in keys and "down2_down.filt" not in keys and ("up1_up.filt" in keys) == (not up)
        missing = m.netG.load_state_dict(pV, strict=False)
        assert not missing.unexpected_keys and all(k.endswith("filt") for k in missing.missing_keys), missing
        fake = m(ir)
        gv = torch.randn(fake.shape, generator=torch.Generator().manual_seed(11))
        fake.backward(gv)
        leaves = {k: v.clone().requires_grad_(True) for k, v in pV.items()}
        fo = O.generator_forward(leaves, ir, no_antialias=True)
        MG.close(fo, fake.detach(), 2e-5, tag + " forward")
        fo.backward(gv)
        for k, p_ in m.netG.named_parameters():
            noise = k.endswith("bias") and not (k.startswith("outc") or k.startswith("up1_up") or k.startswith("up2_up"))
            if p_.grad.abs().max() > 1e-4 and not noise:
                rel = ((leaves[k].grad - p_.grad).norm() / p_.grad.norm()).item()
                assert rel < 5e-3, (tag, k, rel)
            gold[f"{tag}/grad_norm/" + k] = p_.grad.norm().item(); gold[f"{tag}/grad_sample/" + k] = MG.sample(p_.grad)
            gold[f"{tag}/grad_absmax/" + k] = p_.grad.abs().max().item()
        gold[f"{tag}/fake"] = fake.detach().numpy(); gold[f"{tag}/upstream"] = gv.numpy()
    # default graph at a size that is not a multiple of 4 (29 x 38): Downsample rounds up (15 x 19, 8 x 10), UpsampleAA doubles
    # (16 x 20, 30 x 38), and the decoder maps are resized onto the skip connections by F.interpolate(..., align_corners=True)
    # (irc:555-556 in both axes, irc:562-563 in the vertical axis only)
    pO = O.seeded_params(O.generator_shapes(), 999, bias_std=0.02)
    irO, _ = O.synthetic_pair(2, 29, 38)
    cfg = R.Config(); cfg.device = "cpu"
    m = R.IRColorizationModel(cfg)
    missing = m.netG.load_state_dict(pO, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("filt") for k in missing.missing_keys), missing
    fake = m(irO)
    assert tuple(fake.shape) == (2, 3, 29, 38)
    gv = torch.randn(fake.shape, generator=torch.Generator().manual_seed(12))
    fake.backward(gv)
    leaves = {k: v.clone().requires_grad_(True) for k, v in pO.items()}
    fo = O.generator_forward(leaves, irO)
    MG.close(fo, fake.detach(), 2e-5, "odd-size forward")
    fo.backward(gv)
    for k, p_ in m.netG.named_parameters():
        noise = k.endswith("bias") and not k.startswith("outc")
        if p_.grad.abs().max() > 1e-4 and not noise:
            rel = ((leaves[k].grad - p_.grad).norm() / p_.grad.norm()).item()
            assert rel < 5e-3, ("odd", k, rel)
        gold["odd/grad_norm/" + k] = p_.grad.norm().item(); gold["odd/grad_sample/" + k] = MG.sample(p_.grad)
        gold["odd/grad_absmax/" + k] = p_.grad.abs().max().item()
    gold["odd/fake"] = fake.detach().numpy(); gold["odd/upstream"] = gv.numpy()
    # norm='none' (irc:158-163, :452-455, :590-593): Identity layers and bias-free convolutions, generator and discriminator
    pN = O.seeded_params(O.generator_shapes(norm="none"), 555, bias_std=0.02)
    cfg = R.Config(); cfg.device = "cpu"; cfg.norm = "none"
    m = R.IRColorizationModel(cfg)
    keys = set(m.netG.state_dict().keys())
    assert "inc.1.bias" not in keys and "resblocks.0.conv_block.1.bias" not in keys and "outc.1.bias" in keys and len(keys) == len(pN) + 4, len(keys)
    missing = m.netG.load_state_dict(pN, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("filt") for k in missing.missing_keys), missing
    fake = m(ir)
    gv = torch.randn(fake.shape, generator=torch.Generator().manual_seed(13))
    fake.backward(gv)
    leaves = {k: v.clone().requires_grad_(True) for k, v in pN.items()}
    fo = O.generator_forward(leaves, ir)
    MG.close(fo, fake.detach(), 2e-5, "norm none forward")
    fo.backward(gv)
    for k, p_ in m.netG.named_parameters():
        rel = ((leaves[k].grad - p_.grad).norm() / p_.grad.norm()).item()
        assert rel < 5e-3, ("nn", k, rel)
        gold["nn/grad_norm/" + k] = p_.grad.norm().item(); gold["nn/grad_sample/" + k] = MG.sample(p_.grad); gold["nn/grad_absmax/" + k] = p_.grad.abs().max().item()
    gold["nn/fake"] = fake.detach().numpy(); gold["nn/upstream"] = gv.numpy()
    pDn = O.seeded_params(O.discriminator_shapes(norm="none"), 556, bias_std=0.02)
    netD = R.NLayerDiscriminator(4, 64, 3, R.get_norm_layer("none"))
    assert set(netD.state_dict().keys()) == set(pDn.keys())
    netD.load_state_dict(pDn)
    xd = torch.cat([ir, rgb], 1).clone().requires_grad_(True)
    pred = netD(xd)
    gp = torch.randn(pred.shape, generator=torch.Generator().manual_seed(14))
    pred.backward(gp)
    leavesD = {k: v.clone().requires_grad_(True) for k, v in pDn.items()}
    xo = torch.cat([ir, rgb], 1).clone().requires_grad_(True)
    po = O.discriminator_forward(leavesD, xo)
    MG.close(po, pred.detach(), 2e-5, "norm none D forward")
    po.backward(gp)
    MG.close(xo.grad, xd.grad, 1e-4, "norm none D input grad")
    for k, p_ in netD.named_parameters():
        assert ((leavesD[k].grad - p_.grad).norm() / p_.grad.norm()).item() < 5e-3, k
        gold["nnD/grad/" + k] = MG.sample(p_.grad, 512)
    gold["nnD/pred"] = pred.detach().numpy(); gold["nnD/upstream"] = gp.numpy(); gold["nnD/dx"] = xd.grad.numpy()
    # norm='batch' (irc:158-159): nn.BatchNorm2d with affine parameters and running statistics, train() and eval() mode
    pB = O.seeded_params(O.generator_shapes(norm="batch"), 333, bias_std=0.05)
    cfg = R.Config(); cfg.device = "cpu"; cfg.norm = "batch"
    m = R.IRColorizationModel(cfg)
    sd = m.netG.state_dict()
    assert "inc.2.running_mean" in sd and "resblocks.8.conv_block.6.num_batches_tracked" in sd and "inc.1.bias" not in sd
    assert set(k for k in sd if not k.endswith(("filt", "running_mean", "running_var", "num_batches_tracked"))) == set(pB.keys())
    missing = m.netG.load_state_dict(pB, strict=False)
    assert not missing.unexpected_keys
    m.train()
    fake = m(ir)
    gv = torch.randn(fake.shape, generator=torch.Generator().manual_seed(15))
    fake.backward(gv)
    state = O.new_bn_state(O.generator_bn_sites())
    leaves = {k: v.clone().requires_grad_(True) for k, v in pB.items()}
    fo = O.generator_forward(leaves, ir, bn_state=state, training=True)
    MG.close(fo, fake.detach(), 2e-5, "batch-norm forward (train)")
    fo.backward(gv)
    for k, p_ in m.netG.named_parameters():
        rel = ((leaves[k].grad - p_.grad).norm() / p_.grad.norm()).item()
        assert rel < 5e-3, ("bn", k, rel)
        gold["bn/grad_norm/" + k] = p_.grad.norm().item(); gold["bn/grad_sample/" + k] = MG.sample(p_.grad); gold["bn/grad_absmax/" + k] = p_.grad.abs().max().item()
    for k in ("inc.2", "down2.1", "resblocks.4.conv_block.6", "up2_conv.1"):
        MG.close(state[k][0], sd[k + ".running_mean"], 1e-5, k + " running_mean"); MG.close(state[k][1], sd[k + ".running_var"], 1e-5, k + " running_var")
        gold["bn/rm/" + k] = sd[k + ".running_mean"].numpy().copy(); gold["bn/rv/" + k] = sd[k + ".running_var"].numpy().copy()
    assert int(sd["inc.2.num_batches_tracked"]) == 1
    gold["bn/fake"] = fake.detach().numpy(); gold["bn/upstream"] = gv.numpy()
    m.eval()
    with torch.no_grad():
        fe = m(ir)
    MG.close(O.generator_forward(pB, ir, bn_state=state, training=False), fe, 2e-5, "batch-norm forward (eval)")
    gold["bn/fake_eval"] = fe.numpy()
    # discriminator with BatchNorm: one call on 2 images (train mode)
    pDb = O.seeded_params(O.discriminator_shapes(norm="batch"), 334, bias_std=0.05)
    netD = R.NLayerDiscriminator(4, 64, 3, R.get_norm_layer("batch"))
    assert set(k for k in netD.state_dict() if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))) == set(pDb.keys())
    netD.load_state_dict(pDb, strict=False)
    netD.train()
    xd = torch.cat([ir, rgb], 1).clone().requires_grad_(True)
    pred = netD(xd)
    gp = torch.randn(pred.shape, generator=torch.Generator().manual_seed(16))
    pred.backward(gp)
    stateD = O.new_bn_state(O.DISCRIMINATOR_BN_SITES)
    leavesD = {k: v.clone().requires_grad_(True) for k, v in pDb.items()}
    xo = torch.cat([ir, rgb], 1).clone().requires_grad_(True)
    po = O.discriminator_forward(leavesD, xo, bn_state=stateD, training=True)
    MG.close(po, pred.detach(), 2e-5, "batch-norm D forward")
    po.backward(gp)
    MG.close(xo.grad, xd.grad, 1e-4, "batch-norm D input grad")
    for k, p_ in netD.named_parameters():
        assert ((leavesD[k].grad - p_.grad).norm() / p_.grad.norm()).item() < 5e-3, k
        gold["bnD/grad/" + k] = MG.sample(p_.grad, 512)
    gold["bnD/pred"] = pred.detach().numpy(); gold["bnD/upstream"] = gp.numpy(); gold["bnD/dx"] = xd.grad.numpy()
    gold["bnD/rm/model.6"] = netD.state_dict()["model.6.running_mean"].numpy().copy(); gold["bnD/rv/model.6"] = netD.state_dict()["model.6.running_var"].numpy().copy()
    # one full D + G iteration with BatchNorm networks (irc:1636-1681): the discriminator normalises the real and the fake batch
    # separately (two calls), the generator runs twice on the same batch, every call updates the running statistics
    pV = O.seeded_params(O.vgg_shapes(), 13, kaiming=True, bias_std=0.05)
    mV = MG.ref_vgg(pV)
    cfgb = R.Config(); cfgb.device = "cpu"; cfgb.norm = "batch"
    mG = R.IRColorizationModel(cfgb); mG.netG.load_state_dict(pB, strict=False); mG.train()
    mD = R.NLayerDiscriminator(4, 64, 3, R.get_norm_layer("batch")); mD.load_state_dict(pDb, strict=False); mD.train()
    optG = torch.optim.Adam(mG.netG.parameters(), lr=cfgb.lr_G, betas=(cfgb.beta1, cfgb.beta2))
    optD = torch.optim.Adam(mD.parameters(), lr=cfgb.lr_D, betas=(cfgb.beta1, cfgb.beta2))
    optD.zero_grad()
    with torch.no_grad():
        fake_det = mG(ir)
    pred_real = mD(torch.cat([ir, rgb], 1)); pred_fake = mD(torch.cat([ir, fake_det], 1))
    loss_D = 0.5 * (torch.relu(1.0 - pred_real).mean() + torch.relu(1.0 + pred_fake).mean())
    loss_D.backward(); optD.step()
    optG.zero_grad()
    fake = mG(ir)
    gan = -mD(torch.cat([ir, fake], 1)).mean()
    l1 = torch.nn.L1Loss()(fake, rgb) * cfgb.lambda_L1
    perc = torch.nn.functional.l1_loss(mV(fake), mV(rgb)) * cfgb.lambda_perc
    tv = R.tv_loss(fake) * cfgb.lambda_tv
    ssim = R.ssim_loss_torch((fake + 1.0) / 2.0, (rgb + 1.0) / 2.0) * cfgb.lambda_ssim
    loss_G = cfgb.lambda_gan * gan + l1 + perc + tv + ssim
    loss_G.backward(); optG.step()
    oG = {k: v.clone() for k, v in pB.items()}; oD = {k: v.clone() for k, v in pDb.items()}
    sG, sD = O.new_bn_state(O.generator_bn_sites()), O.new_bn_state(O.DISCRIMINATOR_BN_SITES)
    losses, _, _ = O.train_step(oG, oD, pV, O.AdamState(oG), O.AdamState(oD), ir, rgb, bn=(sG, sD))
    MG.close(losses["D"], loss_D.detach(), 1e-5, "bn step loss_D")
    for k, r in (("G", loss_G), ("GAN", gan), ("L1", l1), ("perc", perc), ("SSIM", ssim)):
        MG.close(losses[k], r.detach(), 2e-5, "bn step loss " + k)
    sdG, sdD = mG.netG.state_dict(), mD.state_dict()
    assert int(sdG["inc.2.num_batches_tracked"]) == 2 and int(sdD["model.3.num_batches_tracked"]) == 3      # 2 generator, 3 discriminator calls
    for k in ("inc.2", "resblocks.0.conv_block.2", "up1_conv.1"):
        MG.close(sG[k][0], sdG[k + ".running_mean"], 1e-5, k); MG.close(sG[k][1], sdG[k + ".running_var"], 1e-5, k)
        gold["bnstep/rmG/" + k] = sdG[k + ".running_mean"].numpy().copy(); gold["bnstep/rvG/" + k] = sdG[k + ".running_var"].numpy().copy()
    gold.update({"bnstep/loss_D": loss_D.item(), "bnstep/loss_G": loss_G.item(), "bnstep/loss_GAN": gan.item(), "bnstep/loss_L1": l1.item(),
                 "bnstep/loss_perc": perc.item(), "bnstep/loss_SSIM": ssim.item()})
    for k in ("model.3", "model.6", "model.9"):
        MG.close(sD[k][0], sdD[k + ".running_mean"], 1e-5, k); MG.close(sD[k][1], sdD[k + ".running_var"], 1e-5, k)
        gold["bnstep/rmD/" + k] = sdD[k + ".running_mean"].numpy().copy(); gold["bnstep/rvD/" + k] = sdD[k + ".running_var"].numpy().copy()
    for k, p_ in mD.named_parameters():
        gold["bnstep/pD_after/" + k] = MG.sample(p_)
    for k in ("outc.1.weight", "up2_conv.1.weight", "up2_conv.1.bias", "inc.2.weight"):
        gold["bnstep/pG_after/" + k] = MG.sample(dict(mG.netG.named_parameters())[k])
    np.savez_compressed(os.path.join(MG.OUT, "ref_variants.npz"), **gold)
    print("wrote ref_variants.npz keys:", len(gold))


if __name__ == "__main__":
    main()
