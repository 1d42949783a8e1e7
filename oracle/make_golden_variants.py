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
    np.savez_compressed(os.path.join(MG.OUT, "ref_variants.npz"), **gold)
    print("wrote ref_variants.npz keys:", len(gold))


if __name__ == "__main__":
    main()
