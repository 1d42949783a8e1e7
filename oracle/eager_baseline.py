"""The reference's train step as the SAME sequence of PyTorch library calls the reference makes (irc = Code/ir_colorization.py),
device-agnostic, for timing "PyTorch eager + cuDNN on the same B200" next to the hand-written kernels (SURVEY.md §2.1,
§8d: the stated secondary bar).  BASELINE / TEST INFRASTRUCTURE ONLY - the product never imports it.

irc_oracle.py restates the stencils as explicit index gathers (good for pinning the arithmetic, slow on a GPU); this file
keeps the reference's own operator choices instead - F.pad(reflect) + grouped F.conv2d for Downsample / UpsampleAA
(irc:292-310, :340-355), F.interpolate(bilinear, align_corners=True) (irc:351), F.instance_norm (irc:161), a 2-D 11 x 11
grouped-conv SSIM (irc:706-750), torch.optim.Adam (irc:1601-1604) - so that what is timed is what the unmodified
reference would launch with cfg.device = "cuda".  tests/test_oracle_golden.py checks it against irc_oracle on the CPU."""
from __future__ import annotations

import torch
import torch.nn.functional as F

LAMBDAS = dict(L1=30.0, perc=30.0, tv=1e-4, ssim=2.0, gan=0.1)   # irc:100-104
VGG_IDX = (0, 2, 5, 7, 10, 12, 14)
VGG_POOL_AFTER = (2, 7)


def _filt(C, ref):
    a = torch.tensor([1.0, 2.0, 1.0], device=ref.device, dtype=ref.dtype)
    f = a[:, None] * a[None, :]
    return (f / f.sum())[None, None].repeat(C, 1, 1, 1)


def downsample(x):
    """irc:307-310"""
    C = x.shape[1]
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), _filt(C, x), stride=2, groups=C)


def upsample_aa(x):
    """irc:350-355"""
    C = x.shape[1]
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), _filt(C, x), stride=1, groups=C)


def _cna(x, w, b, pad=0, reflect=0, act=True):
    if reflect:
        x = F.pad(x, (reflect,) * 4, mode="reflect")
    x = F.instance_norm(F.conv2d(x, w, b, padding=pad))
    return torch.relu(x) if act else x


def generator(p, x, n_blocks=9):
    """irc:533-569 (default graph)"""
    x0 = _cna(x, p["inc.1.weight"], p["inc.1.bias"], reflect=3)
    x1 = downsample(_cna(x0, p["down1.0.weight"], p["down1.0.bias"], pad=1))
    x2 = downsample(_cna(x1, p["down2.0.weight"], p["down2.0.bias"], pad=1))
    h = x2
    for b in range(n_blocks):
        y = _cna(h, p[f"resblocks.{b}.conv_block.1.weight"], p[f"resblocks.{b}.conv_block.1.bias"], reflect=1)
        h = h + _cna(y, p[f"resblocks.{b}.conv_block.5.weight"], p[f"resblocks.{b}.conv_block.5.bias"], reflect=1, act=False)
    y = upsample_aa(h)
    y = _cna(torch.cat([y, x1], 1), p["up1_conv.0.weight"], p["up1_conv.0.bias"], pad=1)
    y = upsample_aa(y)
    y = _cna(torch.cat([y, x0], 1), p["up2_conv.0.weight"], p["up2_conv.0.bias"], pad=1)
    return torch.tanh(F.conv2d(F.pad(y, (3, 3, 3, 3), mode="reflect"), p["outc.1.weight"], p["outc.1.bias"]))


def discriminator(p, x):
    """irc:598-635"""
    h = F.leaky_relu(F.conv2d(x, p["model.0.weight"], p["model.0.bias"], stride=2, padding=1), 0.2)
    for i, s in ((2, 2), (5, 2), (8, 1)):
        h = F.leaky_relu(F.instance_norm(F.conv2d(h, p[f"model.{i}.weight"], p[f"model.{i}.bias"], stride=s, padding=1)), 0.2)
    return F.conv2d(h, p["model.11.weight"], p["model.11.bias"], stride=1, padding=1)


def vgg(p, x):
    """irc:677-683"""
    mean = torch.tensor((0.485, 0.456, 0.406), device=x.device, dtype=x.dtype).view(1, 3, 1, 1)
    std = torch.tensor((0.229, 0.224, 0.225), device=x.device, dtype=x.dtype).view(1, 3, 1, 1)
    h = ((x + 1.0) / 2.0 - mean) / std
    for i in VGG_IDX:
        h = torch.relu(F.conv2d(h, p[f"features.{i}.weight"], p[f"features.{i}.bias"], padding=1))
        if i in VGG_POOL_AFTER:
            h = F.max_pool2d(h, 2, 2)
    return h


def tv_loss(x):
    """irc:686-694"""
    return (x[:, :, 1:, :] - x[:, :, :-1, :]).abs().mean() + (x[:, :, :, 1:] - x[:, :, :, :-1]).abs().mean()


def ssim_loss(a, b, n=11, sigma=1.5):
    """irc:699-750: the 2-D window is rebuilt on every call, five grouped convolutions"""
    c = torch.arange(n, device=a.device, dtype=a.dtype) - (n - 1) / 2.0
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2)); g = g / g.sum()
    C = a.shape[1]
    win = (g[:, None] @ g[None, :])[None, None].expand(C, 1, n, n).contiguous()
    f = lambda t: F.conv2d(t, win, padding=n // 2, groups=C)
    mu1, mu2 = f(a), f(b)
    s11, s22, s12 = f(a * a) - mu1 * mu1, f(b * b) - mu2 * mu2, f(a * b) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s11 + s22 + C2))
    return 1.0 - m.mean()


class EagerTrainer:
    """leaf parameters + torch.optim.Adam; `step` is the loop body irc:1636-1681 as written (two generator forwards,
    discriminator gradients also accumulated in the G step)"""

    def __init__(self, pG, pD, pV, device, lr=2e-4, betas=(0.5, 0.999), lambdas=LAMBDAS, channels_last=False, autocast=None):
        mv = lambda d, rg: {k: v.detach().to(device).clone().requires_grad_(rg) for k, v in d.items()}
        self.G, self.D, self.V = mv(pG, True), mv(pD, True), mv(pV, False)
        if channels_last:
            for d in (self.G, self.D, self.V):
                for k, v in d.items():
                    if v.dim() == 4:
                        v.data = v.data.contiguous(memory_format=torch.channels_last)
        self.optG = torch.optim.Adam(list(self.G.values()), lr=lr, betas=betas)
        self.optD = torch.optim.Adam(list(self.D.values()), lr=lr, betas=betas)
        self.lam, self.cl, self.autocast, self.device = lambdas, channels_last, autocast, torch.device(device)

    def _ctx(self):
        if self.autocast is None:
            import contextlib
            return contextlib.nullcontext()
        return torch.autocast(self.device.type, dtype=self.autocast)

    def step(self, ir, rgb):
        lam = self.lam
        if self.cl:
            ir = ir.contiguous(memory_format=torch.channels_last); rgb = rgb.contiguous(memory_format=torch.channels_last)
        self.optD.zero_grad()
        with self._ctx():
            with torch.no_grad():
                fake_d = generator(self.G, ir)
            pred_real = discriminator(self.D, torch.cat([ir, rgb], 1))
            pred_fake = discriminator(self.D, torch.cat([ir, fake_d.detach()], 1))
            loss_D = 0.5 * (torch.relu(1.0 - pred_real.float()).mean() + torch.relu(1.0 + pred_fake.float()).mean())
        loss_D.backward()
        self.optD.step()
        self.optG.zero_grad()
        with self._ctx():
            fake = generator(self.G, ir)
            gan = -discriminator(self.D, torch.cat([ir, fake], 1)).float().mean()
            l1 = F.l1_loss(fake.float(), rgb) * lam["L1"]
            perc = F.l1_loss(vgg(self.V, fake).float(), vgg(self.V, rgb).float()) * lam["perc"]
        ff = fake.float()
        tv = tv_loss(ff) * lam["tv"]
        ssim = ssim_loss((ff + 1.0) / 2.0, (rgb + 1.0) / 2.0) * lam["ssim"]
        loss_G = lam["gan"] * gan + l1 + perc + tv + ssim
        loss_G.backward()
        self.optG.step()
        return dict(D=loss_D.detach(), G=loss_G.detach(), GAN=gan.detach(), L1=l1.detach(), perc=perc.detach(), TV=tv.detach(), SSIM=ssim.detach())
