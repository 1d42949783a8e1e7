"""CPU oracle for the IR-colorization train / test hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 restatement (torch CPU tensor arithmetic + numpy) of the
algorithm in the reference script ``Code/ir_colorization.py`` (cited below as
``irc:LINE``).  It is *not* part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and only as the checker or the timed CPU baseline.  The product
path (the ``irc_b200`` package) never imports anything from ``oracle/`` and fails
loudly when its CUDA library is missing.

Parity pinning: every function here is checked against the *imported* reference
module by ``oracle/make_golden.py`` (run in the build container, where
``/root/reference`` exists) and against the fixtures that script commits under
``tests/golden/`` (``tests/test_oracle_golden.py``, runs anywhere).  Two things stay
"parity unpinned" because their third-party artefacts are absent from the
reference checkout *and* this image: the ImageNet VGG-16 weights
(torchvision ``vgg16-397923af.pth``; the trunk is exercised with seeded random
weights instead) and scikit-image's ``structural_similarity`` (restated from its
documented defaults in :func:`skimage_ssim`, irc:1208-1215).

Everything is functional: networks take a ``dict`` of tensors keyed exactly like
the reference ``state_dict``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

# ----------------------------------------------------------------------------
# deterministic weight / input recipes (shared by oracle, tests, bench)
# ----------------------------------------------------------------------------

def generator_shapes(input_nc=1, output_nc=3, ngf=64, n_blocks=9, no_antialias_up=False, norm="instance"):
    """Parameter shapes of the generator, irc:457-531 (no_antialias_up: ConvTranspose2d up-sampling, irc:495-516; norm='none':
    get_norm_layer returns a lambda, so use_bias is False and only outc keeps its bias, irc:452-455)."""
    s = {
        "inc.1.weight": (ngf, input_nc, 7, 7), "inc.1.bias": (ngf,),
        "down1.0.weight": (2 * ngf, ngf, 3, 3), "down1.0.bias": (2 * ngf,),
        "down2.0.weight": (4 * ngf, 2 * ngf, 3, 3), "down2.0.bias": (4 * ngf,),
    }
    for b in range(n_blocks):
        for j in (1, 5):
            s[f"resblocks.{b}.conv_block.{j}.weight"] = (4 * ngf, 4 * ngf, 3, 3)
            s[f"resblocks.{b}.conv_block.{j}.bias"] = (4 * ngf,)
    if no_antialias_up:
        s["up1_up.weight"] = (4 * ngf, 4 * ngf, 3, 3); s["up1_up.bias"] = (4 * ngf,)
        s["up2_up.weight"] = (2 * ngf, 2 * ngf, 3, 3); s["up2_up.bias"] = (2 * ngf,)
    s["up1_conv.0.weight"] = (2 * ngf, 6 * ngf, 3, 3); s["up1_conv.0.bias"] = (2 * ngf,)
    s["up2_conv.0.weight"] = (ngf, 3 * ngf, 3, 3); s["up2_conv.0.bias"] = (ngf,)
    s["outc.1.weight"] = (output_nc, ngf, 7, 7); s["outc.1.bias"] = (output_nc,)
    if norm != "instance":
        s = {k: v for k, v in s.items() if not k.endswith(".bias") or k == "outc.1.bias"}
    if norm == "batch":
        # nn.BatchNorm2d(C) behind every convolution but outc: affine weight / bias (running statistics are buffers, not listed)
        for key, c in generator_bn_sites(ngf, n_blocks):
            s[key + ".weight"] = (c,); s[key + ".bias"] = (c,)
    return s


def generator_bn_sites(ngf=64, n_blocks=9):
    """(state_dict prefix, channels) of the norm layers of ResnetUNetGenerator (irc:457-524, :388-412)"""
    sites = [("inc.2", ngf), ("down1.1", 2 * ngf), ("down2.1", 4 * ngf)]
    for b in range(n_blocks):
        sites += [(f"resblocks.{b}.conv_block.2", 4 * ngf), (f"resblocks.{b}.conv_block.6", 4 * ngf)]
    return sites + [("up1_conv.1", 2 * ngf), ("up2_conv.1", ngf)]


DISCRIMINATOR_BN_SITES = [("model.3", 128), ("model.6", 256), ("model.9", 512)]      # irc:611-624 with n_layers=3, ndf=64


def new_bn_state(sites):
    """running_mean = 0, running_var = 1 per norm site (nn.BatchNorm2d defaults)"""
    return {k: (torch.zeros(c), torch.ones(c)) for k, c in sites}


def discriminator_shapes(input_nc=4, ndf=64, norm="instance"):
    """irc:598-630 with n_layers=3 (without InstanceNorm model.2/5/8 are built with bias=False, irc:590-593)."""
    chans = [(input_nc, ndf), (ndf, 2 * ndf), (2 * ndf, 4 * ndf), (4 * ndf, 8 * ndf), (8 * ndf, 1)]
    s = {}
    for idx, (ci, co) in zip((0, 2, 5, 8, 11), chans):
        s[f"model.{idx}.weight"] = (co, ci, 4, 4)
        if norm == "instance" or idx in (0, 11):
            s[f"model.{idx}.bias"] = (co,)
    if norm == "batch":
        for key, c in DISCRIMINATOR_BN_SITES:
            s[key + ".weight"] = (c,); s[key + ".bias"] = (c,)
    return s


VGG_CFG = [(0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256), (14, 256, 256)]
VGG_POOL_AFTER = (2, 7)  # max-pool follows features.2 (conv1_2) and features.7 (conv2_2); irc:664


def vgg_shapes():
    s = {}
    for idx, ci, co in VGG_CFG:
        s[f"features.{idx}.weight"] = (co, ci, 3, 3)
        s[f"features.{idx}.bias"] = (co,)
    return s


def seeded_params(shapes, seed, std=0.02, bias_std=0.0, kaiming=False) -> Params:
    """Deterministic parameters: one torch.Generator, tensors drawn in sorted-key order.

    Weights ~ N(0, std) (irc:181) — or He fan-out for the stand-in VGG trunk —
    biases 0 (irc:191) unless bias_std > 0 (tests use non-zero biases so the bias
    path is exercised)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(shapes):
        shp = shapes[k]
        if k.endswith("weight") and len(shp) == 1:
            out[k] = 1.0 + torch.randn(shp, generator=g) * 0.1          # BatchNorm scale: N(1, .) (irc:191-193 uses N(1, 0.02))
        elif k.endswith("weight"):
            sd = math.sqrt(2.0 / (shp[0] * shp[2] * shp[3])) if kaiming else std
            out[k] = torch.randn(shp, generator=g) * sd
        else:
            out[k] = torch.randn(shp, generator=g) * bias_std
    return out


def synthetic_pair(B, H, W, rank=0) -> Tuple[torch.Tensor, torch.Tensor]:
    """SURVEY.md §8d recipe: ir, rgb ~ U(-1, 1)."""
    g = torch.Generator().manual_seed(7 + rank)
    ir = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    rgb = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    return ir, rgb


# ----------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------

def lr_factor(epoch0: int, decay_start: int = 40, epochs: int = 50) -> float:
    """irc:219-231: LambdaLR factor for the 0-based scheduler epoch."""
    e = epoch0 + 1
    if e <= decay_start:
        return 1.0
    if e >= epochs:
        return 0.0
    return max(0.0, 1.0 - (e - decay_start) / float(max(1, epochs - decay_start)))


def binomial3() -> torch.Tensor:
    """irc:251, :264-265: [1,2,1] outer product / 16."""
    a = torch.tensor([1.0, 2.0, 1.0])
    return (a[:, None] * a[None, :]) / 16.0


def _reflect(i: torch.Tensor, n: int) -> torch.Tensor:
    """reflect-without-edge-repeat index map (-1 -> 1, n -> n-2)."""
    i = i.abs()
    return torch.where(i > n - 1, 2 * (n - 1) - i, i)


def blur_down(x: torch.Tensor) -> torch.Tensor:
    """Downsample, irc:307-310: reflect-pad 1, depthwise [1,2,1]x[1,2,1]/16, stride 2.

    Written as a gather over explicit reflected indices."""
    B, C, H, W = x.shape
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    w = torch.tensor([1.0, 2.0, 1.0], dtype=x.dtype) / 4.0
    oy = torch.arange(Ho); ox = torch.arange(Wo)
    out = torch.zeros(B, C, Ho, Wo, dtype=x.dtype)
    for a in range(3):
        ry = _reflect(2 * oy - 1 + a, H)
        rows = x[:, :, ry, :]
        for b in range(3):
            rx = _reflect(2 * ox - 1 + b, W)
            out = out + (w[a] * w[b]) * rows[:, :, :, rx]
    return out


def bilinear_up2(x: torch.Tensor) -> torch.Tensor:
    """F.interpolate(scale_factor=2, bilinear, align_corners=True), irc:351-352, as explicit lerps."""
    B, C, H, W = x.shape

    def axis(n):
        o = torch.arange(2 * n, dtype=torch.float64)
        s = o * (n - 1) / (2 * n - 1) if n > 1 else torch.zeros_like(o)
        i0 = s.floor().clamp(max=n - 1).long()
        i1 = (i0 + 1).clamp(max=n - 1)
        f = (s - i0).to(x.dtype)
        return i0, i1, f

    y0, y1, fy = axis(H)
    x0, x1, fx = axis(W)
    top = x[:, :, y0, :]; bot = x[:, :, y1, :]
    v = top + (bot - top) * fy[None, None, :, None]
    l = v[:, :, :, x0]; r = v[:, :, :, x1]
    return l + (r - l) * fx[None, None, None, :]


def blur_same(u: torch.Tensor) -> torch.Tensor:
    """reflect-pad 1 + depthwise binomial 3x3, stride 1 (irc:353-354)."""
    B, C, H, W = u.shape
    w = torch.tensor([1.0, 2.0, 1.0], dtype=u.dtype) / 4.0
    oy = torch.arange(H); ox = torch.arange(W)
    out = torch.zeros_like(u)
    for a in range(3):
        rows = u[:, :, _reflect(oy - 1 + a, H), :]
        for b in range(3):
            out = out + (w[a] * w[b]) * rows[:, :, :, _reflect(ox - 1 + b, W)]
    return out


def upsample_aa(x: torch.Tensor) -> torch.Tensor:
    """UpsampleAA, irc:350-355."""
    return blur_same(bilinear_up2(x))


def instance_norm(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.InstanceNorm2d defaults (irc:161): per-(n,c) biased variance, no affine."""
    m = x.mean(dim=(2, 3), keepdim=True)
    v = ((x - m) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - m) / torch.sqrt(v + eps)


def _norm_at(p: Params, conv_bias_key: str, bn_key: str, bn_state, training: bool):
    """the layer get_norm_layer(cfg.norm) puts behind a convolution, recognised from the parameters (irc:148-165): BatchNorm2d when
    its affine weight exists, InstanceNorm2d when the convolution kept its bias, Identity otherwise"""
    if bn_key + ".weight" in p:
        rm, rv = bn_state[bn_key]
        return lambda t: F.batch_norm(t, rm, rv, p[bn_key + ".weight"], p[bn_key + ".bias"], training, 0.1, 1e-5)
    return instance_norm if conv_bias_key in p else (lambda t: t)


def _conv(x, w, b, stride=1, pad=0, reflect=0):
    if reflect:
        x = F.pad(x, (reflect,) * 4, mode="reflect")
    return F.conv2d(x, w, b, stride=stride, padding=pad)


def resnet_block(p: Params, prefix: str, x: torch.Tensor, nrm1=instance_norm, nrm2=None) -> torch.Tensor:
    """irc:417-418 with reflect padding; nrm1 / nrm2 = the two norm layers (instance norm by default)."""
    h = _conv(x, p[prefix + "conv_block.1.weight"], p.get(prefix + "conv_block.1.bias"), reflect=1)
    h = torch.relu(nrm1(h))
    h = _conv(h, p[prefix + "conv_block.5.weight"], p.get(prefix + "conv_block.5.bias"), reflect=1)
    return x + (nrm2 or nrm1)(h)


def generator_forward(p: Params, x: torch.Tensor, n_blocks: int = 9, taps: Optional[dict] = None, no_antialias: bool = False,
                      bn_state: Optional[dict] = None, training: bool = True) -> torch.Tensor:
    """ResnetUNetGenerator.forward (irc:540-569); returns the image only.  Default config; ConvTranspose2d up-sampling when the
    parameters carry up{1,2}_up.weight (no_antialias_up=True, irc:495-516); stride-2 down-sampling convolutions without the blur
    modules when no_antialias=True (irc:468, :474, :482).

    ``taps`` (optional dict) receives named intermediate activations for per-layer parity."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    # the norm layer shows in the parameters (irc:148-165): BatchNorm2d affine weights (bn_state: running statistics, updated in
    # place when training), a convolution bias in front of InstanceNorm2d, or neither (norm='none')
    N = lambda conv, bn: _norm_at(p, conv + ".bias", bn, bn_state, training)
    x0 = tap("x0", torch.relu(N("inc.1", "inc.2")(_conv(x, p["inc.1.weight"], p.get("inc.1.bias"), reflect=3))))
    sd = 2 if no_antialias else 1
    d1 = tap("down1", torch.relu(N("down1.0", "down1.1")(_conv(x0, p["down1.0.weight"], p.get("down1.0.bias"), stride=sd, pad=1))))
    x1 = tap("x1", d1 if no_antialias else blur_down(d1))
    d2 = tap("down2", torch.relu(N("down2.0", "down2.1")(_conv(x1, p["down2.0.weight"], p.get("down2.0.bias"), stride=sd, pad=1))))
    x2 = tap("x2", d2 if no_antialias else blur_down(d2))
    h = x2
    for b in range(n_blocks):
        pre = f"resblocks.{b}.conv_block."
        h = tap(f"res{b}", resnet_block(p, f"resblocks.{b}.", h, N(pre + "1", pre + "2"), N(pre + "5", pre + "6")))
    convT = "up1_up.weight" in p          # no_antialias_up=True: nn.ConvTranspose2d(C, C, 3, 2, 1, 1) instead of UpsampleAA (irc:495-499)
    up = (lambda t, k: F.conv_transpose2d(t, p[k + ".weight"], p.get(k + ".bias"), stride=2, padding=1, output_padding=1)) if convT \
        else (lambda t, k: upsample_aa(t))
    y = tap("up1_up", up(h, "up1_up"))
    if y.shape[-2:] != x1.shape[-2:]:  # irc:555-556
        y = F.interpolate(y, size=x1.shape[-2:], mode="bilinear", align_corners=True)
    y = torch.cat([y, x1], dim=1)
    y = tap("up1", torch.relu(N("up1_conv.0", "up1_conv.1")(_conv(y, p["up1_conv.0.weight"], p.get("up1_conv.0.bias"), pad=1))))
    y = tap("up2_up", up(y, "up2_up"))
    if y.shape[-2:] != x0.shape[-2:]:  # irc:562-563
        y = F.interpolate(y, size=x0.shape[-2:], mode="bilinear", align_corners=True)
    y = torch.cat([y, x0], dim=1)
    y = tap("up2", torch.relu(N("up2_conv.0", "up2_conv.1")(_conv(y, p["up2_conv.0.weight"], p.get("up2_conv.0.bias"), pad=1))))
    return tap("out", torch.tanh(_conv(y, p["outc.1.weight"], p["outc.1.bias"], reflect=3)))


def discriminator_forward(p: Params, x: torch.Tensor, taps: Optional[dict] = None, bn_state: Optional[dict] = None, training: bool = True) -> torch.Tensor:
    """NLayerDiscriminator.forward, irc:598-635 (n_layers=3, instance norm)."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t
    N = lambda conv, bn: _norm_at(p, conv + ".bias", bn, bn_state, training)      # BatchNorm2d / InstanceNorm2d / Identity (irc:590-593)
    h = tap("d0", F.leaky_relu(_conv(x, p["model.0.weight"], p["model.0.bias"], stride=2, pad=1), 0.2))
    h = tap("d2", F.leaky_relu(N("model.2", "model.3")(_conv(h, p["model.2.weight"], p.get("model.2.bias"), stride=2, pad=1)), 0.2))
    h = tap("d5", F.leaky_relu(N("model.5", "model.6")(_conv(h, p["model.5.weight"], p.get("model.5.bias"), stride=2, pad=1)), 0.2))
    h = tap("d8", F.leaky_relu(N("model.8", "model.9")(_conv(h, p["model.8.weight"], p.get("model.8.bias"), stride=1, pad=1)), 0.2))
    return tap("d11", _conv(h, p["model.11.weight"], p["model.11.bias"], stride=1, pad=1))


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def vgg_forward(p: Params, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """VGGPerceptual.forward, irc:677-683: [-1,1] -> [0,1] -> ImageNet norm -> features[:16]."""
    mean = torch.tensor(IMAGENET_MEAN, dtype=x.dtype).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=x.dtype).view(1, 3, 1, 1)
    h = ((x + 1.0) / 2.0 - mean) / std
    for idx, _, _ in VGG_CFG:
        h = torch.relu(_conv(h, p[f"features.{idx}.weight"], p[f"features.{idx}.bias"], pad=1))
        if idx in VGG_POOL_AFTER:
            h = F.max_pool2d(h, 2, 2)
        if taps is not None:
            taps[f"vgg{idx}"] = h
    return h


def tv_loss(x: torch.Tensor) -> torch.Tensor:
    """irc:692-694: mean |d/dy| + mean |d/dx| with separate denominators."""
    B, C, H, W = x.shape
    dv = (x[:, :, 1:, :] - x[:, :, :-1, :]).abs().sum() / (B * C * (H - 1) * W)
    dh = (x[:, :, :, 1:] - x[:, :, :, :-1]).abs().sum() / (B * C * H * (W - 1))
    return dv + dh


def gaussian_taps(n: int = 11, sigma: float = 1.5, dtype=torch.float32) -> torch.Tensor:
    """irc:699-703."""
    c = torch.arange(n, dtype=dtype) - (n - 1) / 2.0
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _gauss_sep(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """zero-padded separable window (row pass then column pass); equals the 2-D window of irc:706-711."""
    n = g.numel(); C = x.shape[1]
    kx = g.view(1, 1, 1, n).expand(C, 1, 1, n)
    ky = g.view(1, 1, n, 1).expand(C, 1, n, 1)
    x = F.conv2d(x, kx, padding=(0, n // 2), groups=C)
    return F.conv2d(x, ky, padding=(n // 2, 0), groups=C)


def ssim_loss(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """ssim_loss_torch, irc:714-750: 1 - mean(ssim_map) on [0,1] images."""
    g = gaussian_taps(window_size, 1.5, img1.dtype)
    mu1 = _gauss_sep(img1, g); mu2 = _gauss_sep(img2, g)
    s11 = _gauss_sep(img1 * img1, g) - mu1 * mu1
    s22 = _gauss_sep(img2 * img2, g) - mu2 * mu2
    s12 = _gauss_sep(img1 * img2, g) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s11 + s22 + C2))
    return 1.0 - (m.mean() if size_average else m.mean(dim=(1, 2, 3)))


# ----------------------------------------------------------------------------
# train step (irc:1636-1681) with a restated Adam (torch.optim.Adam defaults)
# ----------------------------------------------------------------------------

LAMBDAS = dict(L1=30.0, perc=30.0, tv=1e-4, ssim=2.0, gan=0.1)  # irc:100-104


class AdamState:
    """Bias-corrected Adam, eps 1e-8, no weight decay/amsgrad (irc:1601-1604 defaults)."""

    def __init__(self, params: Params, lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta1, beta2, eps, 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, params: Params, grads: Params, lr_scale: float = 1.0) -> None:
        self.t += 1
        bc1 = 1.0 - self.b1 ** self.t
        bc2 = 1.0 - self.b2 ** self.t
        for k, p in params.items():
            g = grads[k]
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-(self.lr * lr_scale) / bc1)


def d_loss_and_grads(pD: Params, ir, rgb, fake_detached, bn_state=None):
    """irc:1639-1650 (bn_state: running statistics of a BatchNorm discriminator, updated by each of the two calls)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in pD.items()}
    pred_real = discriminator_forward(leaves, torch.cat([ir, rgb], 1), bn_state=bn_state)
    pred_fake = discriminator_forward(leaves, torch.cat([ir, fake_detached], 1), bn_state=bn_state)
    loss = 0.5 * (torch.relu(1.0 - pred_real).mean() + torch.relu(1.0 + pred_fake).mean())
    grads = torch.autograd.grad(loss, list(leaves.values()))
    return loss.detach(), dict(zip(leaves.keys(), grads))


def g_loss_and_grads(pG: Params, pD: Params, pV: Params, ir, rgb, lambdas=LAMBDAS, want_fake_grad=False, bn=(None, None)):
    """irc:1657-1680. Returns (loss dict, grads wrt G params[, dL/dfake]).  bn = (generator, discriminator) BatchNorm states."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in pG.items()}
    fake = generator_forward(leaves, ir, bn_state=bn[0])
    if want_fake_grad:
        fake.retain_grad()
    gan = -discriminator_forward(pD, torch.cat([ir, fake], 1), bn_state=bn[1]).mean()
    l1 = (fake - rgb).abs().mean() * lambdas["L1"]
    perc = (vgg_forward(pV, fake) - vgg_forward(pV, rgb)).abs().mean() * lambdas["perc"]
    tv = tv_loss(fake) * lambdas["tv"]
    ssim = ssim_loss((fake + 1.0) / 2.0, (rgb + 1.0) / 2.0) * lambdas["ssim"]
    total = lambdas["gan"] * gan + l1 + perc + tv + ssim
    total.backward()
    grads = {k: v.grad.detach() for k, v in leaves.items()}
    losses = dict(G=total.detach(), GAN=gan.detach(), L1=l1.detach(), perc=perc.detach(), TV=tv.detach(), SSIM=ssim.detach())
    if want_fake_grad:
        return losses, grads, fake.grad.detach(), fake.detach()
    return losses, grads


def train_step(pG: Params, pD: Params, pV: Params, optG: AdamState, optD: AdamState, ir, rgb,
               lr_scale: float = 1.0, lambdas=LAMBDAS, bn=(None, None)):
    """One iteration of the hot loop, irc:1636-1681, mutating pG/pD and the Adam states in place.  With norm='batch' the
    running statistics in bn = (generator state, discriminator state) see two generator and three discriminator forward calls."""
    with torch.no_grad():
        fake_d = generator_forward(pG, ir, bn_state=bn[0])
    loss_D, gD = d_loss_and_grads(pD, ir, rgb, fake_d, bn_state=bn[1])
    with torch.no_grad():
        optD.step(pD, gD, lr_scale)
    losses, gG = g_loss_and_grads(pG, pD, pV, ir, rgb, lambdas, bn=bn)
    with torch.no_grad():
        optG.step(pG, gG, lr_scale)
    losses["D"] = loss_D
    return losses, gG, gD


# ----------------------------------------------------------------------------
# test-mode core (irc:865-876, irc:1184-1217)
# ----------------------------------------------------------------------------

def quantize_u8(fake_chw: torch.Tensor) -> np.ndarray:
    """tensor_to_rgb_image on one CHW image: clip((x+1)/2,0,1)*255 truncated to uint8, HWC."""
    x = fake_chw.detach().cpu().numpy().astype(np.float32)
    x = (x + 1.0) / 2.0
    x = np.clip(x, 0.0, 1.0)
    return np.transpose((x * 255.0).astype(np.uint8), (1, 2, 0))


def compute_metrics(pred_01: np.ndarray, gt_01: np.ndarray):
    """irc:1197-1205: MAE, MSE, PSNR (peak 1.0, +1e-12, inf when mse == 0)."""
    d = pred_01 - gt_01
    mae = float(np.mean(np.abs(d)))
    mse = float(np.mean(d ** 2))
    psnr = float("inf") if mse == 0 else -10.0 * math.log10(mse + 1e-12)
    return mae, mse, psnr


def skimage_ssim(gt_01: np.ndarray, pred_01: np.ndarray) -> float:
    """Restatement of skimage.metrics.structural_similarity(gt, pred, data_range=1.0,
    channel_axis=2) with its documented defaults (irc:1210): 7x7 uniform window, sample
    covariance, K1=.01, K2=.03, 3-pixel border crop, mean over channels.  PARITY UNPINNED:
    scikit-image is not installed in this image."""
    from scipy.ndimage import uniform_filter
    win, K1, K2, R = 7, 0.01, 0.03, 1.0
    NP = win * win
    cov_norm = NP / (NP - 1.0)
    C1, C2 = (K1 * R) ** 2, (K2 * R) ** 2
    vals = []
    for c in range(gt_01.shape[2]):
        x = gt_01[:, :, c].astype(np.float64); y = pred_01[:, :, c].astype(np.float64)
        ux = uniform_filter(x, size=win); uy = uniform_filter(y, size=win)
        uxx = uniform_filter(x * x, size=win); uyy = uniform_filter(y * y, size=win); uxy = uniform_filter(x * y, size=win)
        vx = cov_norm * (uxx - ux * ux); vy = cov_norm * (uyy - uy * uy); vxy = cov_norm * (uxy - ux * uy)
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        pad = (win - 1) // 2
        vals.append(S[pad:-pad, pad:-pad].mean())
    return float(np.mean(vals))
