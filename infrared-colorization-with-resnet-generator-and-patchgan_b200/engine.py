"""Execution plans of the hot path: generator, discriminator and VGG trunk forward/backward as
sequences of libirc_sm100 launches over pre-allocated NHWC bf16 frames.

The engines own every buffer (nothing is allocated per step, so a step can be captured in a
CUDA graph) and take a *backend* object that exposes one method per C-ABI entry point
(`_native.CudaBackend` in the product).  Reference: irc = Code/ir_colorization.py."""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import layout as L
from ._native import IDENTITY, Tables, View

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
EPS = 1e-5   # nn.InstanceNorm2d default (irc:161)

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # irc:667-668
IMAGENET_STD = (0.229, 0.224, 0.225)


def generator_shapes(input_nc=1, output_nc=3, ngf=64, n_blocks=9, no_antialias_up=False, norm="instance") -> Dict[str, tuple]:
    """state_dict parameter shapes of ResnetUNetGenerator (irc:457-531).  no_antialias_up: the two UpsampleAA modules (buffers
    only) become nn.ConvTranspose2d(C, C, 3, 2, 1, 1) with parameters up{1,2}_up.{weight (Cin, Cout, 3, 3), bias} (irc:495-516).
    norm='none': get_norm_layer returns a lambda, so use_bias is False (irc:452-455) and every convolution but outc loses its bias"""
    s = {"inc.1.weight": (ngf, input_nc, 7, 7), "inc.1.bias": (ngf,),
         "down1.0.weight": (2 * ngf, ngf, 3, 3), "down1.0.bias": (2 * ngf,),
         "down2.0.weight": (4 * ngf, 2 * ngf, 3, 3), "down2.0.bias": (4 * ngf,)}
    for b in range(n_blocks):
        for j in (1, 5):
            s[f"resblocks.{b}.conv_block.{j}.weight"] = (4 * ngf, 4 * ngf, 3, 3)
            s[f"resblocks.{b}.conv_block.{j}.bias"] = (4 * ngf,)
    if no_antialias_up:
        s["up1_up.weight"] = (4 * ngf, 4 * ngf, 3, 3); s["up1_up.bias"] = (4 * ngf,)
    s["up1_conv.0.weight"] = (2 * ngf, 6 * ngf, 3, 3); s["up1_conv.0.bias"] = (2 * ngf,)
    if no_antialias_up:
        s["up2_up.weight"] = (2 * ngf, 2 * ngf, 3, 3); s["up2_up.bias"] = (2 * ngf,)
    s["up2_conv.0.weight"] = (ngf, 3 * ngf, 3, 3); s["up2_conv.0.bias"] = (ngf,)
    s["outc.1.weight"] = (output_nc, ngf, 7, 7); s["outc.1.bias"] = (output_nc,)
    if norm != "instance":
        s = {k: v for k, v in s.items() if not k.endswith(".bias") or k == "outc.1.bias"}
    if norm == "batch":
        for key, c in generator_bn_sites(ngf, n_blocks):      # nn.BatchNorm2d affine parameters (running statistics are buffers)
            s[key + ".weight"] = (c,); s[key + ".bias"] = (c,)
    return s


def generator_bn_sites(ngf=64, n_blocks=9) -> List[tuple]:
    """(state_dict prefix, channels) of the norm layers of ResnetUNetGenerator in network order (irc:457-524, :388-412)"""
    sites = [("inc.2", ngf), ("down1.1", 2 * ngf), ("down2.1", 4 * ngf)]
    for b in range(n_blocks):
        sites += [(f"resblocks.{b}.conv_block.2", 4 * ngf), (f"resblocks.{b}.conv_block.6", 4 * ngf)]
    return sites + [("up1_conv.1", 2 * ngf), ("up2_conv.1", ngf)]


DISCRIMINATOR_BN_SITES = [("model.3", 128), ("model.6", 256), ("model.9", 512)]      # irc:611-624 (n_layers 3, ndf 64)


def new_bn_state(sites, device) -> Dict[str, tuple]:
    """running_mean = 0, running_var = 1 per norm site (nn.BatchNorm2d defaults)"""
    return {k: (torch.zeros(c, device=device), torch.ones(c, device=device)) for k, c in sites}


def discriminator_shapes(input_nc=4, ndf=64, norm="instance") -> Dict[str, tuple]:
    """NLayerDiscriminator(n_layers=3) parameter shapes (irc:598-630); without InstanceNorm model.2/5/8 have no bias (irc:590-593)"""
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


def vgg_shapes() -> Dict[str, tuple]:
    s = {}
    for idx, ci, co in VGG_CFG:
        s[f"features.{idx}.weight"] = (co, ci, 3, 3); s[f"features.{idx}.bias"] = (co,)
    return s


def _splits_for(tiles: int, kb_total: int, sms: int = 148) -> int:
    """split-K factor of a weight-gradient GEMM: fill exactly one wave of CTAs (measured: a second, partial wave
    costs more than the longer main loop of a single one — scripts/prof_tn.py)"""
    s = max(1, sms // max(tiles, 1))
    return int(max(1, min(s, kb_total, sms)))


class ConvOp:
    """One convolution in its three GEMM roles over frame buffers."""

    def __init__(self, be, lay: L.WeightLayout, taps: List[int], arena: L.ParamArena, max_rows: int, bias_name: Optional[str] = None,
                 pixels: int = 0, name: str = ""):
        self.be, self.lay, self.taps, self.arena = be, lay, list(taps), arena
        self.bias_name = bias_name
        self.name = name
        # algorithmic work of one pass in any role: 2 * (true weight count) * (forward output pixels)  
        self.flops = 2.0 * lay.param_numel * pixels
        N, T, K = lay.N, lay.T, lay.K
        self.splits = _splits_for(be.tn_gemm_ctas(N, K, T), (max_rows + 63) // 64)
        self.partial = torch.zeros(self.splits * N * T * K, device=arena.device)
        self.grad_flat = arena.grad[lay.param_offset:lay.param_offset + lay.param_numel]

    def bias(self):
        return self.arena.view(self.bias_name) if self.bias_name else None

    def fwd_stats(self, a, a_chan_off, out, stats, row_img, n_img, rows_per_img, z_view, C, H, W, **kw):
        """forward convolution followed by the InstanceNorm statistics of its output (irc:161): from the GEMM epilogue when the
        layer is deep enough for the epilogue to hide behind the MMAs (be.stats_epilogue_min_k) and the frame geometry allows
        it, else by a separate pass over the output"""
        if stats is None:
            # norm='none' (irc:158-163): no statistics, the activation is applied by the GEMM epilogue instead (kw carries it)
            self.fwd(a, a_chan_off, out, **kw)
            return
        # a 64-channel 3 x 3 layer with >= 128 input channels runs on the packed-taps GEMM (a third of the MMA instructions), which
        # has no statistics epilogue: the separate pass over its narrow output costs less than the packing saves (up2: -45 us)
        packed = self.lay.N == 64 and self.lay.T == 9 and self.lay.K >= 128 and getattr(self.be, "conv_reuse", 0) in (-1, 2)
        if not packed and self.lay.T * self.lay.K >= self.be.stats_epilogue_min_k and rows_per_img >= 128 and self.lay.N % 64 == 0 and row_img is not None:
            self.fwd(a, a_chan_off, out, row_img=row_img, in_stats=(stats, n_img, rows_per_img), **kw)
        else:
            self.fwd(a, a_chan_off, out, **kw)
            self.be.in_stats(z_view, C, n_img, H, W, stats)

    def fwd(self, a, a_chan_off, out, **kw):
        n = self.lay.w_f.rows
        if n > 512:
            # the GEMM takes at most 512 output columns per launch: wide outputs (transposed conv: 4 phases x Cout) go in blocks
            assert n % 512 == 0 and not kw.get("in_stats") and not kw.get("tap")
            bias = kw.pop("bias", None)
            for c0 in range(0, n, 512):
                self.be.note = (self.name, "fwd", self.flops * 512 / n)
                self.be.conv_gemm(a, a_chan_off, self.lay.K, self.taps, self.lay.w_f.t[c0:c0 + 512], 512, out, out_chan_off=c0,
                                  bias=None if bias is None else bias[c0:c0 + 512], **kw)
            return
        self.be.note = (self.name, "fwd", self.flops)
        self.be.conv_gemm(a, a_chan_off, self.lay.K, self.taps, self.lay.w_f.t, n, out, **kw)

    def dgrad(self, dz, out, **kw):
        """out[q][k] = sum_t dz[q - tap_t][:] . W[:, t, k]   (gradient w.r.t. the conv input frame)"""
        self.be.note = (self.name, "dgrad", getattr(self, "flops_bwd", self.flops))
        self.be.conv_gemm(dz, 0, self.lay.kd, [-t for t in self.taps], self.lay.w_d.t, self.lay.K, out, **kw)

    def wgrad(self, dz, x, x_chan_off, k_rows):
        lay = self.lay
        self.be.note = (self.name, "wgrad", self.flops)
        self.be.tn_gemm(dz, 0, lay.N, x, x_chan_off, lay.K, k_rows, [0] * lay.T, self.taps, self.partial,
                        lay.K, lay.T * lay.K, 1, self.splits, lay.N * lay.T * lay.K)
        # queued: the network's backward pass flushes all its split-K reductions in one launch (be.flush_sums)
        self.be.gather_sum_deferred(self.partial, lay.unpack, self.splits, lay.N * lay.T * lay.K, self.grad_flat)


class TransposedUp:
    """nn.ConvTranspose2d(C, C, 3, stride=2, padding=1, output_padding=1) (the generator's up-sampling with no_antialias_up=True,
    irc:495-499 / :512-516) between an input frame `src` (pad 1, ZERO ring) and C channels of a frame at twice the resolution.
    Forward = one stride-1 implicit GEMM over the input frame with 4 taps and 4 C output columns (the four sub-pixel phases, see
    layout.layout_convT) + a depth-to-space read; backward = the same layout run through the data- / weight-gradient GEMMs."""

    def __init__(self, be, packer: L.Packer, arena: L.ParamArena, key: str, C: int, src: L.Frame, device, name: str):
        self.be, self.C, self.src, self.arena, self.key = be, C, src, arena, key
        lay = L.layout_convT(packer, arena, key + ".weight", C, C)
        self.op = ConvOp(be, lay, L.taps_convT(src.wp), arena, src.rows, pixels=src.N * src.H * src.W, name=name)
        self.T = L.act_zeros(src.rows, 4 * C, device)          # forward output, depth-to-space order
        self.dT = L.act_zeros(src.rows, 4 * C, device)         # gradient w.r.t. it
        self.bias4 = torch.zeros(4 * C, device=device)
        off = arena.offset[key + ".bias"]
        self.bias_map = (off + torch.arange(4 * C, dtype=torch.int32) % C).to(torch.int32).to(device)
        self.colsum = torch.zeros(4 * C, device=device)
        self.cmap = torch.arange(C, dtype=torch.int32, device=device)
        self.ri = torch.zeros(src.rows, device=device, dtype=torch.int16)      # ring rows of the input frame
        be.row_index(self.ri, src.N, src.hp, src.wp, src.p, src.p + src.H, src.p, src.p + src.W)

    def refresh(self):
        self.be.gather_f32(self.arena.flat, self.bias_map, self.bias4)

    def forward(self, dst: L.Frame, dst_off: int):
        src, C = self.src, self.C
        self.op.fwd(src.t, 0, self.T, bias=self.bias4)
        # output pixel (2y + a, 2x + b) = column group a * 2 + b of input row (y, x): read the GEMM output as 2x2 blocks
        self.be.gather(View(self.T, 0, src.hp, src.wp, 2 * src.p, 2 * src.p, C), dst.view(dst_off), C, src.N, 2 * src.H, 2 * src.W, dst.p, 0)

    def backward(self, g: L.Frame, g_off: int, dx_out: torch.Tensor):
        """g: gradient w.r.t. the up-sampled map (channels [g_off, g_off + C) of frame g); dx_out [src.rows, C] receives the
        gradient w.r.t. the input frame (ring rows zeroed: they are padding, not reflected pixels)"""
        be, src, C = self.be, self.src, self.C
        be.gather(g.view(g_off), View(self.dT, 0, src.hp, src.wp), C, src.N, 2 * src.H, 2 * src.W, 2 * src.p, 0, dst_s2d=1)
        be.colsum(self.dT, 0, 4 * C, self.colsum)
        be.gather_sum(self.colsum, self.cmap, 4, C, self.arena.view(self.key + ".bias", self.arena.grad))
        self.op.wgrad(self.dT, src.t, 0, src.rows)
        self.op.dgrad(self.dT, dx_out, row_img=self.ri)


# ==========================================================================================
# generator (irc:425-569)
# ==========================================================================================
class GeneratorEngine:
    def __init__(self, be, B: int, H: int, W: int, device, ngf: int = 64, n_blocks: int = 9, training: bool = True,
                 arena: L.ParamArena = None, no_antialias_up: bool = False, no_antialias: bool = False, norm: str = "instance"):
        if norm not in ("instance", "none", "batch"):
            raise NotImplementedError(f"Normalization type [{norm}] not supported")          # irc:165
        if norm != "instance" and no_antialias_up:
            raise NotImplementedError("norm='none' / 'batch' together with no_antialias_up=True (bias-free transposed convolutions) is not built")
        self.norm = norm
        self.norm_on = norm != "none"            # statistics are needed (InstanceNorm2d or BatchNorm2d)
        self.bn = norm == "batch"
        self.bn_training, self.bn_updates = True, 1      # train()/eval() mode of the caller; forward calls the reference makes per iteration
        if min(H, W) < 8:
            raise ValueError("the generator needs at least 8 x 8 pixels (ReflectionPad2d(1) at a quarter of the resolution)")
        if no_antialias and (H % 4 or W % 4):
            raise NotImplementedError("no_antialias=True (space-to-depth stride-2 convolutions) needs H and W to be multiples of 4")
        if ngf != 64:
            raise NotImplementedError("ngf must be 64 (channel counts are tiled in units of 64)")
        self.be, self.B, self.H, self.W, self.dev, self.nb, self.training = be, B, H, W, device, n_blocks, training
        # Downsample halves with ceil (irc:307-310); the up-sampling doubles, so for sizes that are not multiples of 4 the
        # decoder maps are resized onto the skip-connection grids (irc:555-556, :562-563)
        self.H2, self.W2 = (H + 1) // 2, (W + 1) // 2
        self.H4, self.W4 = (self.H2 + 1) // 2, (self.W2 + 1) // 2
        if no_antialias_up and (2 * self.H4 != self.H2 or 2 * self.W4 != self.W2 or 2 * self.H2 != H or 2 * self.W2 != W):
            raise NotImplementedError("no_antialias_up=True with sizes that are not multiples of 4 (a bilinear resize behind the transposed "
                                      "convolution, irc:555-556) is not built")
        self.convT = bool(no_antialias_up)
        self.noaa = bool(no_antialias)          # stride-2 down-sampling convolutions instead of conv + blur (irc:468, :474, :482)
        self.arena = arena or L.ParamArena(generator_shapes(1, 3, ngf, n_blocks, no_antialias_up, norm), device)
        if ("inc.1.bias" in self.arena.offset) != (norm == "instance") or ("inc.2.weight" in self.arena.offset) != self.bn:
            raise ValueError("the parameter arena does not match norm=%r (convolution biases exist exactly with InstanceNorm, irc:452-455)" % norm)
        if self.convT and "up1_up.weight" not in self.arena.offset:
            raise ValueError("no_antialias_up=True needs an arena with the ConvTranspose2d parameters up{1,2}_up.{weight,bias}")
        self.packer = L.Packer(self.arena)
        A, P = self.arena, self.packer
        H2, W2, H4, W4 = self.H2, self.W2, self.H4, self.W4
        F = lambda h, w, p, c: L.Frame(B, h, w, p, c, device)
        # ---- buffers (forward)
        self.E_in = L.act_zeros(B * H * W, 64, device)
        self.Z0 = F(H, W, 0, 64)
        self.cat2 = F(H, W, 1, 192)       # [0:128) up2_up output, [128:192) x0
        self.cat1 = F(H2, W2, 1, 384)     # [0:256) up1_up output, [256:384) x1
        if self.noaa:
            # stride-2 convolutions read 2x2 space-to-depth copies of the padded x0 / x1 frames and write on the block grid
            self.hb1, self.wb1, self.hb2, self.wb2 = (H + 2) // 2, (W + 2) // 2, (H2 + 2) // 2, (W2 + 2) // 2
            r1, r2 = B * self.hb1 * self.wb1, B * self.hb2 * self.wb2
            self.Sx0, self.Z1s = L.act_zeros(r1, 256, device), L.act_zeros(r1, 128, device)
            self.Sx1, self.Z2s = L.act_zeros(r2, 512, device), L.act_zeros(r2, 256, device)
        else:
            self.Z1 = F(H, W, 1, 128)
            self.Z2 = F(H2, W2, 1, 256)
        self.X = [F(H4, W4, 1, 256) for _ in range(n_blocks + 1)]
        self.Za = [F(H4, W4, 1, 256) for _ in range(n_blocks)]
        self.Hh = [F(H4, W4, 1, 256) for _ in range(n_blocks)]
        self.Zb = [F(H4, W4, 1, 256) for _ in range(n_blocks)]
        self.Z3 = F(H2, W2, 1, 128)
        self.Z4 = F(H, W, 1, 64)
        self.y4 = F(H, W, 3, 64)
        # per-vertical-tap partial products of the output head: only materialised when the tap reduction is NOT fused into the GEMM
        self.P = torch.zeros(8 if getattr(be, "fused_outc", False) else self.y4.rows, 32, device=device)
        self.fake = torch.zeros(B, 3, H, W, device=device)
        st = lambda c: torch.zeros(B, c, 2, device=device)
        self.st0, self.st1, self.st2, self.st3, self.st4 = st(64), st(128), st(256), st(128), st(64)
        self.sta = [st(256) for _ in range(n_blocks)]
        self.stb = [st(256) for _ in range(n_blocks)]
        self.bsum = st(256)
        if self.bn:
            # nn.BatchNorm2d (irc:158-159): running statistics (the module surface swaps in its registered buffers), effective
            # moments per site, and which statistics buffer belongs to which state_dict prefix
            self.bn_state = new_bn_state(generator_bn_sites(ngf, n_blocks), device)
            sts = [self.st0, self.st1, self.st2] + [s_ for b in range(n_blocks) for s_ in (self.sta[b], self.stb[b])] + [self.st3, self.st4]
            self._bn_site = {id(s_): (k, c) for s_, (k, c) in zip(sts, generator_bn_sites(ngf, n_blocks))}
            self.eff = {k: torch.zeros(B, c, 2, device=device) for k, c in generator_bn_sites(ngf, n_blocks)}
            self._bn_done = set()
        # image index of every frame row (-1 = padding ring) for the convolutions that emit InstanceNorm statistics
        self.ri_full = torch.zeros(self.Z4.rows, device=device, dtype=torch.int16)
        self.ri_half = torch.zeros(self.Z3.rows, device=device, dtype=torch.int16)
        be.row_index(self.ri_full, B, H + 2, W + 2, 1, H + 1, 1, W + 1)
        be.row_index(self.ri_half, B, H2 + 2, W2 + 2, 1, H2 + 1, 1, W2 + 1)
        # ResNet blocks: InstanceNorm statistics from the GEMM epilogue + a streaming apply pass, instead of the cluster kernel that
        # computes them itself (be.res_epilogue_stats; measured per box, see profiles/)
        self.res_epi = bool(getattr(be, "res_epilogue_stats", False)) and self.norm == "instance"
        if self.res_epi:
            self.ri_quarter = torch.zeros(self.X[0].rows, device=device, dtype=torch.int16)
            be.row_index(self.ri_quarter, B, H4 + 2, W4 + 2, 1, H4 + 1, 1, W4 + 1)
        # ---- stencil tables
        mk = lambda my, mx: L.make_tables(my, mx, device)
        self.t_down1 = mk(L.down_matrix(H), L.down_matrix(W))
        self.t_down2 = mk(L.down_matrix(H2), L.down_matrix(W2))
        # UpsampleAA (irc:350-355), followed - only where the doubled size misses the skip connection's - by the bilinear
        # align_corners resize of irc:555-556 / :562-563, composed into ONE separable operator per axis
        up_to = lambda n, m: L.up_matrix(n) if 2 * n == m else L.resize_matrix(m, 2 * n) @ L.up_matrix(n)
        m_up1, m_up2 = (up_to(H4, H2), up_to(W4, W2)), (up_to(H2, H), up_to(W2, W))
        self.t_up1 = mk(*m_up1)
        self.t_up2 = mk(*m_up2)
        self.t_down1_T = mk(L.down_matrix(H).T, L.down_matrix(W).T)
        self.t_down2_T = mk(L.down_matrix(H2).T, L.down_matrix(W2).T)
        self.t_up1_T = mk(m_up1[0].T, m_up1[1].T)      # 6 x 6 taps, taken from a shared-memory patch
        self.t_up2_T = mk(m_up2[0].T, m_up2[1].T)
        self.t_fold1 = mk(L.fold_matrix(H4, 1), L.fold_matrix(W4, 1))
        self.t_fold3 = mk(L.fold_matrix(H, 3), L.fold_matrix(W, 3))
        # ---- weights
        wp1, wp2, wp4, wp3 = self.cat2.wp, self.cat1.wp, self.X[0].wp, self.y4.wp
        self.inc = ConvOp(be, L.layout_im2col(P, A, "inc.1.weight", 64, 1, 7), [0], A, self.Z0.rows, pixels=B * H * W, name="G.inc")
        if self.noaa:
            self.down1 = ConvOp(be, L.layout_s2d_k3(P, A, "down1.0.weight", 128, 64), [0, 1, self.wb1, self.wb1 + 1], A, self.Z1s.shape[0],
                                pixels=B * H2 * W2, name="G.down1")
            self.down2 = ConvOp(be, L.layout_s2d_k3(P, A, "down2.0.weight", 256, 128), [0, 1, self.wb2, self.wb2 + 1], A, self.Z2s.shape[0],
                                pixels=B * H4 * W4, name="G.down2")
        else:
            self.down1 = ConvOp(be, L.layout_std(P, A, "down1.0.weight", 128, 64, 3, 3), L.taps_centered(3, 3, wp1), A, self.Z1.rows, pixels=B * H * W, name="G.down1")
            self.down2 = ConvOp(be, L.layout_std(P, A, "down2.0.weight", 256, 128, 3, 3), L.taps_centered(3, 3, wp2), A, self.Z2.rows, pixels=B * H2 * W2, name="G.down2")
        self.res = []
        for b in range(n_blocks):
            self.res.append(tuple(ConvOp(be, L.layout_std(P, A, f"resblocks.{b}.conv_block.{j}.weight", 256, 256, 3, 3),
                                         L.taps_centered(3, 3, wp4), A, self.X[0].rows, pixels=B * H4 * W4, name="G.res") for j in (1, 5)))
        self.up1 = ConvOp(be, L.layout_std(P, A, "up1_conv.0.weight", 128, 384, 3, 3), L.taps_centered(3, 3, wp2), A, self.Z3.rows, pixels=B * H2 * W2, name="G.up1")
        self.up2 = ConvOp(be, L.layout_std(P, A, "up2_conv.0.weight", 64, 192, 3, 3), L.taps_centered(3, 3, wp1), A, self.Z4.rows, pixels=B * H * W, name="G.up2")
        self.outc = ConvOp(be, L.layout_outc(P, A, "outc.1.weight", 3, 64, 7), [(r - 3) * wp3 for r in range(7)], A, self.y4.rows,
                           bias_name="outc.1.bias", pixels=B * H * W, name="G.outc")
        self.outc_shifts = [(0, s - 3) for s in range(7)]      # (dy, dx) of the horizontal taps
        if self.convT:
            # up-sampling by transposed convolutions (irc:495-499, :512-516): up2_up needs the activated map as a frame of its own
            self.A3 = F(H2, W2, 1, 128)
            self.up1_t = TransposedUp(be, P, A, "up1_up", 256, self.X[n_blocks], device, "G.up1_up")
            self.up2_t = TransposedUp(be, P, A, "up2_up", 128, self.A3, device, "G.up2_up")
        P.finish()
        if training:
            self._alloc_backward()

    # ---- norm='none' (irc:158-163): no statistics; the ReLU moves from the normalise-and-activate pass into the GEMM epilogue
    def _st(self, st):
        return st if self.norm_on else None

    def _bn_eff(self, st, cnt):
        """BatchNorm: per-image sums -> effective moments of the whole batch (once per forward pass and site), running statistics"""
        key, c = self._bn_site[id(st)]
        if key not in self._bn_done:
            A = self.arena
            rm, rv = self.bn_state[key]
            self.be.bn_finalize(st, self.B, self.B, c, float(cnt), A.view(key + ".weight"), A.view(key + ".bias"), rm, rv, self.eff[key],
                                training=self.bn_training, updates=self.bn_updates)
            self._bn_done.add(key)
        return key, self.eff[key]

    def _na(self, st, cnt, act=ACT_RELU):
        """keyword arguments of an apply pass: normalise (InstanceNorm2d / BatchNorm2d) + activation, or a plain copy when there is
        no norm layer and the map is already activated"""
        if self.bn:
            return dict(stats=self._bn_eff(st, cnt)[1], cnt=self.B * cnt, eps=-1.0, act=act)
        return dict(stats=st, cnt=cnt, eps=EPS, act=act) if self.norm_on else {}

    def _nb(self, st, cnt, act=ACT_RELU):
        """keyword arguments of a norm + activation backward pass"""
        if self.bn:
            key, _ = self._bn_site[id(st)]
            A = self.arena
            return dict(stats=self.eff[key], cnt=self.B * cnt, eps=-1.0, act=act,
                        bn=dict(group=self.B, gamma=A.view(key + ".weight"), beta=A.view(key + ".bias"),
                                dgamma=A.view(key + ".weight", A.grad), dbeta=A.view(key + ".bias", A.grad)))
        return dict(stats=st if self.norm_on else None, cnt=cnt, eps=EPS, act=act)

    def _epi(self):
        return {} if self.norm_on else dict(act=ACT_RELU)

    def _alloc_backward(self):
        B, H, W, dev = self.B, self.H, self.W, self.dev
        H2, W2, H4, W4 = self.H2, self.W2, self.H4, self.W4
        F = lambda h, w, p, c: L.Frame(B, h, w, p, c, dev)
        self.E_out = L.act_zeros(self.y4.rows, 64, dev)
        self.G4 = F(H, W, 3, 64)
        self.dZ4 = F(H, W, 1, 64)
        self.Gcat2 = F(H, W, 1, 192)
        self.dZ3 = F(H2, W2, 1, 128)
        self.Gcat1 = F(H2, W2, 1, 384)
        self.dOut = [F(H4, W4, 1, 256), F(H4, W4, 1, 256)]
        self.dZb, self.Gh, self.dZa, self.Gx = (F(H4, W4, 1, 256) for _ in range(4))
        self.dZ0 = F(H, W, 0, 64)
        # gradients w.r.t. the activated layer outputs, after the transposed stencil / reflection fold
        self.g3 = F(H2, W2, 0, 128)
        if self.noaa:
            self.dZ1s, self.dZ2s = torch.zeros_like(self.Z1s), torch.zeros_like(self.Z2s)
            self.dSx0, self.dSx1 = torch.zeros_like(self.Sx0), torch.zeros_like(self.Sx1)
        else:
            self.dZ2 = F(H2, W2, 1, 256)
            self.Gx1 = F(H2, W2, 1, 128)
            self.dZ1 = F(H, W, 1, 128)
            self.Gx0 = F(H, W, 1, 64)
            self.g2 = F(H2, W2, 0, 256)
            self.g1 = F(H, W, 0, 128)
        if self.convT:
            self.GA3 = F(H2, W2, 1, 128)

    # ------------------------------------------------------------------ forward
    def refresh_weights(self):
        self.packer.refresh(self.be)
        if self.convT:
            self.up1_t.refresh(); self.up2_t.refresh()

    def forward(self, ir: torch.Tensor) -> torch.Tensor:
        """ir: fp32 [B,1,H,W] -> fake fp32 [B,3,H,W] (irc:533-569)"""
        be, B, H, W = self.be, self.B, self.H, self.W
        H2, W2, H4, W4 = self.H2, self.W2, self.H4, self.W4
        assert ir.shape == (B, 1, H, W) and ir.dtype == torch.float32 and ir.is_contiguous()
        if self.bn:
            self._bn_done.clear()
        # inc: reflect-pad 3 + 7x7 conv over 49 of 64 operand slots, IN + ReLU into cat2[128:192)
        if getattr(be, "direct_smallk", False):
            # direct convolution from the fp32 image; the im2col operand is only written when the weight gradient will read it
            be.note = ("G.inc", "fwd", self.inc.flops)
            be.smallk_conv_fwd(ir, None, None, None, B, H, W, 7, 1, 3, 1, H, W, 0, self.inc.lay.w_f.t, self.Z0.t, E=self.E_in if self.training else None,
                               **self._epi())
        else:
            be.im2col(ir, None, None, None, B, H, W, 7, 1, 3, 1, H, W, 0, self.E_in)
            self.inc.fwd(self.E_in, 0, self.Z0.t, **self._epi())
        if self.norm_on:
            be.in_stats(self.Z0.view(), 64, B, H, W, self.st0)
        be.gather(self.Z0.view(), self.cat2.view(128), 64, B, H, W, 1, 0, **self._na(self.st0, H * W))
        if self.noaa:
            self._encoder_strided()
        else:
            self._encoder_antialiased()
        self._bottleneck_and_decoder()
        return self.output_head()

    def _vZ1s(self, t):
        return View(t, 0, self.hb1, self.wb1, 0, 0)

    def _vZ2s(self, t):
        return View(t, 0, self.hb2, self.wb2, 0, 0)

    def _encoder_strided(self):
        """no_antialias=True (irc:468): down1 / down2 are 3x3 stride-2 convolutions = 2x2 convolutions over space-to-depth copies of
        the padded x0 / x1 frames; their InstanceNorm + ReLU output IS x1 / x2 (no blur module, irc:474, :482)"""
        be, B, H, W = self.be, self.B, self.H, self.W
        H2, W2, H4, W4 = self.H2, self.W2, self.H4, self.W4
        be.gather(self.Z0.view(), View(self.Sx0, 0, self.hb1, self.wb1), 64, B, H, W, 1, 0, **self._na(self.st0, H * W), dst_s2d=1)
        self.down1.fwd(self.Sx0, 0, self.Z1s, **self._epi())
        v1 = self._vZ1s(self.Z1s)
        if self.norm_on:
            be.in_stats(v1, 128, B, H2, W2, self.st1)
        be.gather(v1, self.cat1.view(256), 128, B, H2, W2, 1, 0, **self._na(self.st1, H2 * W2))
        be.gather(v1, View(self.Sx1, 0, self.hb2, self.wb2), 128, B, H2, W2, 1, 0, **self._na(self.st1, H2 * W2), dst_s2d=1)
        self.down2.fwd(self.Sx1, 0, self.Z2s, **self._epi())
        v2 = self._vZ2s(self.Z2s)
        if self.norm_on:
            be.in_stats(v2, 256, B, H4, W4, self.st2)
        be.gather(v2, self.X[0].view(), 256, B, H4, W4, 1, 1, **self._na(self.st2, H4 * W4))

    def _encoder_antialiased(self):
        be, B, H, W = self.be, self.B, self.H, self.W
        H2, W2, H4, W4 = self.H2, self.W2, self.H4, self.W4
        # down1: 3x3 zero-pad conv at full resolution, then IN + ReLU + blur-downsample fused
        self.down1.fwd_stats(self.cat2.t, 128, self.Z1.t, self._st(self.st1), self.ri_full, B, self.Z1.hp * self.Z1.wp, self.Z1.view(), 128, H, W, **self._epi())
        be.gather(self.Z1.view(), self.cat1.view(256), 128, B, H2, W2, 1, 0, tables=self.t_down1, **self._na(self.st1, H * W))
        # down2
        self.down2.fwd_stats(self.cat1.t, 256, self.Z2.t, self._st(self.st2), self.ri_half, B, self.Z2.hp * self.Z2.wp, self.Z2.view(), 256, H2, W2, **self._epi())
        be.gather(self.Z2.view(), self.X[0].view(), 256, B, H4, W4, 1, 1, tables=self.t_down2, **self._na(self.st2, H2 * W2))

    def _bottleneck_and_decoder(self):
        be, B, H, W = self.be, self.B, self.H, self.W
        H2, W2, H4, W4 = self.H2, self.W2, self.H4, self.W4
        # 9 ResNet blocks (irc:362-418): reflect halo written by the apply pass
        n4 = H4 * W4
        for b in range(self.nb):
            c1, c2 = self.res[b]
            if self.res_epi:
                fr = self.X[b]
                c1.fwd_stats(fr.t, 0, self.Za[b].t, self.sta[b], self.ri_quarter, B, fr.hp * fr.wp, self.Za[b].view(), 256, H4, W4)
                be.gather(self.Za[b].view(), self.Hh[b].view(), 256, B, H4, W4, 1, 1, stats=self.sta[b], cnt=n4, eps=EPS, act=ACT_RELU)
                c2.fwd_stats(self.Hh[b].t, 0, self.Zb[b].t, self.stb[b], self.ri_quarter, B, fr.hp * fr.wp, self.Zb[b].view(), 256, H4, W4)
                halo = 0 if (self.convT and b == self.nb - 1) else 1
                be.gather(self.Zb[b].view(), self.X[b + 1].view(), 256, B, H4, W4, 1, halo, stats=self.stb[b], cnt=n4, eps=EPS, act=ACT_NONE, res=self.X[b].view())
                continue
            c1.fwd(self.X[b].t, 0, self.Za[b].t, **self._epi())
            if self.norm == "instance":
                be.in_apply(self.Za[b].view(), self.Hh[b].view(), 256, B, H4, W4, 1, 1, self.sta[b], eps=EPS, act=ACT_RELU)
            elif self.bn:
                be.in_stats(self.Za[b].view(), 256, B, H4, W4, self.sta[b])
                be.gather(self.Za[b].view(), self.Hh[b].view(), 256, B, H4, W4, 1, 1, **self._na(self.sta[b], n4))
            else:
                be.gather(self.Za[b].view(), self.Hh[b].view(), 256, B, H4, W4, 1, 1)          # reflected ring around the activated map
            c2.fwd(self.Hh[b].t, 0, self.Zb[b].t)
            # the last block output feeds only the up-sampling: a transposed convolution wants a ZERO ring, not the reflected one
            halo = 0 if (self.convT and b == self.nb - 1) else 1
            if self.norm == "instance":
                be.in_apply(self.Zb[b].view(), self.X[b + 1].view(), 256, B, H4, W4, 1, halo, self.stb[b], eps=EPS, act=ACT_NONE, res=self.X[b].view())
            elif self.bn:
                be.in_stats(self.Zb[b].view(), 256, B, H4, W4, self.stb[b])
                be.gather(self.Zb[b].view(), self.X[b + 1].view(), 256, B, H4, W4, 1, halo, res=self.X[b].view(), **self._na(self.stb[b], n4, ACT_NONE))
            else:
                be.gather(self.Zb[b].view(), self.X[b + 1].view(), 256, B, H4, W4, 1, halo, res=self.X[b].view())
        if self.convT:
            # up1 / up2 by ConvTranspose2d (irc:495-499, :512-516)
            self.up1_t.forward(self.cat1, 0)
            self.up1.fwd_stats(self.cat1.t, 0, self.Z3.t, self._st(self.st3), self.ri_half, B, self.Z3.hp * self.Z3.wp, self.Z3.view(), 128, H2, W2, **self._epi())
            be.gather(self.Z3.view(), self.A3.view(), 128, B, H2, W2, 1, 0, **self._na(self.st3, H2 * W2))
            self.up2_t.forward(self.cat2, 0)
        else:
            # up1: UpsampleAA into cat1[0:256), conv on the concatenation
            be.gather(self.X[self.nb].view(), self.cat1.view(0), 256, B, H2, W2, 1, 0, tables=self.t_up1)
            self.up1.fwd_stats(self.cat1.t, 0, self.Z3.t, self._st(self.st3), self.ri_half, B, self.Z3.hp * self.Z3.wp, self.Z3.view(), 128, H2, W2, **self._epi())
            # up2: IN + ReLU + UpsampleAA fused into cat2[0:128)
            be.gather(self.Z3.view(), self.cat2.view(0), 128, B, H, W, 1, 0, tables=self.t_up2, **self._na(self.st3, H2 * W2))
        self.up2.fwd_stats(self.cat2.t, 0, self.Z4.t, self._st(self.st4), self.ri_full, B, self.Z4.hp * self.Z4.wp, self.Z4.view(), 64, H, W, **self._epi())
        be.gather(self.Z4.view(), self.y4.view(), 64, B, H, W, 3, 1, **self._na(self.st4, H * W))

    def output_head(self) -> torch.Tensor:
        """outc (irc:527-531) on the frame y4: 7x7 reflect conv 64->3 as a GEMM over the 7 vertical taps (21 of 32 outputs), then
        the horizontal tap reduction + bias + tanh - inside the GEMM epilogue (tiles overlap by 6 rows), or as a second pass"""
        be, B, H, W = self.be, self.B, self.H, self.W
        if getattr(be, "fused_outc", False):
            self.outc.fwd(self.y4.t, 0, self.P, bias=self.outc.bias(),
                          tap=dict(out=self.fake, nshift=7, nco=3, H=H, W=W, hp=self.y4.hp, wp=self.y4.wp, oy=3, ox=3, act=ACT_TANH))
        else:
            self.outc.fwd(self.y4.t, 0, self.P)
            be.tap_reduce(self.P, self.outc_shifts, 3, B, H, W, self.y4.hp, self.y4.wp, 3, 3, self.outc.bias(), ACT_TANH, self.fake)
        return self.fake

    # ------------------------------------------------------------------ backward
    def backward(self, dfake: torch.Tensor, after_blocks=None) -> None:
        """dfake: fp32 [B,3,H,W] = dL/dfake.  Fills arena.grad for every generator parameter.  `after_blocks`
        (optional callable) runs once the gradients of outc, up2, up1 and all ResNet blocks are final."""
        be, B, H, W = self.be, self.B, self.H, self.W
        H2, W2, H4, W4 = self.H2, self.W2, self.H4, self.W4
        y4 = self.y4
        # outc: tanh' and horizontal tap expansion, then weight / data gradients over the vertical taps
        be.tap_expand(dfake, self.fake, self.outc_shifts, 3, B, H, W, y4.hp, y4.wp, 3, 3, self.E_out,
                      dbias=self.arena.view("outc.1.bias", self.arena.grad), live_cols_only=True)      # E_out's other columns stay zero
        self.outc.wgrad(self.E_out, y4.t, 0, y4.rows)
        self.outc.dgrad(self.E_out, self.G4.t, k_live=21)      # E_out holds 7 taps x 3 channels; its other columns are zeros
        # up2_conv
        be.fold_inplace(self.G4.t, 0, 64, B, H, W, 3)             # ReflectionPad2d(3)^T on the border pixels only
        be.in_bwd(self.Z4.view(), self.G4.view(), self.dZ4.view(), 64, B, H, W, **self._nb(self.st4, H * W, ACT_RELU), bsum=self.bsum)
        self.up2.wgrad(self.dZ4.t, self.cat2.t, 0, self.cat2.rows)
        self.up2.dgrad(self.dZ4.t, self.Gcat2.t)
        # up1_conv (through UpsampleAA^T, or the transposed convolution's data gradient)
        if self.convT:
            self.up2_t.backward(self.Gcat2, 0, self.GA3.t)
            g_up1 = self.GA3
        else:
            be.gather(self.Gcat2.view(0), self.g3.view(), 128, B, H2, W2, 0, 0, tables=self.t_up2_T)
            g_up1 = self.g3
        be.in_bwd(self.Z3.view(), g_up1.view(), self.dZ3.view(), 128, B, H2, W2, **self._nb(self.st3, H2 * W2, ACT_RELU),
                  bsum=self.bsum)
        self.up1.wgrad(self.dZ3.t, self.cat1.t, 0, self.cat1.rows)
        self.up1.dgrad(self.dZ3.t, self.Gcat1.t)
        # gradient w.r.t. the last ResNet block output = UpsampleAA^T of Gcat1[0:256)
        cur = self.dOut[0]
        if self.convT:
            self.up1_t.backward(self.Gcat1, 0, cur.t)
        else:
            be.gather(self.Gcat1.view(0), cur.view(), 256, B, H4, W4, 1, 0, tables=self.t_up1_T)
        n4 = H4 * W4
        for b in reversed(range(self.nb)):
            c1, c2 = self.res[b]
            # `cur` (gradient w.r.t. the reflection-padded X[b+1]) still carries its ring: the fold is linear, so it is applied
            # where the gradient is consumed (here, on load) and once more when the stream leaves the blocks
            be.in_bwd(self.Zb[b].view(), cur.view(), self.dZb.view(), 256, B, H4, W4, **self._nb(self.stb[b], n4, ACT_NONE), bsum=self.bsum,
                      fold_pad=1)
            c2.wgrad(self.dZb.t, self.Hh[b].t, 0, self.dZb.rows)
            c2.dgrad(self.dZb.t, self.Gh.t)
            # ReflectionPad2d(1)^T of Gh is folded inside the backward pass (or by a separate in-place pass when the map is
            # too large for the single-pass cluster kernel)
            be.in_bwd(self.Za[b].view(), self.Gh.view(), self.dZa.view(), 256, B, H4, W4, **self._nb(self.sta[b], n4, ACT_RELU),
                      bsum=self.bsum, fold_pad=1)
            c1.wgrad(self.dZa.t, self.X[b].t, 0, self.dZa.rows)
            # data gradient of conv1 + the residual-stream gradient, ring included (unfolded: fold(a + b) = fold(a) + fold(b))
            nxt = self.dOut[1] if cur is self.dOut[0] else self.dOut[0]
            c1.dgrad(self.dZa.t, nxt.t, addend=View(cur.t, 0, 0, 0))
            cur = nxt
        be.fold_inplace(cur.t, 0, 256, B, H4, W4, 1)
        be.flush_sums()                   # weight gradients of outc, up2, up1 and the ResNet blocks are final
        if after_blocks is not None:
            after_blocks()
        if self.noaa:
            # stride-2 encoder: x2 is the activated down2 output itself; the data gradients come back in space-to-depth order
            v1, v2 = self._vZ1s, self._vZ2s
            be.in_bwd(v2(self.Z2s), cur.view(), v2(self.dZ2s), 256, B, H4, W4, **self._nb(self.st2, H4 * W4, ACT_RELU), bsum=self.bsum)
            self.down2.wgrad(self.dZ2s, self.Sx1, 0, self.Sx1.shape[0])
            self.down2.dgrad(self.dZ2s, self.dSx1)
            # x1 feeds down2 and the up1 skip connection
            be.in_bwd(v1(self.Z1s), self.Gcat1.view(256), v1(self.dZ1s), 128, B, H2, W2, **self._nb(self.st1, H2 * W2, ACT_RELU),
                      g2=View(self.dSx1, 0, self.hb2, self.wb2, 1, 1, 128), bsum=self.bsum)
            self.down1.wgrad(self.dZ1s, self.Sx0, 0, self.Sx0.shape[0])
            self.down1.dgrad(self.dZ1s, self.dSx0)
            # inc: x0 feeds down1 and the up2 skip connection
            be.in_bwd(self.Z0.view(), self.Gcat2.view(128), self.dZ0.view(), 64, B, H, W, **self._nb(self.st0, H * W, ACT_RELU),
                      g2=View(self.dSx0, 0, self.hb1, self.wb1, 1, 1, 64), bsum=self.bsum)
            self.inc.wgrad(self.dZ0.t, self.E_in, 0, self.dZ0.rows)
            be.flush_sums()
            return
        # down2 (through Downsample^T)
        be.gather(cur.view(), self.g2.view(), 256, B, H2, W2, 0, 0, tables=self.t_down2_T)
        be.in_bwd(self.Z2.view(), self.g2.view(), self.dZ2.view(), 256, B, H2, W2, **self._nb(self.st2, H2 * W2, ACT_RELU),
                  bsum=self.bsum)
        self.down2.wgrad(self.dZ2.t, self.cat1.t, 256, self.cat1.rows)
        self.down2.dgrad(self.dZ2.t, self.Gx1.t)
        # down1: x1 feeds down2 and the up1 skip connection
        be.gather(self.Gcat1.view(256), self.g1.view(), 128, B, H, W, 0, 0, tables=self.t_down1_T, src2=self.Gx1.view())
        be.in_bwd(self.Z1.view(), self.g1.view(), self.dZ1.view(), 128, B, H, W, **self._nb(self.st1, H * W, ACT_RELU),
                  bsum=self.bsum)
        self.down1.wgrad(self.dZ1.t, self.cat2.t, 128, self.cat2.rows)
        self.down1.dgrad(self.dZ1.t, self.Gx0.t)
        # inc: x0 feeds down1 and the up2 skip connection
        be.in_bwd(self.Z0.view(), self.Gcat2.view(128), self.dZ0.view(), 64, B, H, W, **self._nb(self.st0, H * W, ACT_RELU),
                  g2=self.Gx0.view(), bsum=self.bsum)
        self.inc.wgrad(self.dZ0.t, self.E_in, 0, self.dZ0.rows)
        be.flush_sums()


# ==========================================================================================
# discriminator (irc:576-635)
# ==========================================================================================
class DiscriminatorEngine:
    def __init__(self, be, n_img: int, H: int, W: int, device, arena: L.ParamArena = None, packer: L.Packer = None, layouts=None,
                 norm: str = "instance"):
        if norm not in ("instance", "none", "batch"):
            raise NotImplementedError(f"Normalization type [{norm}] not supported")          # irc:165
        self.norm = norm
        self.norm_on = norm != "none"
        self.bn = norm == "batch"
        self.bn_training = True
        if H % 16 or W % 16 or H < 32 or W < 32:
            raise NotImplementedError("discriminator engine needs H, W multiples of 16 and >= 32 (70x70 PatchGAN receptive field)")
        self.be, self.n, self.H, self.W, self.dev = be, n_img, H, W, device
        own = packer is None
        self.arena = arena or L.ParamArena(discriminator_shapes(4, 64, norm), device)
        if ("model.2.bias" in self.arena.offset) != (norm == "instance") or ("model.3.weight" in self.arena.offset) != self.bn:
            raise ValueError("the parameter arena does not match norm=%r (model.2/5/8 have biases exactly with InstanceNorm, irc:590-593)" % norm)
        self.packer = packer or L.Packer(self.arena)
        A, P = self.arena, self.packer
        n = n_img
        H1, W1, H2, W2, H3, W3 = H // 2, W // 2, H // 4, W // 4, H // 8, W // 8
        self.H1, self.W1, self.H2, self.W2, self.H3, self.W3 = H1, W1, H2, W2, H3, W3
        self.hb0, self.wb0 = (H1 + 2) // 2, (W1 + 2) // 2          # s2d blocks of the padded model.0 output
        self.hb2, self.wb2 = (H2 + 2) // 2, (W2 + 2) // 2          # s2d blocks of the padded model.2 output
        bf = lambda r, c: L.act_zeros(r, c, device)
        self.rows0 = n * (H1 + 2) * (W1 + 2)
        self.E0 = bf(self.rows0, 64)
        self.row_img0 = torch.zeros(self.rows0, device=device, dtype=torch.int16)
        self.S0 = bf(self.rows0, 64)                                # == [n*hb0*wb0, 256] space-to-depth operand
        self.S0v = self.S0.view(n * self.hb0 * self.wb0, 256)
        self.Z2 = bf(n * self.hb0 * self.wb0, 128)
        self.S2 = bf(n * self.hb2 * self.wb2, 512)
        self.Z5 = bf(n * self.hb2 * self.wb2, 256)
        self.X8 = L.Frame(n, H3, W3, 1, 256, device)
        self.Z8 = bf(self.X8.rows, 512)
        self.H8o, self.W8o = H3 - 1, W3 - 1                         # model.8 output (k4 s1 p1)
        self.X11 = L.Frame(n, self.H8o, self.W8o, 1, 512, device)
        self.Ho, self.Wo = self.H8o - 1, self.W8o - 1               # model.11 output
        self.P11 = torch.zeros(self.X11.rows, 32, device=device)
        self.pred = torch.zeros(n, 1, self.Ho, self.Wo, device=device)
        st = lambda c: torch.zeros(n, c, 2, device=device)
        self.st2, self.st5, self.st8, self.bsum = st(128), st(256), st(512), st(512)
        if self.bn:
            # nn.BatchNorm2d (irc:158-159): batch statistics per forward CALL of the reference (real and fake halves are separate calls)
            self.bn_state = new_bn_state(DISCRIMINATOR_BN_SITES, device)
            self._bn_site = {id(s_): kc for s_, kc in zip((self.st2, self.st5, self.st8), DISCRIMINATOR_BN_SITES)}
            self.eff = {k: torch.zeros(n, c, 2, device=device) for k, c in DISCRIMINATOR_BN_SITES}
            self.group = n
        self.ri2 = torch.zeros(n * self.hb0 * self.wb0, device=device, dtype=torch.int16)      # live rows of the model.2 output grid
        be.row_index(self.ri2, n, self.hb0, self.wb0, 0, H2, 0, W2)
        # weights (shared between the 2B-image and B-image instances)
        if layouts is None:
            layouts = dict(
                l0=L.layout_im2col(P, A, "model.0.weight", 64, 4, 4),
                l2=L.layout_s2d(P, A, "model.2.weight", 128, 64),
                l5=L.layout_s2d(P, A, "model.5.weight", 256, 128),
                l8=L.layout_std(P, A, "model.8.weight", 512, 256, 4, 4),
                l11=L.layout_pointwise_taps(P, A, "model.11.weight", 512, 4))
        self.layouts = layouts
        self.c0 = ConvOp(be, layouts["l0"], [0], A, self.rows0, bias_name="model.0.bias", pixels=n * H1 * W1, name="D.0")
        self.c2 = ConvOp(be, layouts["l2"], [0, 1, self.wb0, self.wb0 + 1], A, self.Z2.shape[0], pixels=n * H2 * W2, name="D.2")
        self.c5 = ConvOp(be, layouts["l5"], [0, 1, self.wb2, self.wb2 + 1], A, self.Z5.shape[0], pixels=n * H3 * W3, name="D.5")
        self.c8 = ConvOp(be, layouts["l8"], L.taps_topleft(4, 4, self.X8.wp), A, self.X8.rows, pixels=n * self.H8o * self.W8o, name="D.8")
        self.c11 = ConvOp(be, layouts["l11"], [0], A, self.X11.rows, bias_name="model.11.bias", pixels=n * self.Ho * self.Wo, name="D.11")
        self.shifts11 = [(r, s) for r in range(4) for s in range(4)]    # (dy, dx) of the 16 taps
        if own:
            P.finish()
        # backward buffers
        self.dpred = torch.zeros_like(self.pred)
        self._scratch = torch.zeros(2, 512, device=device)
        self.E11 = bf(self.X11.rows, 64)
        self.G11 = L.Frame(n, self.H8o, self.W8o, 1, 512, device)
        self.dZ8 = bf(self.X8.rows, 512)
        self.G8 = L.Frame(n, H3, W3, 1, 256, device)
        self.dZ5 = bf(n * self.hb2 * self.wb2, 256)
        self.dS2 = bf(n * self.hb2 * self.wb2, 512)
        self.dZ2 = bf(n * self.hb0 * self.wb0, 128)
        self.dZ0 = bf(self.rows0, 64)
        self.dZ0v = self.dZ0.view(n * self.hb0 * self.wb0, 256)
        self.dE0 = bf(self.rows0, 64)

    def refresh_weights(self):
        self.packer.refresh(self.be)

    def _na(self, st, cnt):
        """keyword arguments of a normalise + LeakyReLU apply pass (InstanceNorm2d or BatchNorm2d)"""
        if not self.bn:
            return dict(stats=st, cnt=cnt, eps=EPS, act=ACT_LRELU, slope=0.2)
        key, c = self._bn_site[id(st)]
        A = self.arena
        rm, rv = self.bn_state[key]
        self.be.bn_finalize(st, self.n, self.group, c, float(cnt), A.view(key + ".weight"), A.view(key + ".bias"), rm, rv, self.eff[key],
                            training=self.bn_training)
        return dict(stats=self.eff[key], cnt=self.group * cnt, eps=-1.0, act=ACT_LRELU, slope=0.2)

    def _nb(self, st, cnt, want_wgrad):
        if not self.bn:
            return dict(stats=st if self.norm_on else None, cnt=cnt, eps=EPS, act=ACT_LRELU, slope=0.2)
        key, _ = self._bn_site[id(st)]
        A = self.arena
        # without want_wgrad (generator step) the parameter gradients land in a scratch row of bsum's sibling buffer
        dg, db = (A.view(key + ".weight", A.grad), A.view(key + ".bias", A.grad)) if want_wgrad else (self._scratch[0, :st.shape[1]], self._scratch[1, :st.shape[1]])
        return dict(stats=self.eff[key], cnt=self.group * cnt, eps=-1.0, act=ACT_LRELU, slope=0.2,
                    bn=dict(group=self.group, gamma=A.view(key + ".weight"), beta=A.view(key + ".bias"), dgamma=dg, dbeta=db))

    def _vZ2(self, t):
        return View(t, 0, self.hb0, self.wb0, 0, 0)

    def _vZ5(self, t):
        return View(t, 0, self.hb2, self.wb2, 0, 0)

    def _vZ8(self, t):
        return View(t, 0, self.X8.hp, self.X8.wp, 0, 0)

    def forward(self, ir: torch.Tensor, img: torch.Tensor, ir_b: torch.Tensor = None, img_b: torch.Tensor = None, keep_operand: bool = True) -> torch.Tensor:
        """D(cat[ir, img]) -> fp32 [n,1,Ho,Wo] raw scores (irc:632-635).  (ir_b, img_b), if given, is a second
        batch processed behind the first in the same launches (real and fake halves of the D step; InstanceNorm
        is per sample, so batching is exact).  keep_operand=False: the backward pass will not ask for weight gradients
        (generator step), so model.0's im2col operand need not be written."""
        be, n, H, W = self.be, self.n, self.H, self.W
        # model.0: 4x4 s2 conv on 4 channels = im2col with exactly 64 slots, rows in 2x2 sub-pixel order so the
        # GEMM output *is* the space-to-depth operand of model.2; bias + LeakyReLU in the epilogue
        n1 = ir.shape[0]
        r1 = n1 * (self.H1 + 2) * (self.W1 + 2)
        assert n1 == n if ir_b is None else n1 + ir_b.shape[0] == n
        parts = [(ir, img, n1, slice(0, r1))] + ([] if ir_b is None else [(ir_b, img_b, n - n1, slice(r1, None))])
        if getattr(be, "direct_smallk", False):
            # direct convolution (bias + LeakyReLU on the accumulators); the operand E0 is kept only for the weight gradient
            for a_, b_, k_, sl in parts:
                be.note = ("D.0", "fwd", self.c0.flops * k_ / n)
                be.smallk_conv_fwd(a_, b_, None, None, k_, H, W, 4, 2, 1, 0, self.H1, self.W1, 2, self.c0.lay.w_f.t, self.S0[sl], bias=self.c0.bias(),
                                   act=ACT_LRELU, slope=0.2, E=self.E0[sl] if keep_operand else None, row_img=self.row_img0[sl])
        else:
            for a_, b_, k_, sl in parts:
                be.im2col(a_, b_, None, None, k_, H, W, 4, 2, 1, 0, self.H1, self.W1, 2, self.E0[sl], row_img=self.row_img0[sl])
            self.c0.fwd(self.E0, 0, self.S0, bias=self.c0.bias(), act=ACT_LRELU, slope=0.2, row_img=self.row_img0)
        # model.2
        if not self.norm_on:
            # norm='none' (irc:158-163): LeakyReLU in the GEMM epilogues, the apply passes become plain copies into the next frame
            lr = dict(act=ACT_LRELU, slope=0.2)
            self.c2.fwd(self.S0v, 0, self.Z2, **lr)
            be.gather(self._vZ2(self.Z2), View(self.S2, 0, self.hb2, self.wb2), 128, n, self.H2, self.W2, 1, 0, dst_s2d=1)
            self.c5.fwd(self.S2, 0, self.Z5, **lr)
            be.gather(self._vZ5(self.Z5), self.X8.view(), 256, n, self.H3, self.W3, 1, 0)
            self.c8.fwd(self.X8.t, 0, self.Z8, **lr)
            be.gather(self._vZ8(self.Z8), self.X11.view(), 512, n, self.H8o, self.W8o, 1, 0)
            self.c11.fwd(self.X11.t, 0, self.P11)
            be.tap_reduce(self.P11, self.shifts11, 1, n, self.Ho, self.Wo, self.X11.hp, self.X11.wp, 0, 0, self.c11.bias(), ACT_NONE, self.pred)
            return self.pred
        if self.bn:
            self.group = n1          # one set of batch statistics per forward call of the reference (irc:1642, :1643, :1659)
        self.c2.fwd_stats(self.S0v, 0, self.Z2, self.st2, self.ri2, n, self.hb0 * self.wb0, self._vZ2(self.Z2), 128, self.H2, self.W2)
        be.gather(self._vZ2(self.Z2), View(self.S2, 0, self.hb2, self.wb2), 128, n, self.H2, self.W2, 1, 0, dst_s2d=1, **self._na(self.st2, self.H2 * self.W2))
        # model.5
        self.c5.fwd(self.S2, 0, self.Z5)
        if self.bn:
            be.in_stats(self._vZ5(self.Z5), 256, n, self.H3, self.W3, self.st5)
            be.gather(self._vZ5(self.Z5), self.X8.view(), 256, n, self.H3, self.W3, 1, 0, **self._na(self.st5, self.H3 * self.W3))
        else:
            be.in_apply(self._vZ5(self.Z5), self.X8.view(), 256, n, self.H3, self.W3, 1, 0, self.st5, eps=EPS, act=ACT_LRELU, slope=0.2)
        # model.8 (stride 1)
        self.c8.fwd(self.X8.t, 0, self.Z8)
        if self.bn:
            be.in_stats(self._vZ8(self.Z8), 512, n, self.H8o, self.W8o, self.st8)
            be.gather(self._vZ8(self.Z8), self.X11.view(), 512, n, self.H8o, self.W8o, 1, 0, **self._na(self.st8, self.H8o * self.W8o))
        else:
            be.in_apply(self._vZ8(self.Z8), self.X11.view(), 512, n, self.H8o, self.W8o, 1, 0, self.st8, eps=EPS, act=ACT_LRELU, slope=0.2)
        # model.11: 512->1, per-tap partial products then the 16-tap shifted reduction
        self.c11.fwd(self.X11.t, 0, self.P11)
        be.tap_reduce(self.P11, self.shifts11, 1, n, self.Ho, self.Wo, self.X11.hp, self.X11.wp, 0, 0, self.c11.bias(), ACT_NONE, self.pred)
        return self.pred

    def backward(self, dpred: torch.Tensor, want_wgrad: bool, dinput: Optional[torch.Tensor] = None, c_first: int = 1,
                 accumulate: bool = True) -> None:
        """dpred = dL/dscores.  want_wgrad fills arena.grad; dinput (fp32 [n,c,H,W]) receives the gradient w.r.t.
        input channels [c_first, c_first+c) of cat[ir, img] (default: accumulate into the 3 image channels)."""
        be, n = self.be, self.n
        G = self.arena.grad
        be.tap_expand(dpred, None, self.shifts11, 1, n, self.Ho, self.Wo, self.X11.hp, self.X11.wp, 0, 0, self.E11,
                      dbias=self.arena.view("model.11.bias", G) if want_wgrad else None, live_cols_only=True)
        if want_wgrad:
            self.c11.wgrad(self.E11, self.X11.t, 0, self.X11.rows)
        self.c11.dgrad(self.E11, self.G11.t)
        be.in_bwd(self._vZ8(self.Z8), self.G11.view(), self._vZ8(self.dZ8), 512, n, self.H8o, self.W8o, bsum=self.bsum,
                  **self._nb(self.st8, self.H8o * self.W8o, want_wgrad))
        if want_wgrad:
            self.c8.wgrad(self.dZ8, self.X8.t, 0, self.X8.rows)
        self.c8.dgrad(self.dZ8, self.G8.t)
        be.in_bwd(self._vZ5(self.Z5), self.G8.view(), self._vZ5(self.dZ5), 256, n, self.H3, self.W3, bsum=self.bsum,
                  **self._nb(self.st5, self.H3 * self.W3, want_wgrad))
        if want_wgrad:
            self.c5.wgrad(self.dZ5, self.S2, 0, self.S2.shape[0])
        self.c5.dgrad(self.dZ5, self.dS2)
        be.in_bwd(self._vZ2(self.Z2), View(self.dS2, 0, self.hb2, self.wb2, 1, 1, 128), self._vZ2(self.dZ2), 128, n, self.H2, self.W2,
                  bsum=self.bsum, **self._nb(self.st2, self.H2 * self.W2, want_wgrad))
        if want_wgrad:
            self.c2.wgrad(self.dZ2, self.S0v, 0, self.S0v.shape[0])
        # data gradient of model.2 lands in space-to-depth order == the row order of model.0's output; the
        # LeakyReLU mask of model.0 is applied in the GEMM epilogue
        self.c2.dgrad(self.dZ2, self.dZ0v, mask=View(self.S0v, 0, 0, 0), mask_slope=0.2)
        if want_wgrad:
            be.colsum(self.dZ0, 0, 64, self.arena.view("model.0.bias", G), row_img=self.row_img0)
            self.c0.wgrad(self.dZ0, self.E0, 0, self.rows0)
        be.flush_sums()
        if dinput is not None:
            self.c0.dgrad(self.dZ0, self.dE0)
            be.col2im(self.dE0, 4, c_first, dinput.shape[1], n, self.H, self.W, 4, 2, 1, self.H1, self.W1, 2, None, dinput, accumulate)


# ==========================================================================================
# VGG-16 features[:16] (irc:642-683)
# ==========================================================================================
class VggEngine:
    """Frozen trunk.  n_img images go forward; backward (data gradient only) runs on the first
    n_bwd of them (the `fake` half when fake and target are batched together)."""

    def __init__(self, be, n_img: int, n_bwd: int, H: int, W: int, device, arena: L.ParamArena = None):
        if H % 4 or W % 4:
            raise NotImplementedError("VGG engine needs H, W multiples of 4")
        self.be, self.n, self.nb, self.H, self.W, self.dev = be, n_img, n_bwd, H, W, device
        self.arena = arena or L.ParamArena(vgg_shapes(), device)
        self.packer = L.Packer(self.arena)
        A, P = self.arena, self.packer
        res = [(H, W), (H, W), (H // 2, W // 2), (H // 2, W // 2), (H // 4, W // 4), (H // 4, W // 4), (H // 4, W // 4)]
        self.res = res
        self.act: List[L.Frame] = []
        self.convs: List[ConvOp] = []
        for i, (idx, ci, co) in enumerate(VGG_CFG):
            h, w = res[i]
            fr = L.Frame(n_img, h, w, 1, co, device)
            self.act.append(fr)
            if i == 0:
                lay = L.layout_im2col(P, A, f"features.{idx}.weight", co, ci, 3)
                taps = [0]
            else:
                lay = L.layout_std(P, A, f"features.{idx}.weight", co, ci, 3, 3)
                taps = L.taps_centered(3, 3, fr.wp)
            self.convs.append(ConvOp(be, lay, taps, A, 64, bias_name=f"features.{idx}.bias", pixels=n_img * h * w, name=f"V.{idx}"))
            self.convs[-1].flops_bwd = 2.0 * lay.param_numel * n_bwd * h * w
        # data gradient of conv1_1 (64 -> 3 input planes) as a GEMM over the kernel rows + horizontal tap reduction in the epilogue:
        # column j * 3 + c of tap r holds W[k][c][r][2 - j]  (dx[c][y][x] = sum_{r,s,k} dz[y + 1 - r][x + 1 - s][k] W[k][c][r][s])
        self.fused_dgrad0 = bool(getattr(be, "direct_smallk", False)) and n_bwd > 0
        if self.fused_dgrad0:
            idx = np.full((32, 3, 64), -1, np.int64)
            base = A.offset["features.0.weight"]
            for j in range(3):
                for c in range(3):
                    for r in range(3):
                        idx[j * 3 + c, r, :] = L.oihw_index(base, np.arange(64), c, r, 2 - j, 3, 3, 3)
            self.w0_dg = P.add(idx.reshape(32, 3 * 64))
            self.P0 = torch.zeros(8, 32, device=device)
        P.finish()
        self.pool = {1: L.Frame(n_img, H // 2, W // 2, 1, 64, device), 3: L.Frame(n_img, H // 4, W // 4, 1, 128, device)}
        # im2col operand of conv1_1: only materialised on the im2col + one-tap GEMM path
        self.E = None if getattr(be, "direct_smallk", False) else L.act_zeros(self.act[0].rows, 64, device)
        self.row_img = []
        for (h, w) in ((H, W), (H // 2, W // 2), (H // 4, W // 4)):
            ri = torch.zeros(n_img * (h + 2) * (w + 2), device=device, dtype=torch.int16)
            be.row_index(ri, n_img, h + 2, w + 2, 1, h + 1, 1, w + 1)
            self.row_img.append(ri)
        mean = torch.tensor(IMAGENET_MEAN, device=device); std = torch.tensor(IMAGENET_STD, device=device)
        self.scale = (0.5 / std).contiguous()                # ((x+1)/2 - mean)/std = x*scale + shift
        self.shift = ((0.5 - mean) / std).contiguous()
        if n_bwd:
            self.dz = [L.Frame(n_bwd, h, w, 1, co, device) for (h, w), (_, _, co) in zip(res, VGG_CFG)]
            self.dpool = {1: L.Frame(n_bwd, H // 2, W // 2, 1, 64, device), 3: L.Frame(n_bwd, H // 4, W // 4, 1, 128, device)}
            self.dE = None if self.fused_dgrad0 else L.act_zeros(self.dz[0].rows, 64, device)

    def refresh_weights(self):
        self.packer.refresh(self.be)

    def _ri(self, i):
        return self.row_img[0 if i < 2 else (1 if i < 4 else 2)]

    def forward(self, x1: torch.Tensor, x2: Optional[torch.Tensor] = None) -> L.Frame:
        """images in [-1,1], fp32 NCHW; x2 (optional) is batched after x1.  Returns the relu3_3 frame."""
        be = self.be
        n1 = x1.shape[0]
        H, W = self.H, self.W
        rows1 = self.act[0].rows_of(n1)
        parts = [(x1, n1, slice(0, rows1))] + ([] if x2 is None else [(x2, x2.shape[0], slice(rows1, None))])
        direct = getattr(be, "direct_smallk", False)
        c0 = self.convs[0]
        for x_, k_, sl in parts:
            if direct:
                # conv1_1 straight from the fp32 image (input normalisation folded into the tile fill), bias + ReLU, zero ring
                be.note = (c0.name, "fwd", c0.flops * k_ / self.n)
                be.smallk_conv_fwd(x_, None, self.scale, self.shift, k_, H, W, 3, 1, 1, 0, H, W, 1, c0.lay.w_f.t, self.act[0].t[sl], bias=c0.bias(), act=ACT_RELU)
            else:
                be.im2col(x_, None, self.scale, self.shift, k_, H, W, 3, 1, 1, 0, H, W, 1, self.E[sl])
        src = self.E
        for i, conv in enumerate(self.convs):
            fr = self.act[i]
            if not (direct and i == 0):
                # the frames that feed a max-pool (conv1_2, conv2_2) are only ever read at interior pixels: their ring rows need no
                # zeroing, which spares the epilogue the row_img lookup (21 us at the conv1_2 shape, scripts/prof_epi.py)
                conv.fwd(src, 0, fr.t, bias=conv.bias(), act=ACT_RELU, row_img=None if i in self.pool else self._ri(i))
            if i in self.pool:
                h, w = self.res[i + 1]
                be.maxpool2(fr.view(), self.pool[i].view(), fr.C, self.n, h, w)
                src = self.pool[i].t
            else:
                src = fr.t
        return self.act[-1]

    def backward(self, dfake: torch.Tensor) -> None:
        """self.dz[-1] must hold dL/d(pre-ReLU conv3_3) of the first n_bwd images; accumulates into dfake."""
        be, nb = self.be, self.nb
        for i in range(len(self.convs) - 1, 0, -1):
            conv = self.convs[i]
            dz = self.dz[i].t
            if (i - 1) in self.pool:
                # input of conv i is the pooled frame: data gradient w.r.t. it, then route through the max-pool
                # and the ReLU of conv i-1
                conv.dgrad(dz, self.dpool[i - 1].t)
                h, w = self.res[i]
                src = self.act[i - 1]
                be.maxpool2_bwd(View(src.t[:src.rows_of(nb)], 0, src.hp, src.wp, 1, 1), self.dpool[i - 1].view(), self.dz[i - 1].view(),
                                src.C, nb, h, w)
            else:
                prev = self.act[i - 1]
                conv.dgrad(dz, self.dz[i - 1].t, mask=View(prev.t[:prev.rows_of(nb)], 0, 0, 0), mask_slope=0.0)
        if self.fused_dgrad0:
            d0 = self.dz[0]
            be.note = (self.convs[0].name, "dgrad", self.convs[0].flops_bwd)
            be.conv_gemm(d0.t, 0, 64, [d0.wp, 0, -d0.wp], self.w0_dg.t, 32, self.P0,
                         tap=dict(out=dfake, nshift=3, nco=3, H=self.H, W=self.W, hp=d0.hp, wp=d0.wp, oy=1, ox=1, act=ACT_NONE, scale=self.scale, accumulate=True))
        else:
            self.convs[0].dgrad(self.dz[0].t, self.dE)
            be.col2im(self.dE, 3, 0, 3, nb, self.H, self.W, 3, 1, 1, self.H, self.W, 1, self.scale, dfake, True)
