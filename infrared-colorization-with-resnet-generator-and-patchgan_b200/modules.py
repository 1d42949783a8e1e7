"""Drop-in module surface of the hot path: the same names, constructor arguments, `forward`
signatures and `state_dict` keys as Code/ir_colorization.py (cited irc:LINE), with every
operation executed by libirc_sm100.so through the engines.

Usage is the reference's: build modules, call them on fp32 NCHW CUDA tensors, call
`loss.backward()`, step any torch optimizer.  (The fused `TrainStep` of train_step.py is the
fast path that `train_kaist` uses; these modules exist so that code written against the
reference keeps working, and so that parity can be tested module by module.)"""
from __future__ import annotations

import functools
import math
import warnings
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import layout as L
from . import engine as E
from ._native import CudaBackend, Tables, View

_BACKEND = None


def backend():
    """The process-wide CUDA backend (raises if libirc_sm100.so or an sm_100 device is missing)."""
    global _BACKEND
    if _BACKEND is None:
        _BACKEND = CudaBackend()
    return _BACKEND


def set_backend(be) -> None:
    """Test hook: lets the host-logic tests drive the modules with the torch restatement of the primitives."""
    global _BACKEND
    _BACKEND = be


# ------------------------------------------------------------------------------------------
# Config (irc:32-142): identical attribute names and defaults
# ------------------------------------------------------------------------------------------
class Config:
    def __init__(self):
        self.mode = "test"
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        self.img_size = 256
        self.input_nc = 1
        self.output_nc = 3
        self.ngf = 64
        self.norm = "instance"
        self.no_antialias = False
        self.no_antialias_up = False
        self.save_every = 5
        self.save_dir = r".\Weights\trained_w_night\checkpoints_kaist"
        self.output_dir = r".\results"
        self.test_G_weights = r".\Weights\trained_w_night\checkpoints_kaist\netG_best.pth"
        self.train_roots = [r"kaist-dataset\versions\1\set00", r"kaist-dataset\versions\1\set01",
                            r"kaist-dataset\versions\1\set03", r"kaist-dataset\versions\1\set04"]
        self.kaist_root = self.train_roots[0]
        self.batch_size = 4
        self.epochs = 50
        self.lr_G = 2e-4
        self.lr_D = 2e-4
        self.beta1 = 0.5
        self.beta2 = 0.999
        self.lambda_L1 = 30.0
        self.lambda_perc = 30.0
        self.lambda_tv = 1e-4
        self.lambda_ssim = 2.0
        self.lambda_gan = 0.1
        self.num_workers = 4
        self.val_ratio = 0.1
        self.lr_decay_start_epoch = 40
        self.init_G_weights = None
        self.test_roots = [r"kaist-dataset\versions\1\set02", r"kaist-dataset\versions\1\set05"]
        self.save_comparisons = True
        self.comparison_dirname = "Comparisons"
        self.comparison_add_text = False
        self.comparison_pad = 8
        self.comparison_font_scale = 0.6
        self.comparison_thickness = 2
        self.best50_copy_preds = True
        self.best50_copy_collages = True
        self.best50_preds_subdir = "colored"
        self.best50_collages_subdir = "collages"
        self.topk = 50
        self.best50_dirname = "Best_50_colored_images"
        # additions of this implementation (defaults keep the reference behaviour)
        self.world_size = 1           # data-parallel replicas (one process per GPU)
        self.synthetic_steps = 0      # >0: train/test on synthetic pairs (the KAIST image I/O layer is out of scope)


class Identity(nn.Module):
    """irc:148-151: pass-through layer used when normalisation is disabled"""

    def forward(self, x):
        return x


def _no_norm(num_features):
    return Identity()


def get_norm_layer(norm_type="instance"):
    """irc:154-165.  The returned object is only used as a tag by the engines: nn.InstanceNorm2d (the default graph) or - 'none' /
    None - a factory of Identity layers, which also switches off the convolution biases exactly like the reference's
    `use_bias = (norm_layer == nn.InstanceNorm2d)` (irc:452-455, :590-593); or nn.BatchNorm2d ('batch': affine parameters in the
    arena, running statistics as buffers, batch statistics in train() mode and running statistics in eval() mode)."""
    if norm_type == "instance":
        return nn.InstanceNorm2d
    if norm_type == "none" or norm_type is None:
        return _no_norm
    if norm_type == "batch":
        return nn.BatchNorm2d
    raise NotImplementedError(f"Normalization type [{norm_type}] not supported")


def _norm_tag(norm_layer) -> str:
    if norm_layer is nn.InstanceNorm2d:
        return "instance"
    if norm_layer is _no_norm:
        return "none"
    if norm_layer is nn.BatchNorm2d:
        return "batch"
    raise NotImplementedError("norm_layer must come from get_norm_layer('instance' | 'batch' | 'none')")


def get_lr_lambda(cfg):
    """irc:212-233"""
    def lr_lambda(epoch):
        e = epoch + 1
        if e <= cfg.lr_decay_start_epoch:
            return 1.0
        if e >= cfg.epochs:
            return 0.0
        return max(0.0, 1.0 - (e - cfg.lr_decay_start_epoch) / float(max(1, cfg.epochs - cfg.lr_decay_start_epoch)))
    return lr_lambda


def get_filter(filt_size=3):
    """irc:240-266 (binomial row outer product, normalised)"""
    rows = {1: [1.], 2: [1., 1.], 3: [1., 2., 1.], 4: [1., 3., 3., 1.], 5: [1., 4., 6., 4., 1.]}
    a = torch.tensor(rows[filt_size])
    f = a[:, None] * a[None, :]
    return f / f.sum()


def init_weights(net, init_type="normal", init_gain=0.02):
    """irc:168-197: conv weights ~ N(0, gain), biases 0 (only 'normal' is ever used, irc:778, :1597)"""
    if init_type != "normal":
        raise NotImplementedError("only init_type='normal' is used by the reference path")
    with torch.no_grad():
        bn_weights = {id(mod.weight) for mod in net.modules() if isinstance(mod, _BNParams)}
        for name, p in net.named_parameters():
            if id(p) in bn_weights:
                p.normal_(1.0, init_gain)                     # irc:190-194: norm layers N(1, gain)
            elif name.endswith("weight"):
                p.normal_(0.0, init_gain)
            elif name.endswith("bias"):
                p.zero_()


def init_net(net, init_type="normal", init_gain=0.02, device="cuda", initialize_weights=True):
    """irc:200-209"""
    net.to(device)
    if initialize_weights:
        init_weights(net, init_type, init_gain)
    return net


# ------------------------------------------------------------------------------------------
# parameter plumbing: nn.Parameters that alias a flat arena
# ------------------------------------------------------------------------------------------
class _ConvParams(nn.Module):
    """holds `weight` / `bias` of one convolution under the reference's key"""

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor]):
        super().__init__()
        self.weight = nn.Parameter(w)
        if b is None:
            self.register_parameter("bias", None)          # nn.Conv2d(bias=False): no `bias` key in the state_dict
        else:
            self.bias = nn.Parameter(b)


class _BNParams(nn.Module):
    """affine parameters (arena views) and running-statistics buffers of one nn.BatchNorm2d under the reference's keys"""

    def __init__(self, w: torch.Tensor, b: torch.Tensor):
        super().__init__()
        self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(b)
        self.register_buffer("running_mean", torch.zeros(w.numel(), device=w.device))
        self.register_buffer("running_var", torch.ones(w.numel(), device=w.device))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long, device=w.device))


class _Filt(nn.Module):
    """Downsample / UpsampleAA carry their binomial filter as a buffer in the state_dict (irc:289, :337)"""

    def __init__(self, channels: int):
        super().__init__()
        self.register_buffer("filt", get_filter(3)[None, None].repeat(channels, 1, 1, 1))


class _ArenaModule(nn.Module):
    """A module whose parameters all live in one ParamArena; engines are built per input shape and pooled so that
    several forward passes may be alive in the autograd graph at once (irc:1642-1643 calls D twice before
    backward)."""

    def _init_arena(self, shapes: Dict[str, tuple], device):
        self.arena = L.ParamArena(shapes, device)
        self._free = {}
        self._shapes = shapes

    def _apply(self, fn, recurse=True):
        # .to(device) / .cuda(): parameters must keep aliasing one flat arena, so move the arena and rebind
        probe = fn(torch.zeros(1, device=self.arena.device))
        if probe.device != self.arena.flat.device:
            old = self.arena
            self.arena = L.ParamArena(self._shapes, probe.device)
            self.arena.flat.copy_(old.flat)
            self._free = {}
            self._rebind()
            for m in self.modules():
                for k, b in list(m._buffers.items()):
                    if b is not None:
                        m._buffers[k] = fn(b)
            return self
        return super()._apply(fn, recurse)

    def _named_holders(self):
        raise NotImplementedError

    def _rebind(self):
        for key, holder in self._named_holders():
            holder.weight = nn.Parameter(self.arena.view(key + ".weight"))
            if key + ".bias" in self.arena.offset:
                holder.bias = nn.Parameter(self.arena.view(key + ".bias"))

    def _bn_state(self):
        """{state_dict prefix: (running_mean, running_var)} of the BatchNorm holders, handed to the engine for each forward pass"""
        return {name: (h.running_mean, h.running_var) for name, h in self.named_modules() if isinstance(h, _BNParams)}

    def _bn_tick(self, calls: int = 1):
        for h in self.modules():
            if isinstance(h, _BNParams):
                h.num_batches_tracked += calls

    def _acquire(self, key, make):
        pool = self._free.setdefault(key, [])
        return pool.pop() if pool else make()

    def _release(self, key, eng):
        self._free.setdefault(key, []).append(eng)

    def _params(self):
        return [p for _, p in sorted(self.named_parameters(), key=lambda kv: self.arena.offset[kv[0]])]

    def _grads_out(self, needs):
        """clone the arena gradients in parameter order"""
        g = self.arena.grads()
        return [g[k].clone() if need else None for (k, need) in needs]


def _param_order(mod: _ArenaModule):
    return sorted((k for k, _ in mod.named_parameters()), key=lambda k: mod.arena.offset[k])


# ------------------------------------------------------------------------------------------
# Downsample / UpsampleAA (irc:269-355), stand-alone on fp32 NCHW
# ------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=64)
def _stencil_tables(kind: str, H: int, W: int, device: str):
    mat = L.down_matrix if kind == "down" else L.up_matrix
    fwd = L.make_tables(mat(H), mat(W), device)
    bwd = L.make_tables(mat(H).T, mat(W).T, device)
    return fwd, bwd


class _StencilFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind):
        be = backend()
        x = x.contiguous().float()
        n, c, h, w = x.shape
        fwd, bwd = _stencil_tables(kind, h, w, str(x.device))
        ho, wo = ((h + 1) // 2, (w + 1) // 2) if kind == "down" else (2 * h, 2 * w)
        out = torch.empty(n, c, ho, wo, device=x.device)
        be.stencil_nchw(x, out, fwd)
        ctx.bwd, ctx.shape = bwd, (n, c, h, w)
        return out

    @staticmethod
    def backward(ctx, g):
        dx = torch.empty(ctx.shape, device=g.device)
        backend().stencil_nchw(g.contiguous().float(), dx, ctx.bwd)
        return dx, None


class Downsample(_Filt):
    """irc:269-310: reflect-pad + depthwise binomial 3x3, stride 2"""

    def __init__(self, channels, pad_type="reflect", filt_size=3, stride=2, pad_off=0):
        if (pad_type, filt_size, stride, pad_off) != ("reflect", 3, 2, 0):
            raise NotImplementedError("only the default Downsample(channels) configuration is built")
        super().__init__(channels)
        self.channels = channels

    def forward(self, x):
        return _StencilFn.apply(x, "down")


class UpsampleAA(_Filt):
    """irc:313-355: bilinear x2 (align_corners=True) + reflect-pad + depthwise binomial 3x3"""

    def __init__(self, channels, filt_size=3, stride=2, pad_type="reflect"):
        if (filt_size, stride, pad_type) != (3, 2, "reflect"):
            raise NotImplementedError("only the default UpsampleAA(channels) configuration is built")
        super().__init__(channels)
        self.channels = channels

    def forward(self, x):
        return _StencilFn.apply(x, "up")


# ------------------------------------------------------------------------------------------
# generator (irc:425-569)
# ------------------------------------------------------------------------------------------
class _GenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, *params):
        B, _, H, W = x.shape
        # needs_input_grad stays True for parameters under torch.no_grad() and forward() itself always runs with grad mode
        # off, so the module records the caller's grad mode (mod._grad_on); without it the engine would never return to
        # the pool in inference and be rebuilt (0.3 s of host work) on every call
        grad = mod._grad_on and any(ctx.needs_input_grad)
        key = (B, H, W, grad)                  # inference engines carry no backward buffers
        eng = mod._acquire(key, lambda: E.GeneratorEngine(backend(), B, H, W, x.device, arena=mod.arena, training=grad,
                                                          no_antialias_up=mod.no_antialias_up, no_antialias=mod.no_antialias, norm=mod.norm))
        eng.refresh_weights()
        if mod.norm == "batch":
            # train(): batch statistics + running-statistics update (irc:1622); eval(): running statistics (irc:1357, :1527)
            eng.bn_state, eng.bn_training, eng.bn_updates = mod._bn_state(), mod.training, 1
            if mod.training:
                mod._bn_tick()
        out = eng.forward(x.contiguous().float()).clone()
        if grad:
            ctx.mod, ctx.eng, ctx.key = mod, eng, key
            ctx.needs = [(k, p.requires_grad) for k, p in zip(_param_order(mod), params)]
        else:
            mod._release(key, eng)
        return out

    @staticmethod
    def backward(ctx, g):
        ctx.eng.backward(g.contiguous().float())
        grads = ctx.mod._grads_out(ctx.needs)
        ctx.mod._release(ctx.key, ctx.eng)
        return (None, None, *grads)


class _ResBlockHolder(nn.Module):
    def __init__(self, arena, b):
        super().__init__()
        bias = lambda j: arena.view(f"resblocks.{b}.conv_block.{j}.bias") if f"resblocks.{b}.conv_block.{j}.bias" in arena.offset else None
        self.conv_block = nn.ModuleDict({str(j): _ConvParams(arena.view(f"resblocks.{b}.conv_block.{j}.weight"), bias(j)) for j in (1, 5)})
        for j in (2, 6):                       # nn.BatchNorm2d behind each convolution (norm='batch')
            if f"resblocks.{b}.conv_block.{j}.weight" in arena.offset:
                self.conv_block[str(j)] = _BNParams(arena.view(f"resblocks.{b}.conv_block.{j}.weight"), arena.view(f"resblocks.{b}.conv_block.{j}.bias"))


class ResnetUNetGenerator(_ArenaModule):
    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=nn.InstanceNorm2d, use_dropout=False, n_blocks=9,
                 padding_type="reflect", no_antialias=False, no_antialias_up=False):
        super().__init__()
        assert n_blocks >= 0                                          # irc:449
        self.norm = _norm_tag(norm_layer)
        if (input_nc, output_nc, ngf) != (1, 3, 64) or use_dropout or padding_type != "reflect":
            raise NotImplementedError("built: the generator graph 1->3, ngf 64, reflect padding, InstanceNorm or no normalisation; "
                                      "anti-aliased (Downsample) or - no_antialias=True - stride-2 down-sampling, UpsampleAA or "
                                      "- no_antialias_up=True - ConvTranspose2d up-sampling; dropout / BatchNorm are not (SURVEY.md §8f-4)")
        self.n_blocks = n_blocks
        self.no_antialias_up = bool(no_antialias_up)
        self.no_antialias = bool(no_antialias)
        dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self._init_arena(E.generator_shapes(input_nc, output_nc, ngf, n_blocks, self.no_antialias_up, self.norm), dev)
        A = self.arena
        hold = lambda k: _ConvParams(A.view(k + ".weight"), A.view(k + ".bias") if k + ".bias" in A.offset else None)
        bn = lambda k: {k.rsplit(".", 1)[1]: _BNParams(A.view(k + ".weight"), A.view(k + ".bias"))} if self.norm == "batch" else {}
        self.inc = nn.ModuleDict({"1": hold("inc.1"), **bn("inc.2")})
        self.down1 = nn.ModuleDict({"0": hold("down1.0"), **bn("down1.1")})
        self.down1_down = None if self.no_antialias else _Filt(2 * ngf)          # irc:474: no blur module, no `filt` buffer
        self.down2 = nn.ModuleDict({"0": hold("down2.0"), **bn("down2.1")})
        self.down2_down = None if self.no_antialias else _Filt(4 * ngf)          # irc:482
        self.resblocks = nn.ModuleList([_ResBlockHolder(A, b) for b in range(n_blocks)])
        # UpsampleAA carries only its `filt` buffer; nn.ConvTranspose2d (irc:495-499, :512-516) carries weight (Cin, Cout, 3, 3) + bias
        self.up1_up = hold("up1_up") if self.no_antialias_up else _Filt(4 * ngf)
        self.up1_conv = nn.ModuleDict({"0": hold("up1_conv.0"), **bn("up1_conv.1")})
        self.up2_up = hold("up2_up") if self.no_antialias_up else _Filt(2 * ngf)
        self.up2_conv = nn.ModuleDict({"0": hold("up2_conv.0"), **bn("up2_conv.1")})
        self.outc = nn.ModuleDict({"1": hold("outc.1")})

    def _named_holders(self):
        for name, m in self.named_modules():
            if isinstance(m, (_ConvParams, _BNParams)):
                yield name, m

    def forward(self, x, layers=None, encode_only=False):
        """irc:533-569: returns (image, None).  `encode_only` feature taps are not part of the hot path."""
        if encode_only or layers:
            raise NotImplementedError("encode_only / layers feature taps are never used by the reference's train/test path")
        self._grad_on = torch.is_grad_enabled()
        return _GenFn.apply(self, x, *self._params()), None


# ------------------------------------------------------------------------------------------
# ResnetBlock (irc:362-418), stand-alone
# ------------------------------------------------------------------------------------------
class _ResBlockEngine:
    def __init__(self, be, B, H, W, dim, device, arena):
        self.be, self.B, self.H, self.W, self.C = be, B, H, W, dim
        self.packer = L.Packer(arena)
        F = lambda: L.Frame(B, H, W, 1, dim, device)
        self.X, self.Za, self.Hh, self.Zb, self.Y = F(), F(), F(), F(), F()
        self.dY, self.dZb, self.Gh, self.dZa, self.Gx, self.dX = F(), F(), F(), F(), F(), F()
        self.sta, self.stb, self.bsum = (torch.zeros(B, dim, 2, device=device) for _ in range(3))
        taps = L.taps_centered(3, 3, self.X.wp)
        self.c1 = E.ConvOp(be, L.layout_std(self.packer, arena, "conv_block.1.weight", dim, dim, 3, 3), taps, arena, self.X.rows, pixels=B * H * W)
        self.c2 = E.ConvOp(be, L.layout_std(self.packer, arena, "conv_block.5.weight", dim, dim, 3, 3), taps, arena, self.X.rows, pixels=B * H * W)
        self.packer.finish()
        self.fold = L.make_tables(L.fold_matrix(H, 1), L.fold_matrix(W, 1), device)
        self.nchw_in = torch.zeros(B, H, W, dim, device=device)

    def forward(self, x):
        be, B, H, W, C = self.be, self.B, self.H, self.W, self.C
        self.packer.refresh(be)
        # NCHW fp32 -> reflect-padded NHWC frame (layout change only; torch copies are plumbing)
        xi = torch.nn.functional.pad(x, (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1)
        self.X.t.copy_(xi.reshape(-1, C))
        n = H * W
        self.c1.fwd(self.X.t, 0, self.Za.t)
        be.in_stats(self.Za.view(), C, B, H, W, self.sta)
        be.gather(self.Za.view(), self.Hh.view(), C, B, H, W, 1, 1, stats=self.sta, cnt=n, eps=E.EPS, act=E.ACT_RELU)
        self.c2.fwd(self.Hh.t, 0, self.Zb.t)
        be.in_stats(self.Zb.view(), C, B, H, W, self.stb)
        be.gather(self.Zb.view(), self.Y.view(), C, B, H, W, 1, 1, stats=self.stb, cnt=n, eps=E.EPS, act=E.ACT_NONE, res=self.X.view())
        return self._interior(self.Y)

    def _interior(self, fr):
        return fr.t.view(self.B, fr.hp, fr.wp, self.C)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().contiguous()

    def backward(self, g):
        be, B, H, W, C = self.be, self.B, self.H, self.W, self.C
        n = H * W
        self.dY.t.view(B, self.dY.hp, self.dY.wp, C)[:, 1:-1, 1:-1].copy_(g.permute(0, 2, 3, 1))
        be.in_bwd(self.Zb.view(), self.dY.view(), self.dZb.view(), C, B, H, W, stats=self.stb, cnt=n, eps=E.EPS, act=E.ACT_NONE, bsum=self.bsum)
        self.c2.wgrad(self.dZb.t, self.Hh.t, 0, self.dZb.rows)
        self.c2.dgrad(self.dZb.t, self.Gh.t)
        be.in_bwd(self.Za.view(), self.Gh.pview(), self.dZa.view(), C, B, H, W, stats=self.sta, cnt=n, eps=E.EPS, act=E.ACT_RELU,
                  tables=self.fold, bsum=self.bsum)
        self.c1.wgrad(self.dZa.t, self.X.t, 0, self.dZa.rows)
        be.flush_sums()
        self.c1.dgrad(self.dZa.t, self.Gx.t)
        be.gather(self.Gx.pview(), self.dX.view(), C, B, H, W, 1, 0, tables=self.fold, res=self.dY.view())
        return self._interior(self.dX)


class _ResFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, *params):
        B, C, H, W = x.shape
        key = (B, H, W)
        eng = mod._acquire(key, lambda: _ResBlockEngine(backend(), B, H, W, C, x.device, mod.arena))
        out = eng.forward(x.contiguous().float())
        if mod._grad_on and any(ctx.needs_input_grad):
            ctx.mod, ctx.eng, ctx.key = mod, eng, key
            ctx.needs = [(k, p.requires_grad) for k, p in zip(_param_order(mod), params)]
        else:
            mod._release(key, eng)
        return out

    @staticmethod
    def backward(ctx, g):
        dx = ctx.eng.backward(g.contiguous().float())
        grads = ctx.mod._grads_out(ctx.needs)
        ctx.mod._release(ctx.key, ctx.eng)
        return (None, dx, *grads)


class ResnetBlock(_ArenaModule):
    """irc:362-418 with reflect padding and InstanceNorm: x + IN(conv(pad(ReLU(IN(conv(pad(x)))))))"""

    def __init__(self, dim, padding_type="reflect", norm_layer=nn.InstanceNorm2d, use_dropout=False, use_bias=True):
        super().__init__()
        if padding_type != "reflect" or norm_layer is not nn.InstanceNorm2d or use_dropout or dim % 64:
            raise NotImplementedError("only ResnetBlock(dim % 64 == 0, 'reflect', InstanceNorm2d, no dropout) is built")
        dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        shapes = {"conv_block.1.weight": (dim, dim, 3, 3), "conv_block.1.bias": (dim,),
                  "conv_block.5.weight": (dim, dim, 3, 3), "conv_block.5.bias": (dim,)}
        self._init_arena(shapes, dev)
        A = self.arena
        self.conv_block = nn.ModuleDict({str(j): _ConvParams(A.view(f"conv_block.{j}.weight"), A.view(f"conv_block.{j}.bias")) for j in (1, 5)})

    def _named_holders(self):
        for j in ("1", "5"):
            yield f"conv_block.{j}", self.conv_block[j]

    def forward(self, x):
        self._grad_on = torch.is_grad_enabled()
        return _ResFn.apply(self, x, *self._params())


# ------------------------------------------------------------------------------------------
# discriminator (irc:576-635)
# ------------------------------------------------------------------------------------------
class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, *params):
        n, _, H, W = x.shape
        key = (n, H, W)
        eng = mod._acquire(key, lambda: E.DiscriminatorEngine(backend(), n, H, W, x.device, arena=mod.arena, norm=mod.norm))
        eng.refresh_weights()
        if mod.norm == "batch":
            eng.bn_state, eng.bn_training = mod._bn_state(), mod.training
            if mod.training:
                mod._bn_tick()
        xf = x.float()
        out = eng.forward(xf[:, 0:1].contiguous(), xf[:, 1:4].contiguous()).clone()
        need_p = any(ctx.needs_input_grad[2:])
        if mod._grad_on and any(ctx.needs_input_grad):
            ctx.mod, ctx.eng, ctx.key, ctx.need_p, ctx.need_x = mod, eng, key, need_p, ctx.needs_input_grad[1]
            ctx.needs = [(k, p.requires_grad) for k, p in zip(_param_order(mod), params)]
            ctx.xshape = x.shape
        else:
            mod._release(key, eng)
        return out

    @staticmethod
    def backward(ctx, g):
        dx = torch.zeros(ctx.xshape, device=g.device) if ctx.need_x else None
        ctx.eng.backward(g.contiguous().float(), ctx.need_p, dx, c_first=0, accumulate=False)
        grads = ctx.mod._grads_out(ctx.needs) if ctx.need_p else [None] * len(ctx.needs)
        ctx.mod._release(ctx.key, ctx.eng)
        return (None, dx, *grads)


class NLayerDiscriminator(_ArenaModule):
    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.InstanceNorm2d):
        super().__init__()
        self.norm = _norm_tag(norm_layer)
        if (input_nc, ndf, n_layers) != (4, 64, 3):
            raise NotImplementedError("only NLayerDiscriminator(4, 64, 3, ...) (irc:1590-1595) is built")
        dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self._init_arena(E.discriminator_shapes(input_nc, ndf, self.norm), dev)
        A = self.arena
        self.model = nn.ModuleDict({str(i): _ConvParams(A.view(f"model.{i}.weight"), A.view(f"model.{i}.bias") if f"model.{i}.bias" in A.offset else None)
                                    for i in (0, 2, 5, 8, 11)})
        if self.norm == "batch":
            for i in (3, 6, 9):
                self.model[str(i)] = _BNParams(A.view(f"model.{i}.weight"), A.view(f"model.{i}.bias"))

    def _named_holders(self):
        for i in self.model:
            yield f"model.{i}", self.model[i]

    def forward(self, x):
        self._grad_on = torch.is_grad_enabled()
        return _DiscFn.apply(self, x, *self._params())


# ------------------------------------------------------------------------------------------
# VGG perceptual trunk (irc:642-683)
# ------------------------------------------------------------------------------------------
class _VggFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x):
        n, _, H, W = x.shape
        key = (n, H, W)
        eng = mod._acquire(key, lambda: E.VggEngine(backend(), n, n, H, W, x.device, arena=mod.arena))
        eng.refresh_weights()
        fr = eng.forward(x.contiguous().float())
        out = fr.t.view(n, fr.hp, fr.wp, fr.C)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().contiguous()
        if mod._grad_on and ctx.needs_input_grad[1]:
            ctx.mod, ctx.eng, ctx.key, ctx.shape = mod, eng, key, x.shape
            ctx.save_for_backward(out)
        else:
            mod._release(key, eng)
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        eng = ctx.eng
        dz = eng.dz[-1]
        gz = (g * (out > 0)).permute(0, 2, 3, 1)                       # through relu3_3
        dz.t.view(eng.nb, dz.hp, dz.wp, dz.C)[:, 1:-1, 1:-1].copy_(gz)
        dx = torch.zeros(ctx.shape, device=g.device)
        eng.backward(dx)
        ctx.mod._release(ctx.key, eng)
        return None, dx


class VGGPerceptual(_ArenaModule):
    """Frozen torchvision VGG-16 features[:16].  The ImageNet weights (vgg16-397923af.pth) are a third-party
    artefact: they are loaded when torchvision can find them locally, otherwise the trunk keeps torchvision's
    random initialisation and a warning is issued (the reference would fail to download, irc:659)."""

    def __init__(self, device):
        super().__init__()
        self._init_arena(E.vgg_shapes(), torch.device(device))
        A = self.arena
        self.features = nn.ModuleDict({str(i): _ConvParams(A.view(f"features.{i}.weight"), A.view(f"features.{i}.bias")) for i, _, _ in E.VGG_CFG})
        self.register_buffer("mean", torch.tensor(E.IMAGENET_MEAN, device=device).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor(E.IMAGENET_STD, device=device).view(1, 3, 1, 1))
        loaded = False
        try:
            import torchvision
            try:
                tv = torchvision.models.vgg16(weights=torchvision.models.VGG16_Weights.IMAGENET1K_V1)
                loaded = True
            except Exception:
                tv = torchvision.models.vgg16(weights=None)
            sd = {f"features.{k}": v for k, v in tv.features.state_dict().items() if int(k.split(".")[0]) < 16}
            A.load(sd)
        except Exception:
            with torch.no_grad():
                for k in A.names:
                    v = A.view(k)
                    if k.endswith("weight"):
                        v.normal_(0.0, math.sqrt(2.0 / (v.shape[0] * 9)))
                    else:
                        v.zero_()
        if not loaded:
            warnings.warn("VGG-16 ImageNet weights not available offline: the perceptual trunk is randomly initialised")
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    def _named_holders(self):
        for i, _, _ in E.VGG_CFG:
            yield f"features.{i}", self.features[str(i)]

    def forward(self, x):
        self._grad_on = torch.is_grad_enabled()
        return _VggFn.apply(self, x)


# ------------------------------------------------------------------------------------------
# losses (irc:686-750)
# ------------------------------------------------------------------------------------------
class _TvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        be = backend()
        x = x.contiguous().float()
        n, c, h, w = x.shape
        sums = torch.zeros(3, device=x.device)
        d = torch.empty_like(x)
        cv, ch = n * c * (h - 1) * w, n * c * h * (w - 1)
        be.pixel_loss(x, None, 0.0, 1.0 / cv, 1.0 / ch, sums, d)
        ctx.save_for_backward(d)
        return sums[1] / cv + sums[2] / ch

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return d * g


def tv_loss(x):
    """irc:686-694"""
    return _TvFn.apply(x)


class _SsimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2, size_average):
        from .train_step import gaussian_window
        be = backend()
        a, b = img1.contiguous().float(), img2.contiguous().float()
        n, c, h, w = a.shape
        win = gaussian_window()
        sums = torch.zeros(n, device=a.device)
        ga, gb, gc = (torch.empty_like(a) for _ in range(3))
        be.ssim_fwd(a, b, 1.0, 0.0, win, sums, ga, gb, gc)
        ctx.save_for_backward(a, b, ga, gb, gc)
        ctx.win, ctx.size_average = win, size_average
        per = sums / (c * h * w)
        return 1.0 - (per.mean() if size_average else per)

    @staticmethod
    def backward(ctx, g):
        a, b, ga, gb, gc = ctx.saved_tensors
        n, c, h, w = a.shape
        d = torch.empty_like(a)
        if ctx.size_average:
            backend().ssim_bwd(a, b, 1.0, 0.0, ctx.win, ga, gb, gc, -1.0 / a.numel(), d, False)
            return d * g, None, None
        backend().ssim_bwd(a, b, 1.0, 0.0, ctx.win, ga, gb, gc, -1.0 / (c * h * w), d, False)
        return d * g.view(-1, 1, 1, 1), None, None


def ssim_loss_torch(img1, img2, window_size=11, size_average=True):
    """irc:714-750 (differentiable w.r.t. img1, which is all the reference needs, irc:1675-1677)"""
    assert img1.shape == img2.shape, "SSIM images must have the same shape"
    if window_size != 11:
        raise NotImplementedError("only the default 11-tap window is built")
    return _SsimFn.apply(img1, img2.detach(), size_average)


# ------------------------------------------------------------------------------------------
# model wrapper (irc:757-796)
# ------------------------------------------------------------------------------------------
class IRColorizationModel(nn.Module):
    def __init__(self, cfg: Config):
        super().__init__()
        norm_layer = get_norm_layer(cfg.norm)
        self.netG = ResnetUNetGenerator(cfg.input_nc, cfg.output_nc, cfg.ngf, norm_layer=norm_layer, use_dropout=False, n_blocks=9,
                                        padding_type="reflect", no_antialias=cfg.no_antialias, no_antialias_up=cfg.no_antialias_up)
        self.device = torch.device(cfg.device)
        self.netG = init_net(self.netG, init_type="normal", init_gain=0.02, device=self.device, initialize_weights=True)

    def load_weights(self, path):
        state = torch.load(path, map_location=self.device)
        if isinstance(state, dict) and "state_dict" in state:
            state = state["state_dict"]
        self.netG.load_state_dict(state, strict=False)

    def forward(self, ir_tensor):
        fake_b, _ = self.netG(ir_tensor)
        return fake_b
