"""Host-side layout logic of the hot path: frame geometry, separable stencil tables, the flat
parameter arena and the index maps that pack OIHW parameters into GEMM operands.

Pure index arithmetic (numpy / torch on the host); every FLOP of the path runs in
libirc_sm100.so.  Reference citations: irc = /root/reference/Code/ir_colorization.py."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from ._native import Tables, View


# Storage type of activation frames and packed weights.  The CUDA kernels only take bf16; the
# host-logic tests flip this to float32 (with the torch restatement of the primitives) to check
# the plans exactly.
ACT_DTYPE = torch.bfloat16


def act_zeros(rows: int, cols: int, device) -> torch.Tensor:
    return torch.zeros(rows, cols, device=device, dtype=ACT_DTYPE)


# ----------------------------------------------------------------------------------------
# frames
# ----------------------------------------------------------------------------------------
class Frame:
    """NHWC bf16 activation buffer [N][H+2p][W+2p][C] with its padding ring stored."""

    def __init__(self, N: int, H: int, W: int, p: int, C: int, device):
        self.N, self.H, self.W, self.p, self.C = N, H, W, p, C
        self.hp, self.wp = H + 2 * p, W + 2 * p
        self.rows = N * self.hp * self.wp
        self.t = act_zeros(self.rows, C, device)

    def view(self, chan_off: int = 0) -> View:
        """pixel coordinates of the un-padded image"""
        return View(self.t, chan_off, self.hp, self.wp, self.p, self.p)

    def pview(self, chan_off: int = 0) -> View:
        """pixel coordinates of the padded frame"""
        return View(self.t, chan_off, self.hp, self.wp, 0, 0)

    def rows_of(self, n_img: int) -> int:
        return n_img * self.hp * self.wp


def taps_centered(kh: int, kw: int, wp: int) -> List[int]:
    return [(r - kh // 2) * wp + (s - kw // 2) for r in range(kh) for s in range(kw)]


def taps_topleft(kh: int, kw: int, wp: int) -> List[int]:
    return [r * wp + s for r in range(kh) for s in range(kw)]


# ----------------------------------------------------------------------------------------
# separable stencil operators (Downsample irc:269-310, UpsampleAA irc:313-355, reflection folds)
# ----------------------------------------------------------------------------------------
def _reflect(i: int, n: int) -> int:
    i = abs(i)
    return 2 * (n - 1) - i if i > n - 1 else i


def blur_matrix(n: int) -> np.ndarray:
    """reflect-pad 1 + [1,2,1]/4, stride 1 (irc:353-354) as an n x n matrix"""
    m = np.zeros((n, n))
    for i in range(n):
        for a, w in enumerate((0.25, 0.5, 0.25)):
            m[i, _reflect(i - 1 + a, n)] += w
    return m


def down_matrix(n: int) -> np.ndarray:
    """reflect-pad 1 + [1,2,1]/4, stride 2 (irc:307-310) as a ceil(n/2) x n matrix"""
    no = (n + 1) // 2
    m = np.zeros((no, n))
    for i in range(no):
        for a, w in enumerate((0.25, 0.5, 0.25)):
            m[i, _reflect(2 * i - 1 + a, n)] += w
    return m


def bilinear_matrix(n: int) -> np.ndarray:
    """F.interpolate(scale_factor=2, mode='bilinear', align_corners=True) (irc:351-352), 2n x n"""
    m = np.zeros((2 * n, n))
    scale = np.float32(n - 1) / np.float32(2 * n - 1) if n > 1 else np.float32(0)
    for o in range(2 * n):
        s = np.float32(scale * np.float32(o))
        i0 = min(int(np.floor(s)), n - 1)
        i1 = min(i0 + 1, n - 1)
        l1 = float(np.float32(s - np.float32(i0)))
        m[o, i0] += 1.0 - l1
        m[o, i1] += l1
    return m


def resize_matrix(m: int, n: int) -> np.ndarray:
    """F.interpolate(size=m, mode='bilinear', align_corners=True) of an axis of n samples (irc:555-556, :562-563), m x n;
    float32 index arithmetic as in ATen's area_pixel_compute_source_index"""
    a = np.zeros((m, n))
    scale = np.float32(n - 1) / np.float32(m - 1) if m > 1 else np.float32(0)
    for o in range(m):
        s = np.float32(scale * np.float32(o))
        i0 = min(int(np.floor(s)), n - 1)
        i1 = min(i0 + 1, n - 1)
        l1 = float(np.float32(s - np.float32(i0)))
        a[o, i0] += 1.0 - l1
        a[o, i1] += l1
    return a


def up_matrix(n: int) -> np.ndarray:
    return blur_matrix(2 * n) @ bilinear_matrix(n)


def fold_matrix(n: int, p: int) -> np.ndarray:
    """transpose of ReflectionPad2d(p) on one axis: n x (n+2p); rows index the un-padded axis"""
    m = np.zeros((n, n + 2 * p))
    for Y in range(n + 2 * p):
        m[_reflect(Y - p, n), Y] += 1.0
    return m


def to_ell(m: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    k = int((m != 0).sum(1).max())
    idx = np.zeros((m.shape[0], k), np.int32)
    w = np.zeros((m.shape[0], k), np.float32)
    for i in range(m.shape[0]):
        nz = np.nonzero(m[i])[0]
        idx[i, :len(nz)] = nz
        w[i, :len(nz)] = m[i, nz]
    return idx, w


def make_tables(my, mx, device) -> Tables:
    """either operator may be None (= identity along that axis)"""
    f = lambda a: torch.from_numpy(a).to(device).contiguous()
    iy, wy = (None, None) if my is None else map(f, to_ell(my))
    ix, wx = (None, None) if mx is None else map(f, to_ell(mx))
    return Tables(iy, wy, ix, wx)


# ----------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------
class ParamArena:
    """All parameters of one network in one flat fp32 buffer (plus grad / Adam moments), exposed as
    OIHW views under the reference's state_dict keys."""

    def __init__(self, shapes: Dict[str, Sequence[int]], device):
        self.names = list(shapes)
        self.shapes = {k: tuple(shapes[k]) for k in self.names}
        self.offset: Dict[str, int] = {}
        off = 0
        for k in self.names:
            self.offset[k] = off
            off += (int(np.prod(self.shapes[k])) + 3) // 4 * 4
        self.size = off
        self.device = device
        self.flat = torch.zeros(off, device=device)
        self.grad = torch.zeros(off, device=device)
        self.m = torch.zeros(off, device=device)
        self.v = torch.zeros(off, device=device)

    def numel(self, k: str) -> int:
        return int(np.prod(self.shapes[k]))

    def view(self, k: str, of: torch.Tensor = None) -> torch.Tensor:
        of = self.flat if of is None else of
        return of[self.offset[k]:self.offset[k] + self.numel(k)].view(self.shapes[k])

    def load(self, params: Dict[str, torch.Tensor]) -> None:
        for k in self.names:
            self.view(k).copy_(params[k].to(self.device, torch.float32))

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: self.view(k) for k in self.names}

    def grads(self) -> Dict[str, torch.Tensor]:
        return {k: self.view(k, self.grad) for k in self.names}


class PackedOperand:
    """A bf16 GEMM operand [rows, cols] filled from the arena through an index map."""

    def __init__(self, index: np.ndarray, std=None):
        self.index = index.astype(np.int64)     # [rows, cols] arena indices, -1 = zero
        self.rows, self.cols = index.shape
        self.std = std                          # ("f" | "d", src offset, N, K, T, kd) for the operands of a standard layer
        self.t: torch.Tensor = None             # set by Packer.finish


class Packer:
    """Collects the packed operands of one network.  Refresh = one irc_pack_bf16 launch through an index map for the irregular
    operands + one irc_pack_std launch (shared-memory transposes, no map) for the two operands of every standard convolution."""

    def __init__(self, arena: ParamArena):
        self.arena = arena
        self.ops: List[PackedOperand] = []

    def add(self, index: np.ndarray, std=None) -> PackedOperand:
        op = PackedOperand(index, std)
        self.ops.append(op)
        return op

    def finish(self) -> None:
        dev = self.arena.device
        order = [op for op in self.ops if op.std is None] + [op for op in self.ops if op.std is not None]      # map-packed operands first
        offs, total = {}, 0
        for op in order:
            offs[id(op)] = total
            total += (op.rows * op.cols + 63) // 64 * 64
            if op.std is None:
                self.n_mapped = total
        if not any(op.std is None for op in order):
            self.n_mapped = 0
        self._offs, self._order, self._full_map = offs, order, None
        m = np.full(self.n_mapped, -1, np.int64)
        for op in order:
            if op.std is None:
                o = offs[id(op)]
                m[o:o + op.rows * op.cols] = op.index.reshape(-1)
        assert m.size == 0 or m.max() < 2 ** 31
        self.map = torch.from_numpy(m.astype(np.int32)).to(dev)
        self.packed = torch.zeros(total, device=dev, dtype=ACT_DTYPE)
        for op in order:
            o = offs[id(op)]
            op.t = self.packed[o:o + op.rows * op.cols].view(op.rows, op.cols)
        # job table of the structured launch: one record per standard layer (its forward and data-gradient operands)
        layers = {}
        for op in order:
            if op.std is not None:
                kind, src, N, K, T, kd = op.std
                layers.setdefault((src, N, K, T, kd), {})[kind] = offs[id(op)]
        recs, first = [], 0
        for (src, N, K, T, kd), d in layers.items():
            nb = (N // 16) * (K // 64)
            recs.append(np.array([src, d["f"], d["d"]], np.int64).tobytes() + np.array([N, K, T, kd, first, nb], np.int32).tobytes())
            first += nb
        self.n_jobs, self.job_blocks = len(recs), first
        self.max_taps = max((k[3] for k in layers), default=0)
        self.jobs = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(dev) if recs else None

    def refresh(self, be) -> None:
        if self.n_jobs and hasattr(be, "pack_std") and self.packed.dtype == torch.bfloat16:
            if self.n_mapped:
                be.pack_bf16(self.arena.flat, self.map, self.packed[:self.n_mapped])
            be.pack_std(self.arena.flat, self.packed, self.jobs, self.n_jobs, self.job_blocks, self.max_taps)
            return
        if self._full_map is None:                  # backends without the structured kernel: everything through the map
            m = np.full(self.packed.numel(), -1, np.int64)
            for op in self._order:
                o = self._offs[id(op)]
                m[o:o + op.rows * op.cols] = op.index.reshape(-1)
            self._full_map = torch.from_numpy(m.astype(np.int32)).to(self.arena.device)
        be.pack_bf16(self.arena.flat, self._full_map, self.packed)


def oihw_index(base: int, co, ci, r, s, Cin: int, KH: int, KW: int):
    return base + ((co * Cin + ci) * KH + r) * KW + s


class WeightLayout:
    """Index maps of one convolution's weight in the three GEMM roles.

    fwd_index[n, t, k]: arena index of the weight multiplying input channel-slot k of tap t for
    output n (-1 = structural zero).  From it:
      forward operand   W_f[n, t*K + k]
      data-grad operand W_d[k, t*N + n]      (used with negated tap shifts)
      weight-grad partial layout [n][t][k]   (what irc_tn_gemm writes), inverted by `unpack`."""

    def __init__(self, packer: Packer, fwd_index: np.ndarray, param_offset: int, param_numel: int, n_pad: int = None,
                 k_pad_d: int = None, std: bool = False):
        N, T, K = fwd_index.shape
        self.N, self.T, self.K = N, T, K
        n_pad = n_pad or N
        f = np.full((n_pad, T, K), -1, np.int64); f[:N] = fwd_index
        kd = k_pad_d or N                      # reduction width of the data-grad GEMM (multiple of 64)
        # `std`: the plain OIHW -> [n][t][k] / [k][t][n] permutation of a standard layer, packed without an index map
        std = std and n_pad == N and kd == N and N % 16 == 0 and K % 64 == 0 and T <= 16
        self.w_f = packer.add(f.reshape(n_pad, T * K), std=("f", param_offset, N, K, T, kd) if std else None)
        d = np.full((K, T, kd), -1, np.int64); d[:, :, :N] = np.transpose(fwd_index, (2, 1, 0))
        self.w_d = packer.add(d.reshape(K, T * kd), std=("d", param_offset, N, K, T, kd) if std else None)
        self.kd = kd
        # inverse map: OIHW element -> position in the [N][T][K] partial
        inv = np.full(param_numel, -1, np.int64)
        flat = fwd_index.reshape(-1)
        ok = flat >= 0
        inv[flat[ok] - param_offset] = np.nonzero(ok)[0]
        assert (inv >= 0).all(), "every weight element must appear exactly once in the forward operand"
        self.unpack = torch.from_numpy(inv.astype(np.int32)).to(packer.arena.device)
        self.param_offset, self.param_numel = param_offset, param_numel


def layout_std(packer, arena, name, Cout, Cin, KH, KW) -> WeightLayout:
    co, r, s, ci = np.meshgrid(np.arange(Cout), np.arange(KH), np.arange(KW), np.arange(Cin), indexing="ij")
    idx = oihw_index(arena.offset[name], co, ci, r, s, Cin, KH, KW).reshape(Cout, KH * KW, Cin)
    return WeightLayout(packer, idx, arena.offset[name], arena.numel(name), std=True)


def layout_s2d(packer, arena, name, Cout, Cin) -> WeightLayout:
    """4x4 stride-2 conv as a 2x2 conv over 2x2 space-to-depth blocks: tap (a,b), slot (dy,dx,ci)"""
    co, a, b, dy, dx, ci = np.meshgrid(np.arange(Cout), np.arange(2), np.arange(2), np.arange(2), np.arange(2), np.arange(Cin), indexing="ij")
    idx = oihw_index(arena.offset[name], co, ci, 2 * a + dy, 2 * b + dx, Cin, 4, 4).reshape(Cout, 4, 4 * Cin)
    return WeightLayout(packer, idx, arena.offset[name], arena.numel(name))


def layout_s2d_k3(packer, arena, name, Cout, Cin) -> WeightLayout:
    """3x3 stride-2 pad-1 conv (the generator's down-sampling convolutions with no_antialias=True, irc:468) as a 2x2 conv over the
    2x2 space-to-depth blocks of the padded frame: kernel element (r, s) = (2a + dy, 2b + dx); the elements with r = 3 or s = 3 of
    the enclosing 4x4 kernel are structural zeros"""
    idx = np.full((Cout, 4, 4 * Cin), -1, np.int64)
    base = arena.offset[name]
    co = np.arange(Cout).reshape(-1, 1); ci = np.arange(Cin).reshape(1, -1)
    for a in range(2):
        for b in range(2):
            for dy in range(2):
                for dx in range(2):
                    r, s = 2 * a + dy, 2 * b + dx
                    if r < 3 and s < 3:
                        idx[:, a * 2 + b, (dy * 2 + dx) * Cin:(dy * 2 + dx + 1) * Cin] = oihw_index(base, co, ci, r, s, Cin, 3, 3)
    return WeightLayout(packer, idx, base, arena.numel(name))


def layout_im2col(packer, arena, name, Cout, Cin, k) -> WeightLayout:
    """one tap, 64 slots: slot (r*k+s)*Cin + ci"""
    idx = np.full((Cout, 1, 64), -1, np.int64)
    co, r, s, ci = np.meshgrid(np.arange(Cout), np.arange(k), np.arange(k), np.arange(Cin), indexing="ij")
    idx[co.reshape(-1), 0, ((r * k + s) * Cin + ci).reshape(-1)] = oihw_index(arena.offset[name], co, ci, r, s, Cin, k, k).reshape(-1)
    return WeightLayout(packer, idx, arena.offset[name], arena.numel(name))


def layout_outc(packer, arena, name, Cout, Cin, k) -> WeightLayout:
    """k x k conv with tiny Cout as a GEMM over the vertical taps: output slot (s*Cout+co), tap r"""
    n = k * Cout
    idx = np.full((n, k, Cin), -1, np.int64)
    co, r, s, ci = np.meshgrid(np.arange(Cout), np.arange(k), np.arange(k), np.arange(Cin), indexing="ij")
    idx[(s * Cout + co).reshape(-1), r.reshape(-1), ci.reshape(-1)] = oihw_index(arena.offset[name], co, ci, r, s, Cin, k, k).reshape(-1)
    return WeightLayout(packer, idx, arena.offset[name], arena.numel(name), n_pad=32, k_pad_d=64)


def layout_pointwise_taps(packer, arena, name, Cin, k) -> WeightLayout:
    """k x k conv with Cout = 1 as a one-tap GEMM producing the k*k per-tap partial products"""
    n = k * k
    idx = np.full((n, 1, Cin), -1, np.int64)
    r, s, ci = np.meshgrid(np.arange(k), np.arange(k), np.arange(Cin), indexing="ij")
    idx[(r * k + s).reshape(-1), 0, ci.reshape(-1)] = oihw_index(arena.offset[name], 0, ci, r, s, Cin, k, k).reshape(-1)
    return WeightLayout(packer, idx, arena.offset[name], arena.numel(name), n_pad=32, k_pad_d=64)


def layout_convT(packer, arena, name, Cin, Cout) -> WeightLayout:
    """nn.ConvTranspose2d(Cin, Cout, 3, stride=2, padding=1, output_padding=1) (irc:495-499, :512-516) as a stride-1 GEMM over
    the INPUT frame: output pixel (2y + a, 2x + b) only sees inputs (y + dy, x + dx), dy, dx in {0, 1}, through kernel element
    (a + 1 - 2 dy, b + 1 - 2 dx) when that lies in [0, 2].  Output slot n = (a * 2 + b) * Cout + co (depth-to-space order),
    tap t = dy * 2 + dx (row shift dy * wp + dx), channel slot k = ci.  PyTorch stores the weight as (Cin, Cout, 3, 3)."""
    idx = np.full((4 * Cout, 4, Cin), -1, np.int64)
    base = arena.offset[name]
    ci = np.arange(Cin)
    for a in range(2):
        for b in range(2):
            for dy in range(2):
                for dx in range(2):
                    r, s = a + 1 - 2 * dy, b + 1 - 2 * dx
                    if not (0 <= r <= 2 and 0 <= s <= 2):
                        continue
                    for co in range(Cout):
                        idx[(a * 2 + b) * Cout + co, dy * 2 + dx, :] = base + ((ci * Cout + co) * 3 + r) * 3 + s
    return WeightLayout(packer, idx, base, arena.numel(name))


def taps_convT(wp: int) -> List[int]:
    return [dy * wp + dx for dy in range(2) for dx in range(2)]
