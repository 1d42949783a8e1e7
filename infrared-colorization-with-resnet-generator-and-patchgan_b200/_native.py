"""ctypes binding of libirc_sm100.so (include/irc_b200.h) and the `CudaBackend` that the
engine drives.

There is no fallback: if the library is missing, or the device is not sm_100, every entry
point raises.  Nothing here touches ``oracle/``."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libirc_sm100.so")
MAX_TAPS = 64

EXPORTS = [
    "irc_version", "irc_arch_check", "irc_last_error", "irc_conv_gemm", "irc_conv_stats_workspace_floats", "irc_conv_stats_finalize", "irc_tn_gemm", "irc_tn_gemm_ctas", "irc_row_index",
    "irc_in_stats", "irc_gather", "irc_in_apply_fused", "irc_in_bwd_reduce", "irc_in_bwd_apply", "irc_in_bwd_fused", "irc_in_bwd_l2", "irc_bn_finalize", "irc_bn_bwd_fix", "irc_maxpool2", "irc_maxpool2_bwd",
    "irc_colsum", "irc_im2col_rows", "irc_im2col", "irc_smallk_conv_fwd", "irc_col2im", "irc_tap_reduce", "irc_tap_expand",
    "irc_pixel_loss", "irc_ssim_fwd", "irc_ssim_bwd", "irc_hinge", "irc_feat_l1", "irc_quantize_metrics", "irc_ssim_metric",
    "irc_adam", "irc_accumulate", "irc_gather_f32", "irc_convT2d_fwd", "irc_pack_bf16", "irc_pack_std", "irc_gather_sum", "irc_gather_sum_multi", "irc_stencil_nchw", "irc_stencil_nchw_stream", "irc_fold_inplace",
    "irc_resize_area_u8", "irc_u8_to_pm1",
]


class IrcError(RuntimeError):
    pass


class ConvGemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_longlong), ("a_ld", C.c_int), ("a_chan_off", C.c_int), ("cin", C.c_int),
        ("ntaps", C.c_int), ("taps", C.c_int * MAX_TAPS),
        ("w", C.c_void_p), ("n_out", C.c_int),
        ("out", C.c_void_p), ("out_ld", C.c_longlong), ("out_chan_off", C.c_int), ("out_fp32", C.c_int),
        ("bias", C.c_void_p), ("act", C.c_int), ("slope", C.c_float),
        ("row_img", C.c_void_p),
        ("mask", C.c_void_p), ("mask_ld", C.c_longlong), ("mask_chan_off", C.c_int), ("mask_slope", C.c_float),
        ("addend", C.c_void_p), ("addend_ld", C.c_longlong), ("addend_chan_off", C.c_int),
        ("bn", C.c_int), ("mt", C.c_int), ("reuse", C.c_int), ("epilogue_direct", C.c_int),
        ("stats_part", C.c_void_p), ("stats_edge", C.c_void_p), ("rows_per_img", C.c_int),
        ("tap_out", C.c_void_p), ("tap_nshift", C.c_int), ("tap_nco", C.c_int), ("tap_H", C.c_int), ("tap_W", C.c_int), ("tap_hp", C.c_int),
        ("tap_wp", C.c_int), ("tap_oy", C.c_int), ("tap_ox", C.c_int), ("tap_act", C.c_int),
        ("tap_scale", C.c_void_p), ("tap_accumulate", C.c_int), ("k_live", C.c_int),
    ]


class TnGemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_longlong), ("a_ld", C.c_int), ("a_chan_off", C.c_int), ("m", C.c_int),
        ("b", C.c_void_p), ("b_rows", C.c_longlong), ("b_ld", C.c_int), ("b_chan_off", C.c_int), ("n", C.c_int),
        ("k_rows", C.c_longlong),
        ("ntaps", C.c_int), ("a_shift", C.c_int * MAX_TAPS), ("b_shift", C.c_int * MAX_TAPS),
        ("out", C.c_void_p),
        ("out_tap_stride", C.c_longlong), ("out_m_stride", C.c_longlong), ("out_n_stride", C.c_longlong),
        ("out_split_stride", C.c_longlong),
        ("splits", C.c_int), ("bn", C.c_int), ("tpc", C.c_int),
    ]


class CView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_longlong), ("chan_off", C.c_int), ("hp", C.c_int), ("wp", C.c_int),
                ("oy", C.c_int), ("ox", C.c_int), ("s2d_c", C.c_int)]


class GatherArgs(C.Structure):
    _fields_ = [
        ("src", CView), ("src2", CView), ("res", CView), ("dst", CView),
        ("C", C.c_int), ("n_img", C.c_int),
        ("stats", C.c_void_p), ("cnt", C.c_float), ("eps", C.c_float), ("act", C.c_int), ("slope", C.c_float),
        ("ty_idx", C.c_void_p), ("ty_w", C.c_void_p), ("ky", C.c_int),
        ("tx_idx", C.c_void_p), ("tx_w", C.c_void_p), ("kx", C.c_int),
        ("H", C.c_int), ("W", C.c_int), ("pad", C.c_int), ("halo_mode", C.c_int), ("dst_s2d", C.c_int),
        ("tile_y", C.c_int), ("tile_x", C.c_int), ("patch_y", C.c_int), ("patch_x", C.c_int),
    ]


class InBwdArgs(C.Structure):
    _fields_ = [
        ("z", CView), ("g1", CView), ("g2", CView), ("dz", CView),
        ("C", C.c_int), ("n_img", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("stats", C.c_void_p), ("cnt", C.c_float), ("eps", C.c_float), ("act", C.c_int), ("slope", C.c_float),
        ("ty_idx", C.c_void_p), ("ty_w", C.c_void_p), ("ky", C.c_int),
        ("tx_idx", C.c_void_p), ("tx_w", C.c_void_p), ("kx", C.c_int),
        ("bsum", C.c_void_p), ("work", C.c_void_p), ("work_floats", C.c_longlong),
    ]


class Im2colArgs(C.Structure):
    _fields_ = [
        ("src1", C.c_void_p), ("c1", C.c_int), ("src2", C.c_void_p), ("c2", C.c_int),
        ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("n_img", C.c_int), ("H", C.c_int), ("W", C.c_int), ("k", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("pad_mode", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int), ("row_mode", C.c_int),
        ("dst", C.c_void_p), ("row_img", C.c_void_p),
    ]


class Col2imArgs(C.Structure):
    _fields_ = [
        ("de", C.c_void_p), ("ld", C.c_longlong),
        ("C", C.c_int), ("c_first", C.c_int), ("c_out", C.c_int), ("n_img", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("k", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int), ("row_mode", C.c_int),
        ("scale", C.c_void_p), ("out", C.c_void_p), ("accumulate", C.c_int),
    ]


class SumJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("map", C.c_void_p), ("dst", C.c_void_p), ("n", C.c_longlong), ("split_stride", C.c_longlong),
                ("start", C.c_longlong), ("splits", C.c_int), ("pad_", C.c_int)]


class TapArgs(C.Structure):
    _fields_ = [("nshift", C.c_int), ("nco", C.c_int), ("dy", C.c_int * MAX_TAPS), ("dx", C.c_int * MAX_TAPS),
                ("n_img", C.c_int), ("H", C.c_int), ("W", C.c_int), ("hp", C.c_int), ("wp", C.c_int),
                ("oy", C.c_int), ("ox", C.c_int), ("live_cols_only", C.c_int)]


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises IrcError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IrcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU or PyTorch fallback for the hot path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.irc_last_error.restype = C.c_char_p
        _lib.irc_im2col_rows.restype = C.c_longlong
        _lib.irc_conv_stats_workspace_floats.restype = C.c_longlong
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise IrcError(f"libirc_sm100 error {rc}: {lib().irc_last_error().decode()}")


def arch_check() -> None:
    check(lib().irc_arch_check())


# ------------------------------------------------------------------------------------------
# engine-facing views and backend
# ------------------------------------------------------------------------------------------
@dataclass
class View:
    """Pixel (n,y,x), channel c of a bf16 [rows, ld] tensor ``t``: see irc_view in the header."""
    t: torch.Tensor
    chan_off: int
    hp: int
    wp: int
    oy: int = 0
    ox: int = 0
    s2d_c: int = 0

    def sub(self, chan_off: int) -> "View":
        return View(self.t, self.chan_off + chan_off, self.hp, self.wp, self.oy, self.ox, self.s2d_c)

    def shifted(self, oy: int, ox: int) -> "View":
        return View(self.t, self.chan_off, self.hp, self.wp, oy, ox, self.s2d_c)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _cview(v: Optional[View]) -> CView:
    c = CView()
    if v is None:
        c.ptr = None
        return c
    assert v.t.dim() == 2 and v.t.dtype == torch.bfloat16 and v.t.is_contiguous()
    c.ptr = v.t.data_ptr(); c.ld = v.t.shape[1]; c.chan_off = v.chan_off
    c.hp = v.hp; c.wp = v.wp; c.oy = v.oy; c.ox = v.ox; c.s2d_c = v.s2d_c
    return c


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Tables:
    """Separable sparse 1-D operators (idx[len,k], w[len,k]) for irc_gather / irc_in_bwd."""

    def __init__(self, ty_idx=None, ty_w=None, tx_idx=None, tx_w=None):
        self.ty_idx, self.ty_w, self.tx_idx, self.tx_w = ty_idx, ty_w, tx_idx, tx_w
        self.ky = 1 if ty_idx is None else ty_idx.shape[1]
        self.kx = 1 if tx_idx is None else tx_idx.shape[1]
        self._host = None
        self._tiling = {}
        self._stream = None

    def _load_host(self):
        if self._host is None:
            f = lambda t: None if t is None else t.cpu().numpy()
            self._host = tuple(f(t) for t in (self.ty_idx, self.ty_w, self.tx_idx, self.tx_w))
        return self._host

    def stream_window(self, kmax: int = 6) -> int:
        """Window K (rows of source kept in registers) for the streaming stencil kernel, or 0 when the tables do not
        qualify: both axes tabulated, the last source row of consecutive output rows non-decreasing, every row's
        non-zero entries within K <= 6 consecutive source rows and at most K x-entries."""
        if self._stream is None or self._stream[0] != kmax:
            import numpy as np
            k = 0
            if self.ty_idx is not None and self.tx_idx is not None:
                iy, wy, _, _ = self._load_host()
                nz = wy != 0
                if nz.any(1).all():
                    lo = np.where(nz, iy, 1 << 30).min(1); hi = np.where(nz, iy, -1).max(1)
                    span = int((hi - lo + 1).max())
                    if (np.diff(hi) >= 0).all() and max(span, self.kx, self.ky) <= kmax:
                        k = max(span, self.kx, self.ky)
            self._stream = (kmax, k)
        return self._stream[1]

    def x_bandwidth(self, L: int) -> int:
        """largest span of source columns needed by L consecutive output columns (blocks aligned at multiples of L): the staged
        row segment of the staged-rows stencil kernel"""
        key = ("xbw", L)
        if key not in self._tiling:
            import numpy as np
            _, _, ix, wx = self._load_host()
            nz = wx != 0
            lo = np.where(nz, ix, 1 << 30).min(1); hi = np.where(nz, ix, -1).max(1)
            bw = 0
            for s0 in range(0, len(lo), L):
                m = hi[s0:s0 + L] >= 0
                if m.any():
                    bw = max(bw, int(hi[s0:s0 + L][m].max() - lo[s0:s0 + L][m].min() + 1))
            self._tiling[key] = bw
        return self._tiling[key]

    def _axis_extent(self, idx, w, n_out, pad, halo_mode, T):
        """largest source index span needed by any T-wide tile of the padded output axis"""
        import numpy as np
        P = np.arange(n_out + 2 * pad) - pad
        if halo_mode == 1:
            v = np.abs(P); v = np.where(v > n_out - 1, 2 * (n_out - 1) - v, v); ok = np.ones_like(v, bool)
        else:
            v = P.copy(); ok = (v >= 0) & (v < n_out); v = np.clip(v, 0, n_out - 1)
        if idx is None:
            lo = hi = v
        else:
            big = np.where(w != 0, idx, 1 << 30).min(1); small = np.where(w != 0, idx, -1).max(1)
            lo, hi = big[v], small[v]
        ext = 1
        for s in range(0, len(P), T):
            m = ok[s:s + T]
            if m.any():
                ext = max(ext, int(hi[s:s + T][m].max() - lo[s:s + T][m].min() + 1))
        return ext

    def tiling(self, H, W, pad, halo_mode):
        """(tile_y, tile_x, patch_y, patch_x) for the shared-memory tiled gather, or zeros for trivial tables"""
        if self.ky * self.kx <= 1:
            return (0, 0, 0, 0)
        key = (H, W, pad, halo_mode)
        if key not in self._tiling:
            iy, wy, ix, wx = self._load_host()
            best = None
            for ty, tx in ((16, 16), (8, 32), (8, 16), (4, 32)):
                py = self._axis_extent(iy, wy, H, pad, halo_mode, ty); px = self._axis_extent(ix, wx, W, pad, halo_mode, tx)
                smem = py * px * 128
                if smem <= 96 * 1024:
                    cost = (py * px) / float(ty * tx)
                    if best is None or cost < best[0]:
                        best = (cost, (ty, tx, py, px))
            self._tiling[key] = best[1] if best else (0, 0, 0, 0)
        return self._tiling[key]


IDENTITY = Tables()


class CudaBackend:
    """Thin, allocation-free launcher of the C ABI on torch's current stream.  `launches`
    counts kernel launches (bench.py reports it)."""

    name = "cuda"

    def __init__(self):
        self.L = lib()
        arch_check()
        self.launches = 0
        # partials of the order-fixed two-stage reductions (InstanceNorm statistics, bias gradients)
        self.work = torch.zeros(1 << 22, device="cuda")
        # optional per-launch timing of the tensor-core kernels (bench.py's roofline pass): when `timers` is a
        # list, every GEMM launch is bracketed by CUDA events on the launching stream
        self.timers = None
        self.note = ("", "", 0.0)
        self.conv_mt = int(os.environ.get("IRC_CONV_MT", "0"))   # 0 = let the library choose the M sub-tiling of conv_gemm
        self.fused_in_bwd = os.environ.get("IRC_FUSED_IN_BWD", "1") != "0"   # cluster-resident single-pass InstanceNorm backward
        self.fused_in_apply = os.environ.get("IRC_FUSED_IN_APPLY", "1") != "0"
        self.inbwd_l2_groups = int(os.environ.get("IRC_INBWD_L2", "16"))      # image groups in flight of the L2-resident InstanceNorm backward (0 = two-pass)
        self.batch_sums = os.environ.get("IRC_BATCH_SUMS", "1") != "0"      # one launch for all split-K weight-gradient reductions
        self._pending, self._sum_tables = [], {}
        self._stats_ws = {}
        # InstanceNorm statistics in the conv epilogue for layers with at least this many reduction elements per output.  Measured
        # (profiles/): performance-neutral at K >= 1024 (the column sums cost the epilogue about what the separate pass cost),
        # a loss below it (down1, K = 576: the epilogue is on the critical path, +0.13 ms for a 0.07 ms pass).  IRC_STATS_EPI=0
        # all layers, =1000000 none
        self.stats_epilogue_min_k = int(os.environ.get("IRC_STATS_EPI", "1024"))
        self.res_epilogue_stats = os.environ.get("IRC_RES_EPI_STATS", "0") != "0"   # ResNet blocks: epilogue statistics + streaming apply instead of the cluster kernel
        self.fused_outc = os.environ.get("IRC_FUSED_OUTC", "1") != "0"      # tap reduction + bias + tanh in the GEMM epilogue of the output head
        self.gather_mode = os.environ.get("IRC_GATHER", "auto")     # lean | tiled | generic (stencil gather kernel choice)
        self.conv_epilogue_direct = int(os.environ.get("IRC_EPI_DIRECT", "0"))
        # tap-run conv_gemm (one staged A box per kernel row of taps): -1 = the library decides per layer, 0 never, 1 wherever taps form runs
        self.conv_reuse = int(os.environ.get("IRC_CONV_REUSE", "-1"))
        # inc / VGG conv1_1 / D model.0 as direct convolutions (no im2col operand round trip); 0 = im2col + one-tap GEMM
        self.direct_smallk = os.environ.get("IRC_DIRECT_SMALLK", "1") != "0"

    def time_all_launchers(self):
        """bench.py --breakdown-all: bracket EVERY launcher call with CUDA events (eager mode) so that the memory-bound
        kernels get the same per-call table as the GEMMs.  Appends (name, tag, e0, e1) to self.timers_all."""
        self.timers_all = []
        names = ["in_stats", "gather", "in_apply", "in_bwd", "fold_inplace", "maxpool2", "maxpool2_bwd", "colsum", "im2col", "col2im",
                 "tap_reduce", "tap_expand", "pixel_loss", "ssim_fwd", "ssim_bwd", "hinge", "feat_l1", "adam", "pack_bf16", "gather_sum",
                 "conv_gemm", "tn_gemm", "zero_", "smallk_conv_fwd", "pack_std"]
        for name in names:
            orig = getattr(self, name)

            def wrapped(*a, _orig=orig, _name=name, **k):
                if self.timers_all is None:
                    return _orig(*a, **k)
                tag = ",".join(str(v) for v in a if isinstance(v, int) and not isinstance(v, bool))[:40]
                if _name in ("conv_gemm", "tn_gemm", "smallk_conv_fwd"):
                    tag = self.note[0] + ":" + self.note[1]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = _orig(*a, **k); e1.record()
                self.timers_all.append((_name, tag, e0, e1))
                return r
            setattr(self, name, wrapped)

    def _timed(self, kind, fn):
        if self.timers is None:
            fn()
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        self.timers.append((kind,) + tuple(self.note) + (e0, e1))

    # ---- tensor-core GEMMs
    def conv_gemm(self, a, a_chan_off, cin, taps: Sequence[int], w, n_out, out, out_chan_off=0, bias=None, act=0, slope=0.0,
                  row_img=None, mask: Optional[View] = None, mask_slope=0.0, addend: Optional[View] = None, in_stats=None, tap=None, k_live=0):
        """in_stats = (stats [n_img, n_out, 2], n_img, rows_per_img): InstanceNorm statistics of the output from the epilogue
        (needs row_img); tap = dict(out, nshift, nco, H, W, hp, wp, oy, ox, act): horizontal tap reduction + bias + activation
        fused into the epilogue, fp32 NCHW result in tap['out'] (`out` is then only a placeholder)."""
        g = ConvGemmArgs()
        g.a = a.data_ptr(); g.a_rows = a.shape[0]; g.a_ld = a.shape[1]; g.a_chan_off = a_chan_off; g.cin = cin
        g.ntaps = len(taps)
        for i, t in enumerate(taps):
            g.taps[i] = int(t)
        assert w.dtype == torch.bfloat16 and w.shape == (n_out, len(taps) * cin), (w.shape, n_out, len(taps), cin)
        g.w = w.data_ptr(); g.n_out = n_out
        assert tap is not None or out.shape[0] == a.shape[0]
        g.out = out.data_ptr(); g.out_ld = out.shape[1]; g.out_chan_off = out_chan_off
        g.out_fp32 = int(out.dtype == torch.float32)
        g.bias = None if bias is None else bias.data_ptr()
        g.act = act; g.slope = slope
        g.row_img = None if row_img is None else row_img.data_ptr()
        if mask is not None:
            assert mask.t.shape[0] == a.shape[0]
            g.mask = mask.t.data_ptr(); g.mask_ld = mask.t.shape[1]; g.mask_chan_off = mask.chan_off; g.mask_slope = mask_slope
        if addend is not None:
            assert addend.t.shape[0] == a.shape[0]
            g.addend = addend.t.data_ptr(); g.addend_ld = addend.t.shape[1]; g.addend_chan_off = addend.chan_off
        g.bn = 0; g.mt = self.conv_mt; g.reuse = self.conv_reuse; g.epilogue_direct = self.conv_epilogue_direct
        g.k_live = int(k_live)            # cin == 64: operand columns >= k_live are structural zeros
        fin = None
        if in_stats is not None:
            stats, n_img, rows_per_img = in_stats
            assert row_img is not None and stats.shape == (n_img, n_out, 2) and stats.dtype == torch.float32
            part, edge = self._stats_workspace(a.shape[0], n_out, n_img)
            g.stats_part = part.data_ptr(); g.stats_edge = edge.data_ptr(); g.rows_per_img = rows_per_img
            fin = (part, edge, n_img, rows_per_img, n_out, stats)
        if tap is not None:
            g.tap_out = tap["out"].data_ptr(); g.tap_nshift = tap["nshift"]; g.tap_nco = tap["nco"]; g.tap_H = tap["H"]; g.tap_W = tap["W"]
            g.tap_hp = tap["hp"]; g.tap_wp = tap["wp"]; g.tap_oy = tap["oy"]; g.tap_ox = tap["ox"]; g.tap_act = tap["act"]
            g.tap_scale = _p(tap.get("scale")); g.tap_accumulate = int(bool(tap.get("accumulate", False)))
        self._timed("conv_gemm", lambda: check(self.L.irc_conv_gemm(C.byref(g), _stream()))); self.launches += 1
        if fin is not None:
            part, edge, n_img, rows_per_img, n_out_, stats = fin
            check(self.L.irc_conv_stats_finalize(_p(part), _p(edge), n_img, rows_per_img, n_out_, _p(stats), _stream())); self.launches += 1

    def _stats_workspace(self, rows, n_out, n_img):
        """partials of the epilogue statistics: one buffer per (rows, n_out) shape, allocated on first use (before any capture)"""
        key = (rows, n_out, n_img)
        ws = self._stats_ws.get(key)
        if ws is None:
            nf = int(self.L.irc_conv_stats_workspace_floats(C.c_longlong(rows), n_out))
            ws = self._stats_ws[key] = (torch.zeros(nf, device="cuda"), torch.zeros((n_img + 1) * n_out * 2, device="cuda"))
        return ws

    def tn_gemm(self, a, a_chan_off, m, b, b_chan_off, n, k_rows, a_shift, b_shift, out, tap_stride, m_stride, n_stride,
                splits, split_stride):
        g = TnGemmArgs()
        g.a = a.data_ptr(); g.a_rows = a.shape[0]; g.a_ld = a.shape[1]; g.a_chan_off = a_chan_off; g.m = m
        g.b = b.data_ptr(); g.b_rows = b.shape[0]; g.b_ld = b.shape[1]; g.b_chan_off = b_chan_off; g.n = n
        g.k_rows = k_rows; g.ntaps = len(a_shift)
        for i in range(len(a_shift)):
            g.a_shift[i] = int(a_shift[i]); g.b_shift[i] = int(b_shift[i])
        assert out.dtype == torch.float32
        g.out = out.data_ptr(); g.out_tap_stride = tap_stride; g.out_m_stride = m_stride; g.out_n_stride = n_stride
        g.out_split_stride = split_stride; g.splits = splits; g.bn = 0; g.tpc = 0
        self._timed("tn_gemm", lambda: check(self.L.irc_tn_gemm(C.byref(g), _stream()))); self.launches += 1

    def tn_gemm_ctas(self, m, n, ntaps):
        return int(self.L.irc_tn_gemm_ctas(m, n, ntaps, 1))

    # ---- frames
    def row_index(self, row_img, n_img, hp, wp, y0, y1, x0, x1):
        check(self.L.irc_row_index(_p(row_img), n_img, hp, wp, y0, y1, x0, x1, _stream())); self.launches += 1

    def in_stats(self, z: View, C_, n_img, H, W, stats):
        cv = _cview(z)
        check(self.L.irc_in_stats(C.byref(cv), C_, n_img, H, W, _p(stats), _p(self.work), C.c_longlong(self.work.numel()), _stream()))
        self.launches += 1

    def gather(self, src: View, dst: View, C_, n_img, H, W, pad, halo_mode, tables: Tables = IDENTITY, src2=None, res=None,
               stats=None, cnt=0.0, eps=1e-5, act=0, slope=0.0, dst_s2d=0):
        g = GatherArgs()
        g.src = _cview(src); g.src2 = _cview(src2); g.res = _cview(res); g.dst = _cview(dst)
        g.C = C_; g.n_img = n_img
        g.stats = None if stats is None else stats.data_ptr(); g.cnt = cnt; g.eps = eps; g.act = act; g.slope = slope
        g.ty_idx = None if tables.ty_idx is None else tables.ty_idx.data_ptr()
        g.ty_w = None if tables.ty_w is None else tables.ty_w.data_ptr(); g.ky = tables.ky
        g.tx_idx = None if tables.tx_idx is None else tables.tx_idx.data_ptr()
        g.tx_w = None if tables.tx_w is None else tables.tx_w.data_ptr(); g.kx = tables.kx
        g.H = H; g.W = W; g.pad = pad; g.halo_mode = halo_mode; g.dst_s2d = dst_s2d
        mode = self.gather_mode
        if mode in ("auto", "stream"):
            k = tables.stream_window()
            ok = (k > 0 and not dst_s2d and not src.s2d_c and res is None and (stats is not None or not act) and
                  (src2 is None or (stats is None and not src2.s2d_c)) and (pad == 0 or (H > 2 * pad + 1 and W > 2 * pad + 1)))
            if ok:
                # rows per strip: enough blocks for >= ~4 waves of 2 x 148 resident blocks, at most 32 rows
                lanes = max(1, 256 // (C_ // 8))
                blocks_per_row = n_img * ((W + lanes - 1) // lanes)
                strip = max(4, min(32, (blocks_per_row * H) // 1184))
                # patch_x = source-column span of a block's L output columns (staged-rows kernel; 0 selects the register-only one)
                g.tile_y, g.tile_x, g.patch_y, g.patch_x = -3, k, strip, tables.x_bandwidth(min(lanes, W))
                check(self.L.irc_gather(C.byref(g), _stream())); self.launches += 1
                return
            # measured (scripts/bench_elem.py): the shared-memory tiled kernel wins for the up-sampling stencils and the
            # transposed ones with many taps / two sources; the register-table kernel wins for stride-2 down-sampling
            mode = "tiled" if (tables.ky >= 6 or src2 is not None or H > src.hp) else "lean"
        g.tile_y, g.tile_x, g.patch_y, g.patch_x = (tables.tiling(H, W, pad, halo_mode) if (C_ % 32 == 0 and mode == "tiled") else
                                                    ((-2, 0, 0, 0) if mode == "generic" else (0, 0, 0, 0)))
        check(self.L.irc_gather(C.byref(g), _stream())); self.launches += 1

    def in_apply(self, z: View, dst: View, C_, n_img, H, W, pad, halo_mode, stats, eps=1e-5, act=0, slope=0.0, res=None, dst_s2d=0):
        """InstanceNorm of the map z (statistics written to `stats` for the backward pass) + activation (+ residual) into the
        frame dst, ring included: one cluster launch for small maps, irc_in_stats + irc_gather otherwise."""
        if self.fused_in_apply and not dst_s2d and not z.s2d_c and C_ % 32 == 0 and H * W <= 4096 and (pad == 0 or min(H, W) > 2 * pad + 1):
            g = GatherArgs()
            g.src = _cview(z); g.src2 = _cview(None); g.res = _cview(res); g.dst = _cview(dst)
            g.C = C_; g.n_img = n_img; g.cnt = float(H * W); g.eps = eps; g.act = act; g.slope = slope
            g.H = H; g.W = W; g.pad = pad; g.halo_mode = halo_mode
            check(self.L.irc_in_apply_fused(C.byref(g), _p(stats), _stream())); self.launches += 1
            return
        self.in_stats(z, C_, n_img, H, W, stats)
        self.gather(z, dst, C_, n_img, H, W, pad, halo_mode, stats=stats, cnt=H * W, eps=eps, act=act, slope=slope, res=res, dst_s2d=dst_s2d)

    def _bwd_args(self, z, g1, g2, dz, C_, n_img, H, W, stats, cnt, eps, act, slope, tables, bsum):
        g = InBwdArgs()
        g.z = _cview(z); g.g1 = _cview(g1); g.g2 = _cview(g2); g.dz = _cview(dz)
        g.C = C_; g.n_img = n_img; g.H = H; g.W = W
        g.stats = None if stats is None else stats.data_ptr(); g.cnt = cnt; g.eps = eps; g.act = act; g.slope = slope
        g.ty_idx = None if tables.ty_idx is None else tables.ty_idx.data_ptr()
        g.ty_w = None if tables.ty_w is None else tables.ty_w.data_ptr(); g.ky = tables.ky
        g.tx_idx = None if tables.tx_idx is None else tables.tx_idx.data_ptr()
        g.tx_w = None if tables.tx_w is None else tables.tx_w.data_ptr(); g.kx = tables.kx
        g.bsum = None if bsum is None else bsum.data_ptr()
        g.work = self.work.data_ptr(); g.work_floats = self.work.numel()
        return g

    def bn_finalize(self, stats, n_img, group, C_, cnt_per_img, gamma, beta, running_mean, running_var, eff, momentum=0.1, eps=1e-5, training=True,
                    updates=1):
        """nn.BatchNorm2d on the InstanceNorm kernels: per-image sums -> effective moments eff [n_img, C, 2] (see irc_bn_finalize)"""
        assert eff.shape == (n_img, C_, 2) and eff.dtype == torch.float32 and gamma.numel() == C_ and running_mean.numel() == C_
        check(self.L.irc_bn_finalize(_p(stats), n_img, group, C_, C.c_float(cnt_per_img), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
                                     C.c_float(momentum), C.c_float(eps), int(bool(training)), int(updates), _p(eff), _stream()))
        self.launches += 1

    def in_bwd(self, z: View, g1: View, dz: View, C_, n_img, H, W, stats=None, cnt=0.0, eps=1e-5, act=0, slope=0.0,
               tables: Tables = IDENTITY, g2=None, bsum=None, fold_pad=0, bn=None):
        """reduce (when normalised) + apply.  fold_pad > 0: g1 is the view of a frame holding the gradient w.r.t. the
        reflection-padded map (ring included); the fold happens inside (fused kernel) or as a separate in-place pass.
        bn = dict(group, gamma, beta, dgamma, dbeta, accumulate): BatchNorm backward - `stats` holds the effective moments of
        bn_finalize (eps < 0), the two passes run separately with irc_bn_bwd_fix between them."""
        g = self._bwd_args(z, g1, g2, dz, C_, n_img, H, W, stats, cnt, eps, act, slope, tables, bsum)
        if bn is not None:
            assert stats is not None and eps < 0 and bsum is not None
            if fold_pad:
                self.fold_inplace(g1.t, g1.chan_off, C_, n_img, H, W, fold_pad)
            check(self.L.irc_in_bwd_reduce(C.byref(g), _stream()))
            check(self.L.irc_bn_bwd_fix(_p(bsum), n_img, int(bn["group"]), C_, _p(bn["gamma"]), _p(bn["beta"]), _p(bn["dgamma"]), _p(bn["dbeta"]),
                                        int(bool(bn.get("accumulate", False))), _stream()))
            check(self.L.irc_in_bwd_apply(C.byref(g), _stream()))
            self.launches += 3
            return
        fused = (self.fused_in_bwd and stats is not None and g2 is None and tables.ty_idx is None and tables.tx_idx is None
                 and C_ % 32 == 0 and H * W <= 4096 and not (z.s2d_c or g1.s2d_c or dz.s2d_c))
        if fused:
            check(self.L.irc_in_bwd_fused(C.byref(g), fold_pad, _stream())); self.launches += 1
            return
        if fold_pad:
            self.fold_inplace(g1.t, g1.chan_off, C_, n_img, H, W, fold_pad)
        if (self.inbwd_l2_groups > 0 and stats is not None and tables.ty_idx is None and tables.tx_idx is None and C_ in (64, 128, 256)
                and W >= 2048 // C_ and not (z.s2d_c or g1.s2d_c or dz.s2d_c or (g2 is not None and g2.s2d_c))):
            # large maps: one launch, every image's second read served by the L2 (3 tensor units over HBM instead of 5)
            check(self.L.irc_in_bwd_l2(C.byref(g), self.inbwd_l2_groups, _stream())); self.launches += 1
            return
        # (running the two passes over L2-sized groups of images was measured: 14.50 ms/step unchunked vs 14.67 / 15.11 /
        # 15.67 ms with 64 / 40 / 24 MB groups - the extra launches cost more than the L2 hits save)
        if stats is not None:
            check(self.L.irc_in_bwd_reduce(C.byref(g), _stream())); self.launches += 1
        check(self.L.irc_in_bwd_apply(C.byref(g), _stream())); self.launches += 1

    def fold_inplace(self, fr_t, chan_off, C_, n_img, H, W, p):
        """fr_t: the [rows, ld] tensor of a frame with pad p"""
        check(self.L.irc_fold_inplace(_p(fr_t), C.c_longlong(fr_t.shape[1]), chan_off, C_, n_img, H, W, p, _stream())); self.launches += 1

    def maxpool2(self, src: View, dst: View, C_, n_img, Ho, Wo):
        a, b = _cview(src), _cview(dst)
        check(self.L.irc_maxpool2(C.byref(a), C.byref(b), C_, n_img, Ho, Wo, _stream())); self.launches += 1

    def maxpool2_bwd(self, src: View, g: View, dsrc: View, C_, n_img, Ho, Wo):
        a, b, c = _cview(src), _cview(g), _cview(dsrc)
        check(self.L.irc_maxpool2_bwd(C.byref(a), C.byref(b), C.byref(c), C_, n_img, Ho, Wo, _stream())); self.launches += 1

    def colsum(self, a, chan_off, C_, out, row_img=None):
        check(self.L.irc_colsum(_p(a), C.c_longlong(a.shape[0]), C.c_longlong(a.shape[1]), chan_off, C_, _p(row_img), _p(out),
                                _p(self.work), C.c_longlong(self.work.numel()), _stream()))
        self.launches += 1

    # ---- degenerate convolutions
    def im2col_rows(self, row_mode, n_img, Ho, Wo):
        return int(self.L.irc_im2col_rows(row_mode, n_img, Ho, Wo))

    def im2col(self, src1, src2, scale, shift, n_img, H, W, k, stride, pad, pad_mode, Ho, Wo, row_mode, dst, row_img=None):
        g = Im2colArgs()
        g.src1 = src1.data_ptr(); g.c1 = src1.shape[1]
        g.src2 = None if src2 is None else src2.data_ptr(); g.c2 = 0 if src2 is None else src2.shape[1]
        g.scale = None if scale is None else scale.data_ptr(); g.shift = None if shift is None else shift.data_ptr()
        g.n_img = n_img; g.H = H; g.W = W; g.k = k; g.stride = stride; g.pad = pad; g.pad_mode = pad_mode
        g.Ho = Ho; g.Wo = Wo; g.row_mode = row_mode
        assert dst.shape[1] == 64 and dst.shape[0] == self.im2col_rows(row_mode, n_img, Ho, Wo)
        g.dst = dst.data_ptr(); g.row_img = None if row_img is None else row_img.data_ptr()
        check(self.L.irc_im2col(C.byref(g), _stream())); self.launches += 1

    def smallk_conv_fwd(self, src1, src2, scale, shift, n_img, H, W, k, stride, pad, pad_mode, Ho, Wo, row_mode, w, out, bias=None, act=0,
                        slope=0.0, E=None, row_img=None):
        """direct small-K convolution (inc, VGG conv1_1, D model.0): im2col geometry of `im2col`, weights w [64, 64] bf16, bf16 rows
        out [rows, 64]; E (optional) also receives the im2col operand for the weight gradient"""
        g = Im2colArgs()
        g.src1 = src1.data_ptr(); g.c1 = src1.shape[1]
        g.src2 = None if src2 is None else src2.data_ptr(); g.c2 = 0 if src2 is None else src2.shape[1]
        g.scale = None if scale is None else scale.data_ptr(); g.shift = None if shift is None else shift.data_ptr()
        g.n_img = n_img; g.H = H; g.W = W; g.k = k; g.stride = stride; g.pad = pad; g.pad_mode = pad_mode
        g.Ho = Ho; g.Wo = Wo; g.row_mode = row_mode
        rows = self.im2col_rows(row_mode, n_img, Ho, Wo)
        assert out.shape == (rows, 64) and out.is_contiguous() and tuple(w.shape) == (64, 64) and w.is_contiguous()
        assert E is None or (E.shape == (rows, 64) and E.is_contiguous())
        g.dst = None if E is None else E.data_ptr(); g.row_img = None if row_img is None else row_img.data_ptr()
        self._timed("smallk_conv", lambda: check(self.L.irc_smallk_conv_fwd(C.byref(g), _p(w), _p(bias), int(act), C.c_float(slope), _p(out), _stream())))
        self.launches += 1

    def col2im(self, de, C_, c_first, c_out, n_img, H, W, k, stride, pad, Ho, Wo, row_mode, scale, out, accumulate):
        g = Col2imArgs()
        g.de = de.data_ptr(); g.ld = de.shape[1]; g.C = C_; g.c_first = c_first; g.c_out = c_out; g.n_img = n_img
        g.H = H; g.W = W; g.k = k; g.stride = stride; g.pad = pad; g.Ho = Ho; g.Wo = Wo; g.row_mode = row_mode
        g.scale = None if scale is None else scale.data_ptr(); g.out = out.data_ptr(); g.accumulate = int(accumulate)
        check(self.L.irc_col2im(C.byref(g), _stream())); self.launches += 1

    @staticmethod
    def _tap(shifts, nco, n_img, H, W, hp, wp, oy, ox):
        t = TapArgs(); t.nshift = len(shifts); t.nco = nco
        for i, (dy, dx) in enumerate(shifts):
            t.dy[i] = int(dy); t.dx[i] = int(dx)
        t.n_img = n_img; t.H = H; t.W = W; t.hp = hp; t.wp = wp; t.oy = oy; t.ox = ox
        return t

    def tap_reduce(self, P, shifts, nco, n_img, H, W, hp, wp, oy, ox, bias, act, out):
        t = self._tap(shifts, nco, n_img, H, W, hp, wp, oy, ox)
        check(self.L.irc_tap_reduce(C.byref(t), _p(P), C.c_longlong(P.shape[1]), _p(bias), act, _p(out), _stream())); self.launches += 1

    def tap_expand(self, g, y, shifts, nco, n_img, H, W, hp, wp, oy, ox, E, dbias=None, live_cols_only=False):
        """live_cols_only: E's columns beyond the tap columns (rounded up to 8) are never written - the caller's buffer holds zeros there"""
        t = self._tap(shifts, nco, n_img, H, W, hp, wp, oy, ox)
        t.live_cols_only = int(bool(live_cols_only))
        assert E.shape == (n_img * hp * wp, 64)
        check(self.L.irc_tap_expand(C.byref(t), _p(g), _p(y), _p(E), _p(dbias), _p(self.work), C.c_longlong(self.work.numel()), _stream()))
        self.launches += 1 if dbias is None else 3

    # ---- losses
    def pixel_loss(self, fake, target, w_l1, w_tvv, w_tvh, sums, dfake):
        n, c, h, w = fake.shape
        check(self.L.irc_pixel_loss(_p(fake), _p(target), n, c, h, w, C.c_float(w_l1), C.c_float(w_tvv), C.c_float(w_tvh),
                                    _p(sums), _p(dfake), _stream())); self.launches += 1

    @staticmethod
    def _win(window):
        """the 11 window taps are a HOST array in the C ABI (they travel as a kernel argument)"""
        assert window.device.type == "cpu" and window.numel() == 11
        return (C.c_float * 11)(*[float(v) for v in window.tolist()])

    def ssim_fwd(self, img1, img2, scale, shift, window, sums, ga=None, gb=None, gc=None):
        n, c, h, w = img1.shape
        check(self.L.irc_ssim_fwd(_p(img1), _p(img2), n, c, h, w, C.c_float(scale), C.c_float(shift), self._win(window), _p(sums),
                                  _p(ga), _p(gb), _p(gc), _stream())); self.launches += 1

    def ssim_bwd(self, img1, img2, scale, shift, window, ga, gb, gc, coef, dimg1, accumulate):
        n, c, h, w = img1.shape
        check(self.L.irc_ssim_bwd(_p(img1), _p(img2), n, c, h, w, C.c_float(scale), C.c_float(shift), self._win(window), _p(ga), _p(gb),
                                  _p(gc), C.c_float(coef), _p(dimg1), int(accumulate), _stream())); self.launches += 1

    def hinge(self, pred, n_real, mode, w_real, w_fake, sums, dpred):
        check(self.L.irc_hinge(_p(pred), C.c_longlong(pred.numel()), C.c_longlong(n_real), mode, C.c_float(w_real),
                               C.c_float(w_fake), _p(sums), _p(dpred), _stream())); self.launches += 1

    def feat_l1(self, feat, rows_half, C_, w, sums, dz):
        check(self.L.irc_feat_l1(_p(feat), C.c_longlong(rows_half), C.c_longlong(feat.shape[1]), C_, C.c_float(w), _p(sums),
                                 _p(dz), C.c_longlong(0 if dz is None else dz.shape[1]), _stream())); self.launches += 1

    def quantize_metrics(self, fake, gt, u8, sums):
        n, c, h, w = fake.shape
        check(self.L.irc_quantize_metrics(_p(fake), _p(gt), n, c, h, w, _p(u8), _p(sums), _stream())); self.launches += 1

    def ssim_metric(self, u8, gt, sums):
        """skimage-style SSIM metric (irc:1208-1215): u8 [n, H, W, 3] uint8 predictions, gt fp32 [n, 3, H, W] in [0, 1], sums fp64 [n]"""
        n, h, w, c = u8.shape
        assert c == 3 and u8.dtype == torch.uint8 and gt.dtype == torch.float32 and tuple(gt.shape) == (n, 3, h, w) and sums.dtype == torch.float64
        check(self.L.irc_ssim_metric(_p(u8), _p(gt), n, h, w, _p(sums), _stream())); self.launches += 1

    # ---- optimizer / layout
    def adam(self, p, g, m, v, hyper, step_dev):
        """hyper: fp64 device vector {lr, b1, b2, eps, lr_scale, grad_scale}; step_dev: int64 device step counter (advanced here)"""
        assert hyper.dtype == torch.float64 and step_dev.dtype == torch.int64
        check(self.L.irc_adam(_p(p), _p(g), _p(m), _p(v), C.c_longlong(p.numel()), _p(hyper), _p(step_dev), _stream())); self.launches += 2

    def accumulate(self, sums, coef, acc):
        assert sums.dtype == torch.float32 and coef.dtype == torch.float32 and acc.dtype == torch.float64 and coef.shape[1] == sums.numel() + 1
        check(self.L.irc_accumulate(_p(sums), sums.numel(), _p(coef), coef.shape[0], _p(acc), _stream())); self.launches += 1

    def gather_f32(self, src, map_, dst):
        assert src.dtype == torch.float32 and dst.dtype == torch.float32 and map_.dtype == torch.int32 and map_.numel() == dst.numel()
        check(self.L.irc_gather_f32(_p(src), _p(map_), C.c_longlong(map_.numel()), _p(dst), _stream())); self.launches += 1

    def pack_bf16(self, src, map_, dst):
        check(self.L.irc_pack_bf16(_p(src), _p(map_), C.c_longlong(map_.numel()), _p(dst), _stream())); self.launches += 1

    def pack_std(self, arena_flat, packed, jobs, njobs, total_blocks, max_taps):
        """structured packing of the standard convolution layers (layout.Packer): jobs = uint8 device tensor of PackStdJob records"""
        check(self.L.irc_pack_std(_p(arena_flat), _p(packed), _p(jobs), njobs, total_blocks, max_taps, _stream())); self.launches += 1

    def gather_sum(self, src, map_, splits, split_stride, dst):
        check(self.L.irc_gather_sum(_p(src), _p(map_), C.c_longlong(map_.numel()), splits, C.c_longlong(split_stride), _p(dst),
                                    _stream())); self.launches += 1

    def gather_sum_deferred(self, src, map_, splits, split_stride, dst):
        """queue a gather_sum; flush_sums() runs everything queued since the last flush in ONE launch"""
        if not self.batch_sums:
            return self.gather_sum(src, map_, splits, split_stride, dst)
        self._pending.append((src, map_, int(splits), int(split_stride), dst))

    def flush_sums(self):
        if not self._pending:
            return
        key = tuple((a.data_ptr(), m.data_ptr(), s, st, d.data_ptr(), m.numel()) for a, m, s, st, d in self._pending)
        ent = self._sum_tables.get(key)
        if ent is None:
            arr = (SumJob * len(key))()
            start = 0
            for j, (a, m, s, st, d, n) in zip(arr, key):
                j.src, j.map, j.dst, j.n, j.split_stride, j.start, j.splits = a, m, d, n, st, start, s
                start += n
            dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
            ent = self._sum_tables[key] = (dev, len(key), start, list(self._pending))      # keeps the buffers alive
        dev, nj, total, _ = ent
        self._pending = []
        check(self.L.irc_gather_sum_multi(_p(dev), nj, C.c_longlong(total), _stream())); self.launches += 1

    def stencil_nchw(self, x, out, tables: Tables, accumulate=False):
        n, c, hi, wi = x.shape
        ho, wo = out.shape[2], out.shape[3]
        assert x.dtype == torch.float32 and out.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous()
        k = tables.stream_window(8) if self.gather_mode in ("auto", "stream") else 0
        if k:
            check(self.L.irc_stencil_nchw_stream(_p(x), _p(out), n * c, hi, wi, ho, wo, _p(tables.ty_idx), _p(tables.ty_w), tables.ky,
                                                 _p(tables.tx_idx), _p(tables.tx_w), tables.kx, k, int(accumulate), _stream()))
            self.launches += 1
            return
        check(self.L.irc_stencil_nchw(_p(x), _p(out), n * c, hi, wi, ho, wo, _p(tables.ty_idx), _p(tables.ty_w), tables.ky,
                                      _p(tables.tx_idx), _p(tables.tx_w), tables.kx, int(accumulate), _stream())); self.launches += 1

    # ---- input pipeline
    def resize_area_u8(self, src, dst, tables, mode, img_max=None):
        """src uint8 [n, Hs, Ws, C] -> dst uint8 [n, Hd, Wd, C], cv2.INTER_AREA; tables = (xi, xw, yi, yw) device tensors (mode 0)"""
        n, hs, ws, c = src.shape
        hd, wd = dst.shape[1], dst.shape[2]
        assert src.dtype == torch.uint8 and dst.dtype == torch.uint8 and src.is_contiguous() and dst.is_contiguous() and dst.shape[3] == c
        xi, xw, yi, yw = tables if tables is not None else (None, None, None, None)
        kx = 0 if xi is None else xi.shape[1]; ky = 0 if yi is None else yi.shape[1]
        check(self.L.irc_resize_area_u8(_p(src), n, hs, ws, c, hd, wd, _p(xi), _p(xw), kx, _p(yi), _p(yw), ky, mode, _p(dst), _p(img_max), _stream()))
        self.launches += 1

    def u8_to_pm1(self, src, out, swap_rb=False, flip=None, img_max=None):
        n, h, w, c = src.shape
        assert src.dtype == torch.uint8 and out.dtype == torch.float32 and out.shape == (n, c, h, w) and src.is_contiguous() and out.is_contiguous()
        check(self.L.irc_u8_to_pm1(_p(src), n, h, w, c, int(swap_rb), _p(flip), _p(img_max), _p(out), _stream())); self.launches += 1

    def zero_(self, t):
        t.zero_(); self.launches += 1
