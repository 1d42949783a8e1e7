"""Plain-torch statement of every primitive of include/irc_b200.h.  TEST INFRASTRUCTURE.

Two uses: (1) `-m gpu` tests run each CUDA primitive and this restatement on the same device
buffers and compare; (2) `-m "not gpu"` tests drive the product's engine (host logic: buffer
geometry, tap tables, gradient routing) with this backend on CPU tensors and compare the
result with the oracle.  The product never imports this file."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _act(v, act, slope):
    if act == 1:
        return torch.relu(v)
    if act == 2:
        return torch.where(v > 0, v, v * slope)
    return v


def _dact(v, act, slope):
    if act == 1:
        return (v > 0).to(v.dtype)
    if act == 2:
        return torch.where(v > 0, torch.ones_like(v), torch.full_like(v, slope))
    return torch.ones_like(v)


def _reflect(i, n):
    i = i.abs()
    return torch.where(i > n - 1, 2 * (n - 1) - i, i)


class RefBackend:
    name = "ref"

    def __init__(self):
        self.launches = 0
        self.note = None

    # ------------------------------------------------------------------ addressing
    @staticmethod
    def _addr(v, n_img, ys, xs, C):
        """row [n,H,W] and first-channel [H,W] index tensors for pixels (ys x xs) of a view."""
        dev = v.t.device
        n = torch.arange(n_img, device=dev).view(-1, 1, 1)
        Y = (ys + v.oy).view(1, -1, 1); X = (xs + v.ox).view(1, 1, -1)
        if v.s2d_c:
            row = (n * v.hp + (Y >> 1)) * v.wp + (X >> 1)
            ch = v.chan_off + ((Y & 1) * 2 + (X & 1)) * v.s2d_c
        else:
            row = (n * v.hp + Y) * v.wp + X
            ch = torch.full_like(Y * X, v.chan_off)
        return row, ch

    def _read(self, v, n_img, ys, xs, C):
        row, ch = self._addr(v, n_img, ys, xs, C)
        c = torch.arange(C, device=v.t.device)
        return v.t[row.unsqueeze(-1), (ch.unsqueeze(-1) + c)].float()

    def _write(self, v, n_img, ys, xs, C, val):
        row, ch = self._addr(v, n_img, ys, xs, C)
        c = torch.arange(C, device=v.t.device)
        v.t[row.unsqueeze(-1), (ch.unsqueeze(-1) + c)] = val.to(v.t.dtype)

    @staticmethod
    def _moments(stats, cnt, eps):
        if eps < 0:          # effective (mean, 1/std) pairs of bn_finalize
            return stats[..., 0][:, None, None, :], stats[..., 1][:, None, None, :]
        s = stats[..., 0] / cnt
        var = (stats[..., 1] / cnt - s * s).clamp_min(0)
        return s[:, None, None, :], torch.rsqrt(var + eps)[:, None, None, :]

    def _table_gather(self, srcs, n_img, H, W, C, tables, pre=None):
        dev = srcs[0].t.device
        ys = torch.arange(H, device=dev); xs = torch.arange(W, device=dev)
        acc = torch.zeros(n_img, H, W, C, device=dev)
        for i in range(tables.ky):
            iy = ys if tables.ty_idx is None else tables.ty_idx[:, i].long()
            wy = torch.ones(H, device=dev) if tables.ty_w is None else tables.ty_w[:, i]
            for j in range(tables.kx):
                ix = xs if tables.tx_idx is None else tables.tx_idx[:, j].long()
                wx = torch.ones(W, device=dev) if tables.tx_w is None else tables.tx_w[:, j]
                v = self._read(srcs[0], n_img, iy, ix, C)
                if pre is not None:
                    v = pre(v)
                for s in srcs[1:]:
                    v = v + self._read(s, n_img, iy, ix, C)
                acc = acc + (wy.view(1, -1, 1, 1) * wx.view(1, 1, -1, 1)) * v
        return acc

    # ------------------------------------------------------------------ GEMMs
    stats_epilogue_min_k = 1024
    fused_outc = True

    def conv_gemm(self, a, a_chan_off, cin, taps, w, n_out, out, out_chan_off=0, bias=None, act=0, slope=0.0,
                  row_img=None, mask=None, mask_slope=0.0, addend=None, in_stats=None, tap=None, k_live=0):
        self.launches += 1
        if tap is not None:
            # horizontal tap reduction + bias + activation fused into the epilogue: the bias belongs to the reduced output
            P = torch.zeros(a.shape[0], n_out, device=a.device)
            self.conv_gemm(a, a_chan_off, cin, taps, w, n_out, P)
            h = (tap["nshift"] - 1) // 2
            n_img = a.shape[0] // (tap["hp"] * tap["wp"])
            res = torch.zeros_like(tap["out"]) if (tap.get("accumulate") or tap.get("scale") is not None) else tap["out"]
            self.tap_reduce(P, [(0, j - h) for j in range(tap["nshift"])], tap["nco"], n_img, tap["H"], tap["W"], tap["hp"], tap["wp"],
                            tap["oy"], tap["ox"], bias, tap["act"], res)
            if res is not tap["out"]:
                if tap.get("scale") is not None:
                    res = res * tap["scale"].view(1, -1, 1, 1)
                if tap.get("accumulate"):
                    tap["out"] += res
                else:
                    tap["out"].copy_(res)
            return
        rows = a.shape[0]
        A = a[:, a_chan_off:a_chan_off + cin].float()
        acc = torch.zeros(rows, n_out, device=a.device)
        q = torch.arange(rows, device=a.device)
        for t, sh in enumerate(taps):
            idx = q + int(sh)
            ok = ((idx >= 0) & (idx < rows)).float().unsqueeze(1)
            acc += (A[idx.clamp(0, rows - 1)] * ok) @ w[:, t * cin:(t + 1) * cin].float().t()
        if bias is not None:
            acc = acc + bias
        if mask is not None:
            m = mask.t[:, mask.chan_off:mask.chan_off + n_out].float()
            acc = acc * torch.where(m > 0, torch.ones_like(m), torch.full_like(m, mask_slope))
        if addend is not None:
            acc = acc + addend.t[:, addend.chan_off:addend.chan_off + n_out].float()
        acc = _act(acc, act, slope)
        if row_img is not None:
            acc = acc * (row_img >= 0).float().unsqueeze(1)
        out[:, out_chan_off:out_chan_off + n_out] = acc.to(out.dtype)
        if in_stats is not None:
            stats, n_img, rows_per_img = in_stats
            assert row_img is not None
            st = out[:, out_chan_off:out_chan_off + n_out].float().view(n_img, rows_per_img, n_out)      # the stored (rounded) values
            stats.copy_(torch.stack([st.sum(1), (st * st).sum(1)], -1))

    def tn_gemm(self, a, a_chan_off, m, b, b_chan_off, n, k_rows, a_shift, b_shift, out, tap_stride, m_stride, n_stride,
                splits, split_stride):
        self.launches += 1
        A = a[:, a_chan_off:a_chan_off + m].float(); B = b[:, b_chan_off:b_chan_off + n].float()
        q = torch.arange(k_rows, device=a.device)
        flat = out.view(-1)
        mi = torch.arange(m, device=a.device).view(-1, 1); ni = torch.arange(n, device=a.device).view(1, -1)
        for t in range(len(a_shift)):
            ia = q + int(a_shift[t]); ib = q + int(b_shift[t])
            oka = ((ia >= 0) & (ia < a.shape[0])).float().unsqueeze(1); okb = ((ib >= 0) & (ib < b.shape[0])).float().unsqueeze(1)
            r = (A[ia.clamp(0, a.shape[0] - 1)] * oka).t() @ (B[ib.clamp(0, b.shape[0] - 1)] * okb)
            for s in range(splits):
                idx = s * split_stride + t * tap_stride + mi * m_stride + ni * n_stride
                flat[idx.reshape(-1)] = (r if s == 0 else torch.zeros_like(r)).reshape(-1)

    def tn_gemm_ctas(self, m, n, ntaps):
        return ((m + 127) // 128) * ((n + 255) // 256) * ntaps

    # ------------------------------------------------------------------ frames
    def row_index(self, row_img, n_img, hp, wp, y0, y1, x0, x1):
        self.launches += 1
        dev = row_img.device
        y = torch.arange(hp, device=dev).view(1, -1, 1); x = torch.arange(wp, device=dev).view(1, 1, -1)
        n = torch.arange(n_img, device=dev).view(-1, 1, 1)
        live = (y >= y0) & (y < y1) & (x >= x0) & (x < x1)
        row_img.copy_(torch.where(live, n.expand(n_img, hp, wp), torch.full((n_img, hp, wp), -1, device=dev)).reshape(-1).short())

    def in_stats(self, z, C, n_img, H, W, stats):
        self.launches += 1
        dev = z.t.device
        v = self._read(z, n_img, torch.arange(H, device=dev), torch.arange(W, device=dev), C)
        stats.copy_(torch.stack([v.sum((1, 2)), (v * v).sum((1, 2))], -1))

    def gather(self, src, dst, C, n_img, H, W, pad, halo_mode, tables=None, src2=None, res=None, stats=None, cnt=0.0,
               eps=1e-5, act=0, slope=0.0, dst_s2d=0):
        from irc_b200._native import IDENTITY
        self.launches += 1
        tables = tables or IDENTITY
        dev = src.t.device
        if stats is not None:
            mu, rs = self._moments(stats, cnt, eps)
            pre = lambda v: _act((v - mu) * rs, act, slope)
        elif act:
            pre = lambda v: _act(v, act, slope)
        else:
            pre = None
        acc = self._table_gather([src] + ([src2] if src2 is not None else []), n_img, H, W, C, tables, pre)
        ys = torch.arange(H, device=dev); xs = torch.arange(W, device=dev)
        if res is not None:
            acc = acc + self._read(res, n_img, ys, xs, C)
        Hp, Wp = H + 2 * pad, W + 2 * pad
        Yp = torch.arange(Hp, device=dev) - pad; Xp = torch.arange(Wp, device=dev) - pad
        if halo_mode == 1:
            full = acc[:, _reflect(Yp, H)][:, :, _reflect(Xp, W)]
        else:
            full = torch.zeros(n_img, Hp, Wp, C, device=dev)
            full[:, pad:pad + H, pad:pad + W] = acc
        if dst_s2d:
            n = torch.arange(n_img, device=dev).view(-1, 1, 1)
            Y = torch.arange(Hp, device=dev).view(1, -1, 1); X = torch.arange(Wp, device=dev).view(1, 1, -1)
            row = (n * (Hp >> 1) + (Y >> 1)) * (Wp >> 1) + (X >> 1)
            ch = dst.chan_off + ((Y & 1) * 2 + (X & 1)) * C
            c = torch.arange(C, device=dev)
            dst.t[row.unsqueeze(-1), ch.unsqueeze(-1) + c] = full.to(dst.t.dtype)
        else:
            self._write(dst, n_img, Yp, Xp, C, full)

    def in_apply(self, z, dst, C, n_img, H, W, pad, halo_mode, stats, eps=1e-5, act=0, slope=0.0, res=None, dst_s2d=0):
        self.in_stats(z, C, n_img, H, W, stats)
        self.gather(z, dst, C, n_img, H, W, pad, halo_mode, stats=stats, cnt=H * W, eps=eps, act=act, slope=slope, res=res, dst_s2d=dst_s2d)

    def bn_finalize(self, stats, n_img, group, C, cnt_per_img, gamma, beta, running_mean, running_var, eff, momentum=0.1, eps=1e-5, training=True,
                    updates=1):
        self.launches += 1
        M = group * cnt_per_img
        for g0 in range(0, n_img, group):
            if training:
                s = stats[g0:g0 + group].sum(0)                     # [C, 2]
                mean = s[:, 0] / M; var = (s[:, 1] / M - mean * mean).clamp_min(0)
                for _ in range(updates):
                    running_mean.mul_(1 - momentum).add_(momentum * mean)
                    running_var.mul_(1 - momentum).add_(momentum * var * (M / max(M - 1, 1)))
            else:
                mean, var = running_mean.clone(), running_var.clone()
            sc = gamma.reshape(-1) * torch.rsqrt(var + eps)
            sc = torch.where(sc.abs() < 1e-20, torch.full_like(sc, 1e-20), sc)
            eff[g0:g0 + group, :, 0] = mean - beta.reshape(-1) / sc
            eff[g0:g0 + group, :, 1] = sc

    def in_bwd(self, z, g1, dz, C, n_img, H, W, stats=None, cnt=0.0, eps=1e-5, act=0, slope=0.0, tables=None, g2=None, bsum=None,
               fold_pad=0, bn=None):
        from irc_b200._native import IDENTITY
        tables = tables or IDENTITY
        if fold_pad:
            self.fold_inplace(g1.t, g1.chan_off, C, n_img, H, W, fold_pad)
        self.launches += 2 if stats is not None else 1
        dev = z.t.device
        ys = torch.arange(H, device=dev); xs = torch.arange(W, device=dev)
        g = self._table_gather([g1] + ([g2] if g2 is not None else []), n_img, H, W, C, tables)
        zv = self._read(z, n_img, ys, xs, C)
        if stats is not None:
            mu, rs = self._moments(stats, cnt, eps)
            xh = (zv - mu) * rs
            gd = g * _dact(xh, act, slope)
            s1 = gd.sum((1, 2)); s2 = (gd * xh).sum((1, 2))
            if bn is not None:
                # BatchNorm: xh is the affine output y; group sums -> parameter gradients and the (A, B) pair of the apply formula
                grp = int(bn["group"]); ga = bn["gamma"].reshape(-1); be_ = bn["beta"].reshape(-1)
                ga = torch.where(ga.abs() < 1e-20, torch.full_like(ga, 1e-20), ga)
                dg = torch.zeros_like(ga); db = torch.zeros_like(ga)
                for g0 in range(0, n_img, grp):
                    S1 = s1[g0:g0 + grp].sum(0); S2 = s2[g0:g0 + grp].sum(0)
                    dgam = (S2 - be_ * S1) / ga
                    Bq = dgam / ga; Aq = S1 - be_ * Bq
                    s1[g0:g0 + grp] = Aq; s2[g0:g0 + grp] = Bq
                    dg += dgam; db += S1
                if bn.get("accumulate"):
                    bn["dgamma"].add_(dg.view_as(bn["dgamma"])); bn["dbeta"].add_(db.view_as(bn["dbeta"]))
                else:
                    bn["dgamma"].copy_(dg.view_as(bn["dgamma"])); bn["dbeta"].copy_(db.view_as(bn["dbeta"]))
            if bsum is not None:
                bsum.view(-1)[:n_img * C * 2] = torch.stack([s1, s2], -1).reshape(-1)
            o = rs * (gd - s1[:, None, None, :] / cnt - xh * s2[:, None, None, :] / cnt)
        else:
            o = g * _dact(zv, act, slope)
        self._write(dz, n_img, ys, xs, C, o)

    def fold_inplace(self, fr_t, chan_off, C, n_img, H, W, p):
        self.launches += 1
        Hp, Wp = H + 2 * p, W + 2 * p
        g = fr_t.view(n_img, Hp, Wp, -1)[..., chan_off:chan_off + C].float()
        dev = fr_t.device
        ry = _reflect(torch.arange(Hp, device=dev) - p, H); rx = _reflect(torch.arange(Wp, device=dev) - p, W)
        out = torch.zeros(n_img, H, W, C, device=dev)
        out.index_put_((torch.arange(n_img, device=dev).view(-1, 1, 1).expand(n_img, Hp, Wp), ry.view(1, -1, 1).expand(n_img, Hp, Wp),
                        rx.view(1, 1, -1).expand(n_img, Hp, Wp)), g, accumulate=True)
        fr_t.view(n_img, Hp, Wp, -1)[..., chan_off:chan_off + C] = 0
        fr_t.view(n_img, Hp, Wp, -1)[:, p:p + H, p:p + W, chan_off:chan_off + C] = out.to(fr_t.dtype)

    def maxpool2(self, src, dst, C, n_img, Ho, Wo):
        self.launches += 1
        dev = src.t.device
        v = self._read(src, n_img, torch.arange(2 * Ho, device=dev), torch.arange(2 * Wo, device=dev), C)
        o = v.view(n_img, Ho, 2, Wo, 2, C).amax((2, 4))
        self._write(dst, n_img, torch.arange(Ho, device=dev), torch.arange(Wo, device=dev), C, o)

    def maxpool2_bwd(self, src, g, dsrc, C, n_img, Ho, Wo):
        self.launches += 1
        dev = src.t.device
        v = self._read(src, n_img, torch.arange(2 * Ho, device=dev), torch.arange(2 * Wo, device=dev), C)
        gv = self._read(g, n_img, torch.arange(Ho, device=dev), torch.arange(Wo, device=dev), C)
        w = v.view(n_img, Ho, 2, Wo, 2, C).permute(0, 1, 3, 5, 2, 4).reshape(n_img, Ho, Wo, C, 4)
        arg = w.argmax(-1)  # first maximum
        mx = w.amax(-1)
        onehot = F.one_hot(arg, 4).float() * (mx > 0).float().unsqueeze(-1)
        o = (onehot * gv.unsqueeze(-1)).view(n_img, Ho, Wo, C, 2, 2).permute(0, 1, 4, 2, 5, 3).reshape(n_img, 2 * Ho, 2 * Wo, C)
        self._write(dsrc, n_img, torch.arange(2 * Ho, device=dev), torch.arange(2 * Wo, device=dev), C, o)

    def colsum(self, a, chan_off, C, out, row_img=None):
        self.launches += 1
        v = a[:, chan_off:chan_off + C].float()
        if row_img is not None:
            v = v * (row_img >= 0).float().unsqueeze(1)
        out.copy_(v.sum(0))

    # ------------------------------------------------------------------ degenerate convolutions
    def im2col_rows(self, row_mode, n_img, Ho, Wo):
        return n_img * Ho * Wo if row_mode == 0 else n_img * (Ho + 2) * (Wo + 2)

    @staticmethod
    def _rowmap(row_mode, n_img, Ho, Wo, dev):
        """flat row index [n, Ho, Wo] of every live output pixel"""
        n = torch.arange(n_img, device=dev).view(-1, 1, 1)
        oy = torch.arange(Ho, device=dev).view(1, -1, 1); ox = torch.arange(Wo, device=dev).view(1, 1, -1)
        if row_mode == 0:
            return (n * Ho + oy) * Wo + ox
        if row_mode == 1:
            return (n * (Ho + 2) + oy + 1) * (Wo + 2) + ox + 1
        hb, wb = (Ho + 2) >> 1, (Wo + 2) >> 1
        yp, xp = oy + 1, ox + 1
        return (((n * hb + (yp >> 1)) * wb + (xp >> 1)) << 2) + ((yp & 1) * 2 + (xp & 1))

    def im2col(self, src1, src2, scale, shift, n_img, H, W, k, stride, pad, pad_mode, Ho, Wo, row_mode, dst, row_img=None):
        self.launches += 1
        x = src1 if src2 is None else torch.cat([src1, src2], 1)
        if scale is not None:
            x = x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        C = x.shape[1]
        xp = F.pad(x, (pad,) * 4, mode="reflect" if pad_mode == 1 else "constant")
        cols = F.unfold(xp, k, stride=stride)                       # [n, C*k*k, Ho*Wo], order (c, r, s)
        cols = cols.view(n_img, C, k * k, Ho, Wo).permute(0, 3, 4, 2, 1).reshape(n_img, Ho, Wo, k * k * C)
        rm = self._rowmap(row_mode, n_img, Ho, Wo, x.device)
        dst.zero_()
        dst[rm.reshape(-1), :k * k * C] = cols.reshape(-1, k * k * C).to(dst.dtype)
        if row_img is not None:
            row_img.fill_(-1)
            row_img[rm.reshape(-1)] = torch.arange(n_img, device=x.device).view(-1, 1, 1).expand_as(rm).reshape(-1).short()

    direct_smallk = True

    def smallk_conv_fwd(self, src1, src2, scale, shift, n_img, H, W, k, stride, pad, pad_mode, Ho, Wo, row_mode, w, out, bias=None, act=0,
                        slope=0.0, E=None, row_img=None):
        e = torch.zeros_like(out) if E is None else E
        ri = torch.zeros(out.shape[0], dtype=torch.int16, device=out.device)
        self.im2col(src1, src2, scale, shift, n_img, H, W, k, stride, pad, pad_mode, Ho, Wo, row_mode, e, row_img=ri)
        v = e.float() @ w.float().t()
        if bias is not None:
            v = v + bias.view(1, -1)
        v = _act(v, act, slope) * (ri >= 0).view(-1, 1)
        out.copy_(v.to(out.dtype))
        if row_img is not None:
            row_img.copy_(ri)

    def col2im(self, de, C, c_first, c_out, n_img, H, W, k, stride, pad, Ho, Wo, row_mode, scale, out, accumulate):
        self.launches += 1
        rm = self._rowmap(row_mode, n_img, Ho, Wo, de.device)
        cols = de[rm.reshape(-1), :k * k * C].float().view(n_img, Ho * Wo, k * k, C).permute(0, 3, 2, 1).reshape(n_img, C * k * k, Ho * Wo)
        g = F.fold(cols, (H + 2 * pad, W + 2 * pad), k, stride=stride)[:, :, pad:pad + H, pad:pad + W]
        g = g[:, c_first:c_first + c_out]
        if scale is not None:
            g = g * scale[c_first:c_first + c_out].view(1, -1, 1, 1)
        if accumulate:
            out += g
        else:
            out.copy_(g)

    def tap_reduce(self, P, shifts, nco, n_img, H, W, hp, wp, oy, ox, bias, act, out):
        self.launches += 1
        dev = P.device
        n = torch.arange(n_img, device=dev).view(-1, 1, 1)
        y = torch.arange(H, device=dev).view(1, -1, 1); x = torch.arange(W, device=dev).view(1, 1, -1)
        q = (n * hp + y + oy) * wp + x + ox
        acc = torch.zeros(n_img, H, W, nco, device=dev)
        for j, (dy, dx) in enumerate(shifts):
            acc += P[(q + int(dy) * wp + int(dx)).reshape(-1), j * nco:(j + 1) * nco].view(n_img, H, W, nco)
        if bias is not None:
            acc = acc + bias.view(1, 1, 1, -1)
        if act == 3:
            acc = torch.tanh(acc)
        out.copy_(acc.permute(0, 3, 1, 2))

    def tap_expand(self, g, y, shifts, nco, n_img, H, W, hp, wp, oy, ox, E, dbias=None, live_cols_only=False):
        self.launches += 1
        dev = g.device
        gp = g if y is None else g * (1 - y * y)
        if dbias is not None:
            dbias.copy_(gp.sum((0, 2, 3)))
        rows = n_img * hp * wp
        img = torch.zeros(rows, nco, device=dev)
        n = torch.arange(n_img, device=dev).view(-1, 1, 1)
        yy = torch.arange(H, device=dev).view(1, -1, 1); xx = torch.arange(W, device=dev).view(1, 1, -1)
        q = ((n * hp + yy + oy) * wp + xx + ox).reshape(-1)
        img[q] = gp.permute(0, 2, 3, 1).reshape(-1, nco)
        E.zero_()
        allq = torch.arange(rows, device=dev)
        for j, (dy, dx) in enumerate(shifts):
            idx = allq - (int(dy) * wp + int(dx))
            ok = ((idx >= 0) & (idx < rows)).float().unsqueeze(1)
            E[:, j * nco:(j + 1) * nco] = (img[idx.clamp(0, rows - 1)] * ok).to(E.dtype)

    # ------------------------------------------------------------------ losses
    def pixel_loss(self, fake, target, w_l1, w_tvv, w_tvh, sums, dfake):
        self.launches += 1
        with torch.enable_grad():       # may be called from inside an autograd.Function.forward
            f = fake.detach().clone().requires_grad_(True)
            l1 = (f - target).abs().sum() if target is not None else f.sum() * 0
            tvv = (f[:, :, 1:] - f[:, :, :-1]).abs().sum(); tvh = (f[:, :, :, 1:] - f[:, :, :, :-1]).abs().sum()
            (g,) = torch.autograd.grad(w_l1 * l1 + w_tvv * tvv + w_tvh * tvh, f)
        sums[0] += l1.detach(); sums[1] += tvv.detach(); sums[2] += tvh.detach()
        if dfake is not None:
            dfake.copy_(g)

    @staticmethod
    def _ssim_map(x, y, window):
        window = window.to(x.device); n = window.numel(); C = x.shape[1]
        kx = window.view(1, 1, 1, n).expand(C, 1, 1, n); ky = window.view(1, 1, n, 1).expand(C, 1, n, 1)
        f = lambda t: F.conv2d(F.conv2d(t, kx, padding=(0, n // 2), groups=C), ky, padding=(n // 2, 0), groups=C)
        m1, m2 = f(x), f(y)
        s1 = f(x * x) - m1 * m1; s2 = f(y * y) - m2 * m2; s12 = f(x * y) - m1 * m2
        return ((2 * m1 * m2 + 1e-4) * (2 * s12 + 9e-4)) / ((m1 * m1 + m2 * m2 + 1e-4) * (s1 + s2 + 9e-4))

    def ssim_fwd(self, img1, img2, scale, shift, window, sums, ga=None, gb=None, gc=None):
        self.launches += 1
        m = self._ssim_map(img1 * scale + shift, img2 * scale + shift, window)
        sums += m.sum((1, 2, 3))
        self._ssim_saved = None

    def ssim_bwd(self, img1, img2, scale, shift, window, ga, gb, gc, coef, dimg1, accumulate):
        self.launches += 1
        with torch.enable_grad():
            a = img1.detach().clone().requires_grad_(True)
            m = self._ssim_map(a * scale + shift, img2 * scale + shift, window)
            (g,) = torch.autograd.grad(m.sum() * coef, a)
        if accumulate:
            dimg1 += g
        else:
            dimg1.copy_(g)

    def hinge(self, pred, n_real, mode, w_real, w_fake, sums, dpred):
        self.launches += 1
        p = pred.reshape(-1)
        if mode == 1:
            sums[2] += p.sum()
            if dpred is not None:
                dpred.fill_(-w_real)
            return
        r, f = p[:n_real], p[n_real:]
        sums[0] += torch.relu(1 - r).sum(); sums[1] += torch.relu(1 + f).sum()
        if dpred is not None:
            d = dpred.view(-1)
            d[:n_real] = torch.where(1 - r > 0, -w_real, 0.0)
            d[n_real:] = torch.where(1 + f > 0, w_fake, 0.0)

    def feat_l1(self, feat, rows_half, C, w, sums, dz):
        self.launches += 1
        f = feat[:rows_half, :C].float(); r = feat[rows_half:2 * rows_half, :C].float()
        sums[0] += (f - r).abs().sum()
        if dz is not None:
            dz[:rows_half, :C] = (w * torch.sign(f - r) * (f > 0).float()).to(dz.dtype)

    def quantize_metrics(self, fake, gt, u8, sums):
        self.launches += 1
        x = ((fake + 1.0) * 0.5).clamp(0, 1)
        q = (x * 255.0).to(torch.uint8)
        if u8 is not None:
            u8.copy_(q.permute(0, 2, 3, 1))
        if gt is not None:
            d = q.float() / 255.0 - gt
            sums.copy_(torch.stack([d.abs().double().sum((1, 2, 3)), (d * d).double().sum((1, 2, 3))], -1))

    # ------------------------------------------------------------------ optimizer / layout
    def ssim_metric(self, u8, gt, sums):
        """7x7 uniform-window SSIM in float64 over the valid region (skimage defaults), summed over channels and positions"""
        self.launches += 1
        x = gt.double(); y = (u8.permute(0, 3, 1, 2).float() / 255.0).double()
        box = lambda t: F.avg_pool2d(t, 7, 1)
        ux, uy, uxx, uyy, uxy = box(x), box(y), box(x * x), box(y * y), box(x * y)
        cn = 49.0 / 48.0
        vx, vy, vxy = cn * (uxx - ux * ux), cn * (uyy - uy * uy), cn * (uxy - ux * uy)
        S = ((2 * ux * uy + 1e-4) * (2 * vxy + 9e-4)) / ((ux * ux + uy * uy + 1e-4) * (vx + vy + 9e-4))
        sums.copy_(S.sum((1, 2, 3)))

    def accumulate(self, sums, coef, acc):
        self.launches += 1
        n = sums.numel()
        acc[:coef.shape[0]] += coef[:, :n].double() @ sums.double() + coef[:, n].double()

    def adam(self, p, g, m, v, hyper, step_dev):
        self.launches += 1
        lr, b1, b2, eps, lrs, gs = [float(t) for t in hyper.tolist()[:6]]
        step_dev += 1
        t = int(step_dev.item())
        lr, bc1, bc2 = lr * lrs, 1.0 - b1 ** t, 1.0 - b2 ** t
        gg = g * gs
        m.mul_(b1).add_(gg, alpha=1 - b1)
        v.mul_(b2).addcmul_(gg, gg, value=1 - b2)
        p.addcdiv_(m, (v.sqrt() / math.sqrt(bc2)).add_(eps), value=-lr / bc1)

    def gather_f32(self, src, map_, dst):
        self.launches += 1
        m = map_.long()
        dst.copy_(torch.where(m >= 0, src[m.clamp_min(0)], torch.zeros_like(src[m.clamp_min(0)])))

    def pack_bf16(self, src, map_, dst):
        self.launches += 1
        m = map_.long()
        dst.copy_(torch.where(m >= 0, src[m.clamp_min(0)], torch.zeros_like(src[m.clamp_min(0)])).to(dst.dtype))

    def gather_sum(self, src, map_, splits, split_stride, dst):
        self.launches += 1
        m = map_.long()
        acc = torch.zeros(m.numel(), device=src.device)
        for s in range(splits):
            acc += src[s * split_stride + m.clamp_min(0)]
        dst.copy_(torch.where(m >= 0, acc, torch.zeros_like(acc)))

    def gather_sum_deferred(self, src, map_, splits, split_stride, dst):
        self.gather_sum(src, map_, splits, split_stride, dst)

    def flush_sums(self):
        pass

    def stencil_nchw(self, x, out, tables, accumulate=False):
        self.launches += 1
        acc = torch.zeros_like(out)
        for i in range(tables.ky):
            for j in range(tables.kx):
                w = tables.ty_w[:, i].view(1, 1, -1, 1) * tables.tx_w[:, j].view(1, 1, 1, -1)
                acc += w * x[:, :, tables.ty_idx[:, i].long()][:, :, :, tables.tx_idx[:, j].long()]
        if accumulate:
            out += acc
        else:
            out.copy_(acc)

    def zero_(self, t):
        self.launches += 1
        t.zero_()
