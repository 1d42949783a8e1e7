"""The fused D+G training iteration (irc:1629-1685) and the batched test-mode core
(irc:1381-1389, :865-876, :1197-1205) on top of the engines.

Result-preserving savings over the reference loop: the no-grad generator
forward of the D update (irc:1638) and the generator forward of the G update (irc:1657) are
the same computation (G is only updated at irc:1681), so it runs once; the G-step backward does
not compute discriminator weight gradients (irc:1636 discards them); vgg(rgb) keeps no
activations; D(real) and D(fake) of the D update run as one 2B-image pass."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import layout as L
from .engine import DiscriminatorEngine, GeneratorEngine, VggEngine

LAMBDAS = dict(L1=30.0, perc=30.0, tv=1e-4, ssim=2.0, gan=0.1)   # irc:100-104


def gaussian_window(n: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """irc:699-703, evaluated in fp32 exactly as the reference does.  Host tensor: the taps are passed to the
    SSIM kernels by value."""
    c = torch.arange(n, dtype=torch.float32) - (n - 1) / 2.0
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return (g / g.sum()).contiguous()


class AdamHyper:
    """Host side of the fused Adam (torch.optim.Adam defaults of irc:1601-1604).  The step counter lives on the DEVICE
    (`step_dev`, advanced by irc_adam itself), so a replayed CUDA graph needs no per-step host write; only the LambdaLR
    factor and the 1/world_size gradient scale travel from the host, and only when they change (per epoch), through a
    pageable staging tensor (the copy is staged by the driver before the call returns: the host cannot overwrite a buffer
    that an earlier, still queued copy has yet to read).  `t` mirrors the device counter on the host."""

    def __init__(self, device, lr: float, beta1: float, beta2: float, eps: float = 1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta1, beta2, eps, 0
        self.dev = torch.zeros(8, device=device, dtype=torch.float64)       # {lr, b1, b2, eps, lr_scale, grad_scale}
        self.step_dev = torch.zeros(1, device=device, dtype=torch.int64)
        self._staged = None

    def advance(self, lr_scale: float, grad_scale: float) -> None:
        """call once per optimizer step, before the step is enqueued"""
        self.t += 1
        want = (self.lr, self.b1, self.b2, self.eps, float(lr_scale), float(grad_scale))
        if want != self._staged:
            self.dev.copy_(torch.tensor(want + (0.0, 0.0), dtype=torch.float64))
            self._staged = want

    def set_step(self, t: int) -> None:
        self.t = int(t)
        self.step_dev.fill_(self.t)


class LossHandle:
    """result of TrainStep.losses_async(): the loss sums of one iteration on their way to the host"""

    def __init__(self, ts, buf, ev):
        self._ts, self._buf, self._ev, self._vals = ts, buf, ev, None

    def _materialize(self):
        if self._vals is None:
            if self._ev is not None:
                self._ev.synchronize()
            self._vals = self._buf.tolist()
            self._buf = None

    def get(self) -> Dict[str, float]:
        self._materialize()
        return self._ts._loss_terms(self._vals)


class TrainStep:
    """One process = one GPU = one replica.  With world_size > 1 the flat gradient arenas are
    all-reduced (sum) over NCCL and the 1/world_size average is folded into the Adam kernel."""

    def __init__(self, be, B: int, H: int, W: int, device, lr_G=2e-4, lr_D=2e-4, beta1=0.5, beta2=0.999, lambdas=LAMBDAS,
                 world_size: int = 1, process_group=None, use_graph: bool = False, arenas=(None, None, None), no_antialias_up: bool = False,
                 no_antialias: bool = False, norm: str = "instance", bn_states=(None, None)):
        self.be, self.B, self.H, self.W, self.dev = be, B, H, W, device
        self.lam = dict(lambdas)
        self.world, self.pg = world_size, process_group
        self.G = GeneratorEngine(be, B, H, W, device, arena=arenas[0], no_antialias_up=no_antialias_up, no_antialias=no_antialias, norm=norm)
        self.D2 = DiscriminatorEngine(be, 2 * B, H, W, device, arena=arenas[1], norm=norm)
        self.D1 = DiscriminatorEngine(be, B, H, W, device, arena=self.D2.arena, packer=self.D2.packer, layouts=self.D2.layouts, norm=norm)
        if norm == "batch":
            # nn.BatchNorm2d: the reference runs the generator twice per iteration on the same batch with the same weights
            # (irc:1638, :1657: two identical running-statistics updates) and the discriminator three times (irc:1642, :1643,
            # :1659: real and fake halves are separate batches); bn_states = the module surface's registered buffers
            self.G.bn_updates = 2
            if bn_states[0] is not None:
                self.G.bn_state = bn_states[0]
            if bn_states[1] is not None:
                self.D2.bn_state = bn_states[1]
            self.D1.bn_state = self.D2.bn_state
        self.V = VggEngine(be, 2 * B, B, H, W, device, arena=arenas[2])
        self.optG = AdamHyper(device, lr_G, beta1, beta2)
        self.optD = AdamHyper(device, lr_D, beta1, beta2)
        self.window = gaussian_window()
        self.sums = torch.zeros(8 + B, device=device)
        # running sums of the D and G losses over the steps since reset_epoch_sums(): the reference averages EVERY step of an
        # epoch (irc:1683-1697); accumulating on the device keeps the loop free of per-step host synchronisation
        self.acc = torch.zeros(4, device=device, dtype=torch.float64)          # {sum D, sum G, steps, -}
        self.coef = self._loss_coefficients().to(device)
        self.dfake = torch.zeros(B, 3, H, W, device=device)
        self.ga, self.gb, self.gc = (torch.zeros(B, 3, H, W, device=device) for _ in range(3))
        self.ir = torch.zeros(B, 1, H, W, device=device)
        self.rgb = torch.zeros(B, 3, H, W, device=device)
        # NCCL collectives are capturable (torch's ProcessGroupNCCL records them on its own stream with event edges), so the
        # data-parallel step replays as one graph too
        self.use_graph = use_graph
        self.graph = None
        self._side = None
        self.n_pred = self.D1.pred[0].numel()

    def _loss_coefficients(self) -> torch.Tensor:
        """D and G losses are affine in the raw sums of one iteration: row j = (coefficients of sums[0:8+B], constant).
        Third row counts the steps."""
        B, H, W, lam = self.B, self.H, self.W, self.lam
        n = 8 + B
        cnt = B * self.D1.pred[0].numel()
        npix = B * 3 * H * W
        c = torch.zeros(3, n + 1, dtype=torch.float32)
        c[0, 0] = c[0, 1] = 0.5 / cnt                                      # irc:1647-1649
        c[1, 2] = -lam["gan"] / cnt                                        # irc:1662, :1679
        c[1, 3] = lam["L1"] / npix                                         # irc:1664
        c[1, 4] = lam["tv"] / (B * 3 * (H - 1) * W); c[1, 5] = lam["tv"] / (B * 3 * H * (W - 1))     # irc:686-694
        c[1, 6] = lam["perc"] / (B * 256 * (H // 4) * (W // 4))            # irc:1669
        c[1, 8:8 + B] = -lam["ssim"] / npix; c[1, n] = lam["ssim"]         # irc:1675-1677
        c[2, n] = 1.0
        return c.contiguous()

    def reset_epoch_sums(self) -> None:
        self.acc.zero_()

    def epoch_means(self):
        """(mean D loss, mean G loss, steps) over the iterations since reset_epoch_sums(), averaged over the ranks
        (irc:1683-1697).  Synchronises."""
        acc = self.acc.clone()
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(acc, group=self.pg)
        a = acc.tolist()
        steps = max(a[2], 1.0)
        return a[0] / steps, a[1] / steps, int(round(a[2] / self.world))

    # ------------------------------------------------------------------ parameters
    def load(self, pG: Dict[str, torch.Tensor], pD: Dict[str, torch.Tensor], pV: Dict[str, torch.Tensor]) -> None:
        self.G.arena.load(pG); self.D2.arena.load(pD); self.V.arena.load(pV)
        self.refresh_weights()

    def refresh_weights(self) -> None:
        self.G.refresh_weights(); self.D2.refresh_weights(); self.V.refresh_weights()

    # ------------------------------------------------------------------ full-state checkpoints (resume)
    def state_dict(self) -> Dict:
        """Everything needed to resume training bit-exactly: generator and discriminator parameters under the reference's
        state_dict keys (`netG` alone is what irc:1706-1715 saves), both Adam states and step counters.  The reference has no
        resume path; `netG` stays loadable by its `load_weights` (irc:781-789)."""
        out = {"format": "irc_b200.train_state.v1",
               "netG": {k: v.detach().clone() for k, v in self.G.arena.state_dict().items()},
               "netD": {k: v.detach().clone() for k, v in self.D2.arena.state_dict().items()}}
        for name, arena, opt in (("optG", self.G.arena, self.optG), ("optD", self.D2.arena, self.optD)):
            out[name] = dict(step=opt.t, lr=opt.lr, betas=(opt.b1, opt.b2), eps=opt.eps, exp_avg=arena.m.detach().clone(),
                             exp_avg_sq=arena.v.detach().clone())
        for name, eng in (("bnG", self.G), ("bnD", self.D2)):          # running statistics of nn.BatchNorm2d networks (norm='batch')
            if getattr(eng, "bn", False):
                out[name] = {k: (rm.detach().clone(), rv.detach().clone()) for k, (rm, rv) in eng.bn_state.items()}
        return out

    def load_state_dict(self, state: Dict) -> None:
        if state.get("format") != "irc_b200.train_state.v1":
            raise ValueError("not an irc_b200 training-state checkpoint (for generator-only checkpoints use IRColorizationModel.load_weights)")
        self.G.arena.load(state["netG"]); self.D2.arena.load(state["netD"])
        for name, arena, opt in (("optG", self.G.arena, self.optG), ("optD", self.D2.arena, self.optD)):
            st = state[name]
            if st["exp_avg"].numel() != arena.m.numel():
                raise ValueError(f"{name}: optimizer state of {st['exp_avg'].numel()} elements does not fit the arena ({arena.m.numel()})")
            arena.m.copy_(st["exp_avg"]); arena.v.copy_(st["exp_avg_sq"])
            opt.set_step(int(st["step"])); opt._staged = None; opt.lr = float(st["lr"]); opt.b1, opt.b2 = (float(b) for b in st["betas"]); opt.eps = float(st["eps"])
        for name, eng in (("bnG", self.G), ("bnD", self.D2)):
            for k, (rm, rv) in state.get(name, {}).items():
                eng.bn_state[k][0].copy_(rm); eng.bn_state[k][1].copy_(rv)
        self.refresh_weights()

    # ------------------------------------------------------------------ the iteration
    def _allreduce_async(self, t: torch.Tensor):
        """sum-allreduce over NCCL/NVLink on NCCL's own stream; returns a handle to wait on (None when single GPU)"""
        if self.world == 1:
            return None
        import torch.distributed as dist
        return dist.all_reduce(t, group=self.pg, async_op=True)

    @staticmethod
    def _wait(*works) -> None:
        for w in works:
            if w is not None:
                w.wait()          # stream-level wait, the host does not block

    def _body(self) -> None:
        be, B, H, W, lam = self.be, self.B, self.H, self.W, self.lam
        ir, rgb = self.ir, self.rgb
        be.zero_(self.sums)
        fake = self.G.forward(ir)                                             # irc:1638 and irc:1657
        # ---------------- D update (irc:1636-1651)
        pred = self.D2.forward(ir, rgb, ir, fake)                              # irc:1639-1643
        cnt = B * self.n_pred
        be.hinge(pred, cnt, 0, 0.5 / cnt, 0.5 / cnt, self.sums[0:3], self.D2.dpred)   # irc:1647-1649
        self.D2.backward(self.D2.dpred, True)                                  # irc:1650
        wD = self._allreduce_async(self.D2.arena.grad)
        # ---------------- G update, the part that does not involve D (overlaps the D-gradient all-reduce)
        npix = B * 3 * H * W
        be.pixel_loss(fake, rgb, lam["L1"] / npix, lam["tv"] / (B * 3 * (H - 1) * W), lam["tv"] / (B * 3 * H * (W - 1)),
                      self.sums[3:6], self.dfake)                              # irc:1664, :1672
        be.ssim_fwd(fake, rgb, 0.5, 0.5, self.window, self.sums[8:8 + B], self.ga, self.gb, self.gc)   # irc:1675-1677
        be.ssim_bwd(fake, rgb, 0.5, 0.5, self.window, self.ga, self.gb, self.gc, -lam["ssim"] / npix, self.dfake, True)
        feat = self.V.forward(fake, rgb)                                       # irc:1667-1668
        nfeat = B * 256 * (H // 4) * (W // 4)
        be.feat_l1(feat.t, feat.rows_of(B), 256, lam["perc"] / nfeat, self.sums[6:7], self.V.dz[-1].t)   # irc:1669
        self.V.backward(self.dfake)
        # ---------------- D optimizer step, then the adversarial term with the updated D (irc:1651, :1659-1662)
        self._wait(wD)
        A = self.D2.arena
        be.adam(A.flat, A.grad, A.m, A.v, self.optD.dev, self.optD.step_dev)                       # irc:1651
        self.D2.refresh_weights()
        pred_f = self.D1.forward(ir, fake, keep_operand=False)                                     # irc:1659
        be.hinge(pred_f, 0, 1, lam["gan"] / cnt, 0.0, self.sums[0:3], self.D1.dpred)   # irc:1662, :1679
        self.D1.backward(self.D1.dpred, False, self.dfake)
        # ---------------- generator backward (irc:1680); gradients of everything past the encoder are final once
        # the ResNet blocks are done, so their all-reduce overlaps the full-resolution encoder backward
        A = self.G.arena
        split = A.offset["resblocks.0.conv_block.1.weight"]
        works = []
        self.G.backward(self.dfake, after_blocks=lambda: works.append(self._allreduce_async(A.grad[split:])))
        works.append(self._allreduce_async(A.grad[:split]))
        self._wait(*works)
        be.adam(A.flat, A.grad, A.m, A.v, self.optG.dev, self.optG.step_dev)                       # irc:1681
        self.G.refresh_weights()
        be.accumulate(self.sums, self.coef, self.acc)                          # irc:1683-1685

    def step(self, ir: torch.Tensor, rgb: torch.Tensor, lr_scale: float = 1.0) -> None:
        """One iteration on a device-resident batch (fp32 NCHW in [-1,1])."""
        if ir.device.type == "cpu" and self.ir.is_cuda:
            self._stage_host_batch(ir, rgb)
        else:
            self.ir.copy_(ir, non_blocking=True); self.rgb.copy_(rgb, non_blocking=True)
        gs = 1.0 / self.world
        self.optD.advance(lr_scale, gs); self.optG.advance(lr_scale, gs)
        if not self.use_graph:
            self._body()
            return
        if self.graph is None:
            # warm-up on a side stream (sets kernel attributes, primes allocators), undo its effect on
            # the optimizer state, then capture
            state = [t for a in (self.G.arena, self.D2.arena) for t in (a.flat, a.m, a.v)] + [self.optG.step_dev, self.optD.step_dev, self.acc]
            for eng in (self.G, self.D2):          # BatchNorm running statistics are advanced by the warm-up run as well
                state += [t for pair in getattr(eng, "bn_state", {}).values() for t in pair]
            snap = [t.clone() for t in state]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._body()
            torch.cuda.current_stream().wait_stream(s)
            for t, c in zip(state, snap):
                t.copy_(c)
            self.refresh_weights()
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            n0 = self.be.launches
            with torch.cuda.graph(self.graph):
                self._body()
            self.launches_per_step = self.be.launches - n0
        self.graph.replay()

    def _stage_host_batch(self, ir: torch.Tensor, rgb: torch.Tensor) -> None:
        """Host (pinned) batch: the H2D copy runs on a copy stream into one of two staging buffers, so the copy of iteration
        i + 1 overlaps the kernels of iteration i whenever the caller does not synchronise in between (`losses_async`); the
        iteration's own stream then waits for the copy and moves the batch into the buffers the captured graph reads (a 17 MB
        device-to-device copy, ~6 us).  Usual non_blocking contract: the host tensors must stay untouched until the copy ran."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage = [(torch.empty_like(self.ir), torch.empty_like(self.rgb)) for _ in range(2)]
            self._stage_ready = [torch.cuda.Event() for _ in range(2)]
            self._stage_free = [torch.cuda.Event() for _ in range(2)]
            self._stage_k = 0
            for e in self._stage_free:
                e.record()
        k = self._stage_k
        self._stage_k ^= 1
        main, cs = torch.cuda.current_stream(), self._copy_stream
        cs.wait_event(self._stage_free[k])          # the iteration that used this staging pair two calls ago has read it
        with torch.cuda.stream(cs):
            self._stage[k][0].copy_(ir, non_blocking=True); self._stage[k][1].copy_(rgb, non_blocking=True)
            self._stage_ready[k].record(cs)
        main.wait_event(self._stage_ready[k])
        self.ir.copy_(self._stage[k][0], non_blocking=True); self.rgb.copy_(self._stage[k][1], non_blocking=True)
        self._stage_free[k].record(main)

    def losses_async(self) -> "LossHandle":
        """Queue the device-to-host read of this iteration's loss sums (pinned ring slot + event) and return a handle; `get()`
        on it synchronises with that read only.  A loop that calls `step(); h = losses_async()` and reads the previous handle
        keeps one iteration of work queued, so host-side batch preparation and the H2D copy overlap the kernels."""
        if not self.sums.is_cuda:
            return LossHandle(self, self.sums.clone(), None)
        if getattr(self, "_loss_ring", None) is None:
            self._loss_ring = [[torch.empty(self.sums.shape, dtype=self.sums.dtype).pin_memory(), torch.cuda.Event(), None] for _ in range(4)]
            self._loss_i = 0
        slot = self._loss_ring[self._loss_i % len(self._loss_ring)]
        self._loss_i += 1
        if slot[2] is not None:
            slot[2]._materialize()          # an unread handle still owns this pinned slot: take its values out first
        buf, ev = slot[0], slot[1]
        buf.copy_(self.sums, non_blocking=True)
        ev.record()
        slot[2] = LossHandle(self, buf, ev)
        return slot[2]

    def losses(self) -> Dict[str, float]:
        """Loss terms of the last iteration, as printed by the reference (irc:1687-1694).  Synchronises."""
        return self._loss_terms(self.sums.tolist())

    def _loss_terms(self, s) -> Dict[str, float]:
        B, H, W, lam = self.B, self.H, self.W, self.lam
        cnt = B * self.n_pred
        npix = B * 3 * H * W
        out = dict(D=0.5 * (s[0] / cnt + s[1] / cnt), GAN=-s[2] / cnt, L1=lam["L1"] * s[3] / npix,
                   TV=lam["tv"] * (s[4] / (B * 3 * (H - 1) * W) + s[5] / (B * 3 * H * (W - 1))),
                   perc=lam["perc"] * s[6] / (B * 256 * (H // 4) * (W // 4)),
                   SSIM=lam["ssim"] * (1.0 - sum(s[8:8 + B]) / npix))
        out["G"] = lam["gan"] * out["GAN"] + out["L1"] + out["perc"] + out["TV"] + out["SSIM"]
        return out
