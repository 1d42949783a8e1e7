"""Train / test drivers of the hot path (irc:1333-1514 core, irc:1521-1542, irc:1549-1723, irc:1730-1748).

The reference reads KAIST image files through cv2/PIL (irc:803-1177); that I/O layer is outside the
accelerated path, so the drivers here take any iterable of `{'ir': Bx1xHxW, 'rgb': Bx3xHxW}`
batches in [-1,1] (the reference's DataLoader output format) or, with `cfg.synthetic_steps > 0`, generate
synthetic pairs of that format."""
from __future__ import annotations

import math
import os
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch

from . import modules as M
from .train_step import TrainStep


class SyntheticPairs:
    """Batches shaped and ranged like KAISTPairDataset + DataLoader (irc:1045-1177, :1576-1581).  Data-parallel: every
    step draws ONE global batch of world*B pairs from the seed and each rank yields its B-sample slice, so that an N-rank
    run sees exactly the batches of a single-process run with batch N*B (the property the DP parity tests rely on)."""

    sharded = True        # train_kaist must not stride this loader again

    def __init__(self, steps: int, B: int, H: int, W: int, seed: int = 7, rank: int = 0, world: int = 1):
        self.steps, self.B, self.H, self.W, self.seed, self.rank, self.world = steps, B, H, W, seed, rank, world

    def __len__(self):
        return self.steps

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        B, r = self.B, self.rank
        for i in range(self.steps):
            ir = torch.rand(self.world * B, 1, self.H, self.W, generator=g) * 2 - 1
            rgb = torch.rand(self.world * B, 3, self.H, self.W, generator=g) * 2 - 1
            yield {"ir": ir[r * B:(r + 1) * B].contiguous(), "rgb": rgb[r * B:(r + 1) * B].contiguous(),
                   "name": [f"synthetic_{i:05d}_{r * B + j}.png" for j in range(B)]}


class _Strided:
    """rank-strided view of a caller-supplied loader: batch i goes to rank i % world; the tail that does not fill a whole
    round is dropped so that every rank takes the same number of steps (each step ends in collective calls)"""

    sharded = True

    def __init__(self, loader, rank: int, world: int):
        self.loader, self.rank, self.world = loader, rank, world

    def __len__(self):
        return len(self.loader) // self.world

    def __iter__(self):
        n, pending = len(self), None
        it = iter(self.loader)
        for _ in range(n):
            for r in range(self.world):
                b = next(it)
                if r == self.rank:
                    pending = b
            yield pending


def init_distributed(cfg):
    """One process per GPU under torchrun: joins the NCCL (CUDA) / gloo (CPU) group from the RANK /
    WORLD_SIZE / LOCAL_RANK / MASTER_* environment, selects cuda:LOCAL_RANK and returns (rank, world).  A group the caller
    has already initialised is used as is.  Single process: (0, 1) and nothing is touched."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            cuda = str(cfg.device).startswith("cuda")
            if cuda:
                local = int(os.environ.get("LOCAL_RANK", str(rank % max(torch.cuda.device_count(), 1))))
                torch.cuda.set_device(local)
                dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            else:
                dist.init_process_group("gloo")
    if world > 1 and str(cfg.device).startswith("cuda"):
        cfg.device = f"cuda:{torch.cuda.current_device()}"
    cfg.world_size = world
    return rank, world


def broadcast_state(ts: TrainStep, src: int = 0) -> None:
    """Replicas must start from identical parameters and optimizer state (each rank draws its own init_weights sample,
    irc:168-209): rank `src` wins.  Also covers warm starts / resumes loaded on every rank."""
    import torch.distributed as dist
    if ts.world == 1:
        return
    for a in (ts.G.arena, ts.D2.arena, ts.V.arena):
        for t in (a.flat, a.m, a.v):
            dist.broadcast(t, src, group=ts.pg)
    for opt in (ts.optG, ts.optD):
        dist.broadcast(opt.step_dev, src, group=ts.pg)
        opt.t = int(opt.step_dev.item())
    ts.refresh_weights()


def tensor_to_rgb_image(t: torch.Tensor) -> np.ndarray:
    """irc:865-876 for one image (1x3xHxW or 3xHxW in [-1,1]) -> HxWx3 uint8, truncating; runs on the device."""
    x = t if t.dim() == 4 else t.unsqueeze(0)
    x = x[:1].contiguous().float()
    u8 = torch.empty(1, x.shape[2], x.shape[3], 3, device=x.device, dtype=torch.uint8)
    M.backend().quantize_metrics(x, None, u8, None)
    return u8[0].cpu().numpy()


def compute_metrics(pred_01, gt_01):
    """irc:1184-1217 on host arrays (HxWx3 float32 in [0,1]).  The SSIM metric needs scikit-image, which is a
    third-party dependency that is not vendored: None when it is not importable, exactly like the reference."""
    diff = pred_01 - gt_01
    mae = float(np.mean(np.abs(diff)))
    mse = float(np.mean(diff ** 2))
    psnr = float("inf") if mse == 0 else 20.0 * math.log10(1.0) - 10.0 * math.log10(mse + 1e-12)
    try:
        from skimage.metrics import structural_similarity as ssim
        ssim_val = float(ssim(gt_01, pred_01, data_range=1.0, channel_axis=2))
    except ImportError:
        ssim_val = None
    return mae, mse, psnr, ssim_val


def batch_metrics(fake: torch.Tensor, gt_01: torch.Tensor, with_ssim: bool = False):
    """Batched device version of tensor_to_rgb_image + compute_metrics: fake Bx3xHxW in [-1,1], gt Bx3xHxW in [0,1].
    Returns (uint8 BxHxWx3 predictions, per-image mae, mse, psnr lists) and - with_ssim - the per-image SSIM list
    (structural_similarity with the defaults the reference uses, irc:1208-1215, evaluated on the device)."""
    B, C, H, W = fake.shape
    u8 = torch.empty(B, H, W, C, device=fake.device, dtype=torch.uint8)
    sums = torch.zeros(B, 2, device=fake.device, dtype=torch.float64)
    gt = gt_01.contiguous().float()
    M.backend().quantize_metrics(fake.contiguous().float(), gt, u8, sums)
    ssum = None
    if with_ssim:
        ssum = torch.zeros(B, device=fake.device, dtype=torch.float64)
        M.backend().ssim_metric(u8, gt, ssum)
    s = (sums / (C * H * W)).cpu().numpy()
    mae = [float(np.float32(v)) for v in s[:, 0]]
    mse = [float(np.float32(v)) for v in s[:, 1]]
    psnr = [float("inf") if m == 0 else -10.0 * math.log10(m + 1e-12) for m in mse]
    if with_ssim:
        return u8, mae, mse, psnr, [float(v) for v in (ssum / (3 * (H - 6) * (W - 6))).cpu().numpy()]
    return u8, mae, mse, psnr


def validate_kaist(model: M.IRColorizationModel, val_loader: Iterable[Dict], device) -> float:
    """irc:1521-1542: sample-weighted mean L1 over the validation loader (over all ranks' shards when data-parallel)"""
    import torch.distributed as dist
    model.eval()
    total, count = 0.0, 0
    be = M.backend()
    with torch.no_grad():
        for batch in val_loader:
            ir = batch["ir"].to(device); rgb = batch["rgb"].to(device)
            fake = model(ir)
            sums = torch.zeros(3, device=fake.device)
            be.pixel_loss(fake.contiguous(), rgb.contiguous().float(), 0.0, 0.0, 0.0, sums, None)
            total += sums[0].item() / fake.numel() * ir.size(0)
            count += ir.size(0)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([total, float(count)], dtype=torch.float64, device=device)
        dist.all_reduce(t)
        total, count = t[0].item(), int(t[1].item())
    model.train()
    return total / max(count, 1)


def train_kaist(cfg: M.Config, train_loader=None, val_loader=None, use_graph: bool = True):
    """irc:1549-1723 with the fused D+G iteration.  Returns the list of per-epoch (avg_D, avg_G, val_L1).

    Data-parallel when launched by torchrun (one process per GPU): the process group is joined here, rank 0's initial
    weights are broadcast, the batches are sharded (synthetic: one global batch per step, sliced; caller-supplied loaders:
    rank-strided), gradients are all-reduced inside the step, the epoch means are averaged over the ranks, and only rank 0
    prints and writes checkpoints."""
    import torch.distributed as dist
    rank, world = init_distributed(cfg)
    device = torch.device(cfg.device)
    H = W = cfg.img_size
    say = print if rank == 0 else (lambda *a, **k: None)
    if train_loader is None:
        if getattr(cfg, "synthetic_steps", 0) <= 0:
            raise RuntimeError("No training data: pass train_loader/val_loader yielding {'ir','rgb'} batches, or set "
                               "cfg.synthetic_steps > 0 (the KAIST file reader, irc:1045-1177, is outside the accelerated path)")
        train_loader = SyntheticPairs(cfg.synthetic_steps, cfg.batch_size, H, W, seed=7, rank=rank, world=world)
        val_loader = SyntheticPairs(max(1, cfg.synthetic_steps // 10), cfg.batch_size, H, W, seed=8, rank=rank, world=world)
    elif world > 1:
        if not getattr(train_loader, "sharded", False):
            train_loader = _Strided(train_loader, rank, world)
        if val_loader is not None and not getattr(val_loader, "sharded", False):
            val_loader = _Strided(val_loader, rank, world)
    if rank == 0:
        os.makedirs(cfg.save_dir, exist_ok=True)

    model = M.IRColorizationModel(cfg)
    if cfg.init_G_weights is not None and os.path.isfile(cfg.init_G_weights):
        say(f"Loading initial generator weights from {cfg.init_G_weights}")
        model.load_weights(cfg.init_G_weights)
    netD = M.init_net(M.NLayerDiscriminator(cfg.input_nc + cfg.output_nc, 64, 3, M.get_norm_layer(cfg.norm)), "normal", 0.02, device)
    vgg = M.VGGPerceptual(device)
    lam = dict(L1=cfg.lambda_L1, perc=cfg.lambda_perc, tv=cfg.lambda_tv, ssim=cfg.lambda_ssim, gan=cfg.lambda_gan)
    ts = TrainStep(M.backend(), cfg.batch_size, H, W, device, cfg.lr_G, cfg.lr_D, cfg.beta1, cfg.beta2, lam, world_size=world,
                   use_graph=use_graph and device.type == "cuda", arenas=(model.netG.arena, netD.arena, vgg.arena),
                   no_antialias_up=cfg.no_antialias_up, no_antialias=cfg.no_antialias, norm=model.netG.norm,
                   bn_states=(model.netG._bn_state(), netD._bn_state()) if model.netG.norm == "batch" else (None, None))
    ts.refresh_weights()
    start_epoch = 1
    resume = getattr(cfg, "resume_from", None)          # not in the reference: full-state checkpoints (D + both Adam states + epoch)
    if resume and os.path.isfile(resume):
        st = torch.load(resume, map_location=device)
        ts.load_state_dict(st)
        start_epoch = int(st.get("epoch", 0)) + 1
        say(f"Resumed training state from {resume} (epoch {start_epoch - 1})")
    broadcast_state(ts)
    lr_lambda = M.get_lr_lambda(cfg)
    best_val, best_path = float("inf"), os.path.join(cfg.save_dir, "netG_best.pth")
    history = []
    for epoch in range(start_epoch, cfg.epochs + 1):
        scale = lr_lambda(epoch - 1)
        ts.reset_epoch_sums()
        for i, batch in enumerate(train_loader, start=1):
            ir = batch["ir"].to(device, non_blocking=True); rgb = batch["rgb"].to(device, non_blocking=True)
            ts.step(ir, rgb, lr_scale=scale)
            if i % 50 == 0 or i == 1:
                l = ts.losses()          # the only host synchronisation of the loop (rank-local values, like a per-rank print)
                say(f"Epoch [{epoch}/{cfg.epochs}] Step [{i}/{len(train_loader)}] D: {l['D']:.4f} | G: {l['G']:.4f} "
                    f"(GAN {l['GAN']:.4f} + L1 {l['L1']:.4f} + Perc {l['perc']:.4f} + TV {l['TV']:.6f} + SSIM {l['SSIM']:.4f})")
        avg_d, avg_g, _ = ts.epoch_means()      # every step of the epoch, accumulated on the device (irc:1683-1697)
        if model.netG.norm == "batch":
            # num_batches_tracked of the BatchNorm holders: two generator and three discriminator forward calls per iteration
            model.netG._bn_tick(2 * i); netD._bn_tick(3 * i)
        val_l1 = validate_kaist(model, val_loader, device) if val_loader is not None else float("nan")
        say(f"Epoch [{epoch}/{cfg.epochs}] DONE | avg D: {avg_d:.4f} | avg G: {avg_g:.4f} | val L1: {val_l1:.4f}")
        history.append((avg_d, avg_g, val_l1))
        if rank == 0 and ((epoch % cfg.save_every == 0) or (epoch == cfg.epochs)):
            path = os.path.join(cfg.save_dir, f"netG_epoch_{epoch:03d}.pth")
            torch.save(model.netG.state_dict(), path)
            say(f"Saved generator checkpoint to {path}")
            if getattr(cfg, "save_full_state", False):
                full = ts.state_dict(); full["epoch"] = epoch
                torch.save(full, os.path.join(cfg.save_dir, "train_state_latest.pth"))
        if val_l1 < best_val:
            best_val = val_l1
            if rank == 0:
                torch.save(model.netG.state_dict(), best_path)
            say(f"New best model saved to {best_path} (val L1={best_val:.4f})")
        say(f"Current LR (G): {cfg.lr_G * lr_lambda(epoch):.6e}")
    say(f"Training finished. Best val L1: {best_val:.4f}, best model: {best_path}")
    train_kaist.last_step = ts          # handle for tests / callers that want the final state
    if ts.graph is not None and world > 1:
        ts.graph = None                 # a captured graph holding NCCL kernels must go before the communicator does
        torch.cuda.synchronize()
    return history


def merge_test_rows(rows: List[Dict]) -> List[Dict]:
    """Data-parallel test mode (BASELINE config 5): every rank has evaluated the batches with
    index % world == rank; gather the per-image rows on all ranks in the loader's order.  Identity without a process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, rows)
    merged = [r for part in parts for r in part]
    merged.sort(key=lambda r: r["_order"])
    return merged


def summarize_rows(rows: List[Dict]) -> Optional[Dict]:
    """running means of irc:1425-1431 / :1467-1480 over the per-image rows"""
    if not rows:
        return None
    count = len(rows)
    have_ssim = any(r.get("ssim") is not None for r in rows)
    return dict(count=count, mean_mae=sum(r["mae"] for r in rows) / count, mean_mse=sum(r["mse"] for r in rows) / count,
                mean_psnr=sum(r["psnr"] for r in rows if np.isfinite(r["psnr"])) / count,
                mean_ssim=(sum(r["ssim"] for r in rows if r.get("ssim") is not None) / count) if have_ssim else None)


def write_topk_ranking(cfg: M.Config, rows: List[Dict]) -> Optional[str]:
    """The ranking CSV of save_best_k_outputs (irc:1236-1278): rows sorted by SSIM when any was computed, else by PSNR,
    descending (stable, like list.sort), top `cfg.topk`, same columns and number formats.  The file copies of
    irc:1280-1330 belong to the image-file layer that is outside the accelerated path."""
    if not rows:
        print("[TOP-K] metrics_list empty, skipping top-K save.")
        return None
    key = "ssim" if any(r.get("ssim") is not None for r in rows) else "psnr"
    valid = [r for r in rows if r.get(key) is not None and np.isfinite(r[key])]
    if not valid:
        print(f"[TOP-K] No valid '{key}' values, skipping top-K save.")
        return None
    valid.sort(key=lambda r: r[key], reverse=True)
    top = valid[:max(1, int(cfg.topk))]
    best_dir = os.path.join(cfg.output_dir, cfg.best50_dirname)
    os.makedirs(best_dir, exist_ok=True)
    path = os.path.join(best_dir, f"top_{len(top)}_ranking.csv")
    with open(path, "w", encoding="utf-8") as f:
        f.write("rank,file,mae,mse,psnr,ssim,metric_used\n")
        for i, m in enumerate(top, start=1):
            ssim_str = "" if m.get("ssim") is None else f"{m['ssim']:.6f}"
            f.write(f"{i},{m['file']},{m['mae']:.8f},{m['mse']:.8f},{m['psnr']:.6f},{ssim_str},{key}\n")
    print(f"[TOP-K] Ranking file   : {path}")
    return path


def run_test(cfg: M.Config, loader=None):
    """Test-mode core of irc:1333-1514: batched generator inference, on-device truncating quantisation and
    MAE/MSE/PSNR per image, `metrics_test.csv` in the reference's format.  Image files, collages and the top-K
    copies (irc:945-1038, :1220-1330) are outside the accelerated path."""
    rank, world = init_distributed(cfg)
    device = torch.device(cfg.device)
    H = W = cfg.img_size
    if loader is None:
        if getattr(cfg, "synthetic_steps", 0) <= 0:
            raise RuntimeError("No test data: pass a loader yielding {'ir', optional 'rgb', optional 'name'} batches or set "
                               "cfg.synthetic_steps > 0")
        loader = SyntheticPairs(cfg.synthetic_steps, cfg.batch_size, H, W, seed=9)
    os.makedirs(cfg.output_dir, exist_ok=True)
    model = M.IRColorizationModel(cfg)
    if cfg.test_G_weights and os.path.isfile(cfg.test_G_weights):
        model.load_weights(cfg.test_G_weights)
        print(f"Loaded generator weights from {cfg.test_G_weights}")
    else:
        print(f"Warning: generator weights not found at {cfg.test_G_weights}. Using randomly initialized model.")
    model.eval()
    rows: List[Dict] = []
    preds = []
    pending = []          # (pinned host buffer, event) of the predictions on their way back

    def shard():          # images are independent: shard the batches, no data-path collective
        for bi_, batch_ in enumerate(loader):
            if bi_ % world == rank:
                batch_ = dict(batch_); batch_["_bi"] = bi_
                yield batch_

    from .data import DevicePrefetcher
    with torch.no_grad():
        # the H2D copy of batch i + 1 overlaps the generator pass of batch i; the uint8 predictions travel back asynchronously
        for batch in DevicePrefetcher(shard(), device):
            bi = batch["_bi"]
            ir = batch["ir"].to(device)
            fake = model(ir)
            names = batch.get("name") or [f"img_{bi:05d}_{j}.png" for j in range(ir.shape[0])]
            if "rgb" in batch and batch["rgb"] is not None:
                gt01 = (batch["rgb"].to(device).float() + 1.0) * 0.5 if batch.get("rgb_range", "pm1") == "pm1" else batch["rgb"].to(device).float()
                # the SSIM column needs scikit-image in the reference (None without it, irc:1208-1215); here it is evaluated on
                # the device with the same defaults unless cfg.ssim_metric is False
                want_ssim = bool(getattr(cfg, "ssim_metric", True)) and min(fake.shape[2], fake.shape[3]) >= 7
                res = batch_metrics(fake, gt01, with_ssim=want_ssim)
                u8, mae, mse, psnr = res[:4]
                ssim_vals = res[4] if want_ssim else [None] * len(mae)
                for j, (n_, a, b, c, s_) in enumerate(zip(names, mae, mse, psnr, ssim_vals)):
                    rows.append({"file": n_, "mae": a, "mse": b, "psnr": c, "ssim": s_, "_order": (bi, j)})
            else:
                u8 = torch.empty(ir.shape[0], H, W, 3, device=device, dtype=torch.uint8)
                M.backend().quantize_metrics(fake.contiguous().float(), None, u8, None)
            if u8.is_cuda:
                host = torch.empty(u8.shape, dtype=u8.dtype, pin_memory=True)
                host.copy_(u8, non_blocking=True)
                ev = torch.cuda.Event(); ev.record()
                pending.append((host, ev, u8))            # u8 kept alive until its copy has run
            else:
                preds.append(u8)
    for host, ev, _ in pending:
        ev.synchronize()
        preds.append(host)
    rows = merge_test_rows(rows)
    print("Test finished.")
    summary = summarize_rows(rows)
    if summary is not None and rank == 0:
        count, mean_mae, mean_mse, mean_psnr = summary["count"], summary["mean_mae"], summary["mean_mse"], summary["mean_psnr"]
        print("\n=== Test Metrics (on images with GT) ===")
        print(f"Count      : {count}")
        print(f"Mean MAE   : {mean_mae:.6f}")
        print(f"Mean MSE   : {mean_mse:.6f}")
        print(f"Mean PSNR  : {mean_psnr:.4f} dB")
        mean_ssim = summary["mean_ssim"]
        print(f"Mean SSIM  : {mean_ssim:.6f}" if mean_ssim is not None else "Mean SSIM  : None (scikit-image not installed)")
        finite = [r for r in rows if np.isfinite(r["psnr"])]
        best = max(finite, key=lambda r: r["psnr"]) if finite else None        # first maximum, like the running `>` of irc:1437-1439
        print(f"Best PSNR  : {best['psnr']:.4f} ({best['file']})" if best else "Best PSNR  : N/A")
        with_s = [r for r in rows if r.get("ssim") is not None]
        best_s = max(with_s, key=lambda r: r["ssim"]) if with_s else None      # irc:1440-1442
        print(f"Best SSIM  : {best_s['ssim']:.6f} ({best_s['file']})" if best_s else "Best SSIM  : N/A")
        path = os.path.join(cfg.output_dir, "metrics_test.csv")
        with open(path, "w", encoding="utf-8") as f:
            f.write("file,mae,mse,psnr,ssim\n")
            for m in rows:
                ssim_str = "" if m.get("ssim") is None else f"{m['ssim']:.6f}"
                f.write(f"{m['file']},{m['mae']:.8f},{m['mse']:.8f},{m['psnr']:.6f},{ssim_str}\n")
            f.write("\n# Summary\n")
            ms = "" if mean_ssim is None else f"{mean_ssim:.6f}"
            f.write(f"# count,{count}\n# mean_mae,{mean_mae:.8f}\n# mean_mse,{mean_mse:.8f}\n# mean_psnr,{mean_psnr:.6f}\n# mean_ssim,{ms}\n")
        print(f"\nMetrics saved to: {path}")
        write_topk_ranking(cfg, rows)
    elif summary is None:
        print("No metrics were computed (no matching GT RGB images found).")
    return summary, rows, preds


def main(cfg: Optional[M.Config] = None):
    """irc:1730-1748"""
    cfg = cfg or M.Config()
    print(f"Running in mode: {cfg.mode}")
    print(f"Device: {cfg.device}")
    if cfg.mode == "train":
        return train_kaist(cfg)
    if cfg.mode == "test":
        return run_test(cfg)
    raise ValueError(f"Unknown mode: {cfg.mode}")
