"""Test-mode throughput, BASELINE.json configs[0] (B=1, 256x256, generator inference) and configs[4] (batched test-mode
inference + MAE/MSE/PSNR on 512x640 pairs, 64 images, sharded over the visible GPUs), through the module surface a user
of the reference calls: IRColorizationModel.forward + on-device truncating quantisation + metrics (irc:1381-1389,
:865-876, :1184-1205).

    python scripts/bench_infer.py [--out FILE]                                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node N scripts/bench_infer.py   # images sharded over N GPUs, no collective

One JSON line per config: `value` = images/s with the inputs resident in HBM; `e2e` = the same with pinned HOST inputs,
the H2D copy of every batch and the D2H copy of the uint8 predictions + per-image sums inside the timed region;
`roofline` = generator conv FLOPs (136.94 GFLOP/image at 256^2, x5 at 512x640) over the measured time
against the measured sustained bf16 peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.distributed as dist

import irc_b200 as R
from irc_b200.train import batch_metrics


def peak_tf():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured bf16_tflops_sustained"
    except Exception:
        return 1400.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = R.Config(); cfg.device = f"cuda:{local}"
    torch.manual_seed(0)
    model = R.IRColorizationModel(cfg).eval()
    pk, pk_src = peak_tf()
    lines = []
    # (label, images in total, per-rank batch, H, W)
    for label, total, H, W in (("configs[0]: test-mode generator inference, 1x256x256, batch 1", 1, 256, 256),
                               ("configs[4]: batched test-mode inference + MAE/MSE/PSNR, 512x640, 64 images", 64, 512, 640)):
        B = max(1, total // world) if total > 1 else 1
        g = torch.Generator().manual_seed(1 + rank)
        ir_h = (torch.rand(B, 1, H, W, generator=g) * 2 - 1).pin_memory(); gt_h = torch.rand(B, 3, H, W, generator=g).pin_memory()
        ir_d, gt_d = ir_h.to(dev), gt_h.to(dev)
        out_h = torch.empty(B, H, W, 3, dtype=torch.uint8).pin_memory()

        def resident():
            fake = model(ir_d)
            return batch_metrics(fake, gt_d)          # reads the per-image sums back (a few hundred bytes)

        out_ring = [out_h, torch.empty_like(out_h).pin_memory()]

        def e2e_loop(n):
            # the loop of run_test (train.py): batches staged one ahead on a copy stream (data.DevicePrefetcher), predictions
            # copied back asynchronously; the per-image metric sums are read on the host every batch
            from irc_b200.data import DevicePrefetcher
            evs = [None, None]
            for i, batch in enumerate(DevicePrefetcher([{"ir": ir_h, "rgb": gt_h}] * n, dev)):
                fake = model(batch["ir"])
                u8, mae, mse, psnr = batch_metrics(fake, batch["rgb"])
                if evs[i & 1] is not None:
                    evs[i & 1].synchronize()
                out_ring[i & 1].copy_(u8, non_blocking=True)
                evs[i & 1] = torch.cuda.Event(); evs[i & 1].record()
            torch.cuda.current_stream().synchronize()

        res = {}
        with torch.no_grad():
            for name, fn in (("resident", resident), ("e2e", None)):
                if fn is None:
                    e2e_loop(3)
                else:
                    for _ in range(3):
                        fn()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if fn is None:
                    e2e_loop(args.reps)
                else:
                    for _ in range(args.reps):
                        fn()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.reps
                if world > 1:
                    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
                res[name] = ms
        imgs = B * (world if total > 1 else 1)
        gflop = 136.94 * H * W / 65536
        tf = gflop * B / res["resident"]                # GFLOP per ms == TFLOP/s, per GPU
        lines.append(dict(metric="test_mode_img_per_s", value=imgs / (res["resident"] * 1e-3), unit="img/s", n_gpus=world if total > 1 else 1,
                          steps=args.reps, warmup=3, ms_per_step=res["resident"], higher_is_better=True, scaling="strong", vs_baseline=None,
                          dtype="bf16", data="synthetic",
                          config=dict(workload=label, images=imgs, per_gpu_batch=B, size=[H, W]),
                          e2e=dict(value=imgs / (res["e2e"] * 1e-3), unit="img/s", h2d_bytes_per_step=B * 4 * H * W * 4, d2h_bytes_per_step=B * H * W * 3 + B * 16),
                          roofline=dict(bound="tensor", achieved=tf, peak=pk, unit="TFLOP/s", frac=tf / pk, traffic=None, peak_source=pk_src,
                                        note="whole test-mode step (generator forward + quantise + metrics) against the conv FLOPs of the generator"),
                          cpu_reference_note="the reference's CPU path: 0.8-1.3 s per 256x256 image, 2.8 s per 512x640 image on 8 threads (SURVEY.md §8a-17)"))
    if rank == 0:
        for l in lines:
            print(json.dumps(l), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                for l in lines:
                    f.write(json.dumps(l) + "\n")
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
