#!/usr/bin/env python
"""Benchmark of the hot path: one full GAN training iteration (D update + G update with hinge +
L1 + VGG + TV + SSIM losses, irc:1636-1681) on synthetic 256x256 pairs, batch 16 per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Prints ONE JSON line (rank 0).  See DESIGN.md §measurement for what each key means."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

UNIT = "img/s"
GFLOP_PER_SAMPLE = 533.9          # minimal result-preserving conv work of one train step at 256^2


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tf=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), hbm=float(p["hbm_gbs"]), src="measured")
    except Exception:
        return dict(tf=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def oracle_step_time(B, H, W, steps, threads):
    """seconds per reference-algorithm train step (oracle port of irc:1636-1681) on the host cores"""
    import torch
    import irc_oracle as O
    torch.set_num_threads(threads)
    pG = O.seeded_params(O.generator_shapes(), 1234)
    pD = O.seeded_params(O.discriminator_shapes(), 1235)
    pV = O.seeded_params(O.vgg_shapes(), 1236, kaiming=True)
    aG, aD = O.AdamState(pG), O.AdamState(pD)
    ir, rgb = O.synthetic_pair(B, H, W)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.train_step(pG, pD, pV, aG, aD, ir, rgb)
        ts.append(time.perf_counter() - t0)
    return ts


WORKLOAD = ("single GAN train step (G+D, hinge+L1+VGG+TV+SSIM losses) at {H}x{W} batch {B} per GPU "
            "(BASELINE.json configs[1]; configs[2] when n_gpus > 1)")


def bench_config(args, world):
    """identical in both arms: what is measured, not how"""
    H, W = args.size
    return dict(workload=WORKLOAD.format(H=H, W=W, B=args.batch), global_batch=args.batch * world, parallelism=f"dp{world}")


def metric_name(args):
    H, W = args.size
    return f"gan_train_img_per_s_{H}x{W}_b{args.batch}_per_gpu"


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm (oracle port; the reference itself is a Python script that cannot
    travel to the GPU box) on all host cores.  EXACTLY --warmup + --steps steps are run; each step is a bounded sample of
    the workload (batch 2 instead of 16 - img/s is per image, and oneDNN's per-image cost does not fall with the batch), so
    that the default 5 + 20 steps end within a few minutes."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    Bs = 2
    H, W = args.size
    steps, warm = max(1, args.steps), max(0, args.warmup)
    ts = oracle_step_time(Bs, H, W, warm + steps, cores)[warm:]
    sec = sum(ts) / len(ts)
    val = Bs / sec
    line = dict(impl="reference", metric=metric_name(args), value=val, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warm,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=bench_config(args, args.gpus),
                cpu_baseline=dict(value=val, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{steps} step(s) (after {warm} warm-up) of batch {Bs} at {H}x{W}: the reference algorithm "
                                         f"(oracle/irc_oracle.py) on torch CPU fp32, {cores} threads; img/s is per image"),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                iters_per_s_at_sample_batch=1.0 / sec, gpu_launches=0)
    print(json.dumps(line), flush=True)


def eager_gpu_baseline(B, H, W, dev, steps=5, warm=2):
    """the stated kernel to beat: the reference's own op sequence in PyTorch eager + cuDNN on this
    GPU (oracle/eager_baseline.py), fp32 storage with the default TF32 convolutions, and bf16 autocast + channels_last"""
    import torch
    import irc_oracle as O
    import eager_baseline as EB
    pG = O.seeded_params(O.generator_shapes(), 1234); pD = O.seeded_params(O.discriminator_shapes(), 1235)
    pV = O.seeded_params(O.vgg_shapes(), 1236, kaiming=True)
    ir, rgb = O.synthetic_pair(B, H, W)
    ir, rgb = ir.to(dev), rgb.to(dev)
    out = {}
    for name, kw in (("tf32", dict()), ("bf16_autocast_channels_last", dict(channels_last=True, autocast=torch.bfloat16))):
        try:
            tr = EB.EagerTrainer(pG, pD, pV, dev, **kw)
            for _ in range(warm):
                tr.step(ir, rgb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                tr.step(ir, rgb)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = dict(ms_per_step=ms, img_per_s=B / (ms * 1e-3), steps=steps, warmup=warm)
            del tr
            torch.cuda.empty_cache()
        except Exception as ex:      # an out-of-memory eager baseline must not take the bench line with it
            out[name] = dict(error=str(ex)[:200])
            torch.cuda.empty_cache()
    out["note"] = ("PyTorch eager + cuDNN on the same GPU, the reference's op sequence as written (two generator forwards, D gradients also in "
                   "the G step): cudnn.allow_tf32 default (True) for 'tf32'; device-resident inputs, CUDA events")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, nargs=2, default=[256, 256], metavar=("H", "W"))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip timing the PyTorch eager + cuDNN restatement on the GPU")
    ap.add_argument("--skip", default="", help="timing ablation (results become wrong): comma-separated backend launchers turned into no-ops")
    ap.add_argument("--breakdown-all", default=None, help="write a per-launcher CUDA-event table of one eager step (all kernels) to this file")
    ap.add_argument("--breakdown", default=None, help="write the per-launch GEMM timing table to this file")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import irc_oracle as O          # weight / input recipes only; the timed path never touches it
    import irc_b200  # noqa: F401
    from irc_b200._native import CudaBackend
    from irc_b200.train_step import TrainStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly ONE JSON line: whatever native libraries print while the job runs (NCCL's version banner ...)
    # goes to stderr; file descriptor 1 is restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    B, (H, W) = args.batch, args.size
    W_, K = max(args.warmup, 3), args.steps

    be = CudaBackend()
    for name in filter(None, args.skip.split(",")):
        assert hasattr(be, name), name
        setattr(be, name, lambda *a, **k: None)
    ts = TrainStep(be, B, H, W, dev, world_size=world, use_graph=(not args.no_graph))
    ts.load(O.seeded_params(O.generator_shapes(), 1234), O.seeded_params(O.discriminator_shapes(), 1235),
            O.seeded_params(O.vgg_shapes(), 1236, kaiming=True))
    ir_h, rgb_h = O.synthetic_pair(B, H, W, rank)
    ir_h, rgb_h = ir_h.pin_memory(), rgb_h.pin_memory()
    ir_d, rgb_d = ir_h.to(dev), rgb_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    # ---- device-resident throughput
    for _ in range(W_):
        ts.step(ir_d, rgb_d)
    l0 = be.launches
    ts.step(ir_d, rgb_d)
    launches_per_step = getattr(ts, "launches_per_step", None) or (be.launches - l0)
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(lambda: ts.step(ir_d, rgb_d), K)
    clocks = sampler.stop() if sampler else None
    losses = ts.losses()

    # ---- end to end through the public call with host buffers: H2D of the batch + D2H of the losses every step
    # (public pipelined form: step() stages the pinned batch on a copy stream, losses_async() queues the 96-byte D2H read of
    # THIS step's loss sums; the host reads the previous step's handle, so the copy of step i + 1 overlaps the kernels of step i)
    pend = []
    def e2e_step():
        ts.step(ir_h, rgb_h)
        pend.append(ts.losses_async())
        if len(pend) > 1:
            pend.pop(0).get()
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, K)
    e2e_losses = pend.pop().get()
    assert all(v == v for v in e2e_losses.values())

    # ---- roofline of the dominant kernel: every tensor-core GEMM launch of one eager step bracketed by CUDA events
    pk = peaks()
    ts.use_graph = False
    ts.step(ir_d, rgb_d)
    be.timers = []
    ts.step(ir_d, rgb_d)
    torch.cuda.synchronize()
    rows = [(kind, name, role, fl, e0.elapsed_time(e1)) for (kind, name, role, fl, e0, e1) in be.timers]
    be.timers = None
    # SURVEY.md §8d classes: tensor-bound = every dense contraction (G down1/down2/resblocks/up1/up2, D model.2/5/8, VGG
    # conv1_2..conv3_3, in all three roles); the degenerate shapes (inc K=49, outc N=3, D model.0 K=64 / model.11 N=1, VGG
    # conv1_1 K=27) are HBM-bound layout passes and are reported against the HBM roofline instead
    DEGENERATE = {"G.inc", "G.outc", "D.0", "D.11", "V.0"}
    conv_all = [r for r in rows if r[0] == "conv_gemm" and r[3] > 0]
    conv = [r for r in conv_all if r[1] not in DEGENERATE]
    tn = [r for r in rows if r[0] == "tn_gemm" and r[3] > 0 and r[1] not in DEGENERATE]
    degen = [r for r in rows if r[1] in DEGENERATE]
    tfl = lambda rs: (sum(r[3] for r in rs) / (sum(r[4] for r in rs) * 1e-3) / 1e12) if rs else 0.0
    conv_tf, tn_tf, conv_all_tf = tfl(conv), tfl(tn), tfl(conv_all)
    gemm_ms = sum(r[4] for r in rows if r[0] in ("conv_gemm", "tn_gemm"))
    if args.breakdown and rank == 0:
        agg = {}
        for kind, name, role, fl, t in rows:
            a = agg.setdefault((kind, name, role), [0, 0.0, 0.0]); a[0] += 1; a[1] += fl; a[2] += t
        with open(args.breakdown, "w") as f:
            f.write("kernel,layer,role,launches,gflop_per_launch,ms_per_launch,tflops,frac_of_peak\n")
            for (kind, name, role), (n, fl, t) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
                f.write(f"{kind},{name},{role},{n},{fl / n / 1e9:.2f},{t / n:.4f},{fl / (t * 1e-3) / 1e12:.1f},{fl / (t * 1e-3) / 1e12 / pk['tf']:.3f}\n")

    graph_used = not args.no_graph

    def shutdown():
        # a captured graph that contains NCCL kernels must be released before the communicator goes away
        ts.graph = None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    # ---- every launcher of one eager step bracketed by CUDA events: the HBM-bound kernels against the measured copy peak
    # (all ranks run the steps - they contain the gradient all-reduces - rank 0 reports)
    hbm_kernels, dom = [], None
    be.time_all_launchers()
    ts.step(ir_d, rgb_d)
    be.timers_all = []
    ts.step(ir_d, rgb_d)
    torch.cuda.synchronize()
    agg = {}
    for name, tag, e0, e1 in be.timers_all:
        a = agg.setdefault((name, tag), [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1)
    be.timers_all = None
    tot = sum(v[1] for v in agg.values())
    if args.breakdown_all and rank == 0:
        with open(args.breakdown_all, "w") as f:
            f.write(f"launcher,args,calls,ms_total,ms_per_call,share   # one eager step, CUDA events per call, total {tot:.3f} ms\n")
            for (name, tag), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{name},\"{tag}\",{n},{t:.4f},{t / n:.4f},{t / tot:.3f}\n")
    # algorithmic bytes (SURVEY.md §8d, bf16 frames / fp32 images): what one ideal pass has to move
    px = B * H * W
    nG, nD = ts.G.arena.size, ts.D2.arena.size
    want = [
        ("Downsample fused with IN+ReLU (down1, 128 ch)", "gather", f"128,{B},{H // 2},{W // 2},1,0", (px + px // 4) * 128 * 2),
        ("UpsampleAA fused with IN+ReLU (up2, 128 ch)", "gather", f"128,{B},{H},{W},1,0", (px + px // 4) * 128 * 2),
        ("UpsampleAA^T (up2 backward)", "gather", f"128,{B},{H // 2},{W // 2},0,0", (px + px // 4) * 128 * 2),
        ("Downsample^T, two sources (down1 backward)", "gather", f"128,{B},{H},{W},0,0", (px + 2 * (px // 4)) * 128 * 2),
        ("InstanceNorm+ReLU backward, 128 ch full res (two passes move 5 units, 3 are algorithmic)", "in_bwd", f"128,{B},{H},{W}", 3 * px * 128 * 2),
        ("InstanceNorm statistics, 128 ch full res", "in_stats", f"128,{B},{H},{W}", px * 128 * 2),
        ("L1 + TV value and gradient", "pixel_loss", "", 3 * B * 3 * H * W * 4),
        ("SSIM forward (2 reads; the 3 saved maps are extra)", "ssim_fwd", "", 2 * B * 3 * H * W * 4),
        ("SSIM backward", "ssim_bwd", "", 3 * B * 3 * H * W * 4),
        ("Adam (G + D arenas, 28 B/param)", "adam", "", 28 * (nG + nD)),
    ]
    for label, name, tag, nbytes in want:
        if (name, tag) in agg:
            n, t = agg[(name, tag)]
            gbs = nbytes / (t * 1e-3) / 1e9
            row = dict(kernel=label, launches=n, algorithmic_mb=round(nbytes / 1e6, 1), us=round(t * 1e3, 1), gb_s=round(gbs, 0),
                       frac=round(gbs / pk["hbm"], 3))
            if name in ("ssim_fwd", "ssim_bwd"):
                # five (forward) / three (backward) separable 11-tap filters per pixel-channel: the fp32 pipe, not HBM, bounds them
                fma = (110 if name == "ssim_fwd" else 66) * B * 3 * H * W
                row.update(bound="fp32 pipe", tfma_s=round(fma / (t * 1e-3) / 1e12, 2),
                           note="110 / 66 FMA per pixel-channel put the floor 2.3x above the HBM floor; frac is against the HBM peak")
            hbm_kernels.append(row)
    if ("conv_gemm", "G.res:fwd") in agg:
        n, t = agg[("conv_gemm", "G.res:fwd")]
        fl = 2.0 * 256 * 256 * 9 * B * (H // 4) * (W // 4)
        dom = dict(kernel="conv_gemm_kernel, ResNet-block shape (M=B*H/4*W/4, N=256, K=2304): 18 of the conv launches, same shape as the "
                          "18 data-gradient launches (CTA-pair kernel, tcgen05.mma.cta_group::2)", launches=n, gflop_per_launch=fl / 1e9, us_per_launch=t / n * 1e3,
                   achieved=fl / (t / n * 1e-3) / 1e12, peak=pk["tf"], unit="TFLOP/s", frac=fl / (t / n * 1e-3) / 1e12 / pk["tf"],
                   traffic=None,
                   traffic_note="not measured in this run (DRAM counters need ncu): see profiles/ for the ncu --set full capture of this shape")

    if rank != 0:
        shutdown()
        return

    d2h_bytes = ts.sums.numel() * 4
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        oracle_step_time(2, H, W, 1, cores)                      # warm-up (oneDNN primitive caches)
        t = oracle_step_time(B, H, W, 1, cores)
        cpu_baseline = dict(value=B / t[0], unit=UNIT, cores=cores, kind="port",
                            sample=f"1 step of the full batch {B} at {H}x{W} (after a batch-2 warm-up step), oracle/irc_oracle.py, torch CPU fp32, "
                                   f"{cores} threads: {t[0]:.1f} s")
    eager = None
    if world == 1 and not args.no_eager_baseline:
        # release this repo's ~6 GB of frames first: the eager autograd graph of a batch-16 step needs tens of GB
        ts.graph = None
        ts.G = ts.D1 = ts.D2 = ts.V = None
        be._pending, be._sum_tables = [], {}
        torch.cuda.synchronize(); torch.cuda.empty_cache()
        eager = eager_gpu_baseline(B, H, W, dev)

    sec = ms * 1e-3 / K
    value = world * B / sec
    step_tflop = GFLOP_PER_SAMPLE * B * (H * W) / (256 * 256) / 1e3
    if eager:
        for k in ("tf32", "bf16_autocast_channels_last"):
            if "ms_per_step" in eager.get(k, {}):
                eager[k]["speedup_of_this_repo"] = eager[k]["ms_per_step"] / (sec * 1e3)
    line = dict(
        metric=metric_name(args), value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W_, ms_per_step=sec * 1e3, higher_is_better=True,
        scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
        config=bench_config(args, world),
        timing=dict(cuda_graph=bool(graph_used), events="CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks",
                    l2="per-step working set (several GB of activations) exceeds the 126 MB L2; no explicit flush"),
        iters_per_s=1.0 / sec, step_tflop_algorithmic=step_tflop, step_tensor_frac=step_tflop / sec / pk["tf"],
        e2e=dict(value=world * B / (ms_e2e * 1e-3 / K), unit=UNIT, h2d_bytes_per_step=ir_h.numel() * 4 + rgb_h.numel() * 4,
                 d2h_bytes_per_step=d2h_bytes),
        gpu_launches=launches_per_step * K,
        clocks=clocks,
        roofline=dict(bound="tensor", kernel="conv_gemm_kernel: every forward + data-gradient launch of the tensor-bound layers (SURVEY.md §8d classes)",
                      achieved=conv_tf, peak=pk["tf"], unit="TFLOP/s", frac=conv_tf / pk["tf"],
                      traffic=None, traffic_note="DRAM bytes are not observable from inside the run; ncu --set full captures are under profiles/",
                      peak_source=pk["src"] + " bf16_tflops_sustained", launches_per_step=len(conv),
                      share_of_step=sum(r[4] for r in conv) / (sec * 1e3),
                      all_conv_gemm_launches=dict(achieved=conv_all_tf, frac=conv_all_tf / pk["tf"], launches_per_step=len(conv_all),
                                                  note="same figure as round 1: includes the degenerate (HBM-bound) one-tap GEMMs")),
        roofline_wgrad=dict(bound="tensor", kernel="tn_gemm_kernel (weight gradients of the tensor-bound layers)", achieved=tn_tf, peak=pk["tf"],
                            unit="TFLOP/s", frac=tn_tf / pk["tf"], launches_per_step=len(tn), share_of_step=sum(r[4] for r in tn) / (sec * 1e3)),
        degenerate_convs=dict(layers=sorted(DEGENERATE), launches_per_step=len(degen), ms_per_step=sum(r[4] for r in degen),
                              note="HBM-bound by construction (K or N of a few units); GEMM launches only - their layout passes are in hbm_kernels"),
        roofline_dominant_launch=dom,
        hbm_kernels=hbm_kernels,
        hbm_peak_gb_s=pk["hbm"],
        gemm_ms_per_step=gemm_ms,
        cpu_baseline=cpu_baseline,
        gpu_eager_baseline=eager,
        losses={k: round(v, 5) for k, v in losses.items()},
    )
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    shutdown()


if __name__ == "__main__":
    main()
