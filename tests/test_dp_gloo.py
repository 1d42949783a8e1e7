"""Data-parallel path on CPU: world_size 2 over gloo.  Each rank runs the product's TrainStep (torch restatement of
the primitives, float32 storage) on its shard; the all-reduced gradients, averaged, must equal the gradients of the
single-process step on the concatenated batch (all losses are plain means and InstanceNorm is per sample)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import irc_oracle as O
    import irc_b200  # noqa: F401
    from irc_b200 import layout as L
    from irc_b200.train_step import TrainStep
    from ref_backend import RefBackend
    L.ACT_DTYPE = torch.float32
    H = W = 32
    pG = O.seeded_params(O.generator_shapes(), 1, bias_std=0.02)
    pD = O.seeded_params(O.discriminator_shapes(), 2, bias_std=0.02)
    pV = O.seeded_params(O.vgg_shapes(), 3, kaiming=True, bias_std=0.05)
    ir, rgb = O.synthetic_pair(world, H, W)
    ts = TrainStep(RefBackend(), 1, H, W, "cpu", world_size=world)
    ts.load(pG, pD, pV)
    ts.step(ir[rank:rank + 1].contiguous(), rgb[rank:rank + 1].contiguous())
    # every rank must hold identical parameters after the step
    flat = ts.G.arena.flat.clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        ref = TrainStep(RefBackend(), world, H, W, "cpu")
        ref.load(pG, pD, pV)
        # the D step of the DP run used all-reduced D gradients: same update => same D for the G step
        ref.step(ir, rgb)
        rel = lambda a, b: ((a - b).norm() / b.norm()).item()
        eG = rel(ts.G.arena.grad / world, ref.G.arena.grad)
        eD = rel(ts.D2.arena.grad / world, ref.D2.arena.grad)
        eP = (ts.G.arena.flat - ref.G.arena.flat).abs().max().item()
        torch.save(dict(same=same, eG=eG, eD=eD, eP=eP, hyper=ts.optG.dev.tolist()), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_equals_single_process_on_concatenated_batch(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["same"], "ranks diverged"
    assert r["eD"] < 1e-4 and r["eG"] < 1e-4, r
    assert r["eP"] < 4.1e-4, r          # Adam's first step is +-lr wherever the gradient sign is numerically fragile
    assert abs(r["hyper"][5] - 0.5) < 1e-12   # the 1/world average is folded into the Adam kernel


def _merge_worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import irc_b200  # noqa: F401
    from irc_b200.train import merge_test_rows, summarize_rows
    allrows = _fake_rows()
    mine = [r for r in allrows if r["_order"][0] % world == rank]        # run_test's sharding rule
    merged = merge_test_rows(mine)
    if rank == 0:
        torch.save(dict(files=[r["file"] for r in merged], summary=summarize_rows(merged)), out)
    dist.barrier()
    dist.destroy_process_group()


def _fake_rows():
    g = torch.Generator().manual_seed(3)
    rows = []
    for bi in range(5):
        for j in range(3):
            mse = float(torch.rand(1, generator=g)) * 0.1 + 1e-3
            rows.append({"file": f"img_{bi}_{j}.png", "mae": float(torch.rand(1, generator=g)), "mse": mse,
                         "psnr": -10.0 * __import__("math").log10(mse), "ssim": float(torch.rand(1, generator=g)), "_order": (bi, j)})
    return rows


@pytest.mark.timeout(300)
def test_sharded_test_mode_metrics_merge_to_the_single_process_summary(tmp_path):
    """config 5 across ranks: batches are sharded round-robin, the per-image rows are gathered in loader order and the
    running means (irc:1425-1431) equal those of one process over all images"""
    sys.path.insert(0, ROOT)
    import irc_b200  # noqa: F401
    from irc_b200.train import summarize_rows
    out = str(tmp_path / "merge.pt")
    mp.spawn(_merge_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    ref = _fake_rows()
    assert r["files"] == [x["file"] for x in ref]
    want = summarize_rows(ref)
    for k in want:
        assert abs(r["summary"][k] - want[k]) < 1e-12, k


# ---------------------------------------------------------------------------------------------------------------
# data parallelism through the PRODUCT entry point (train_kaist), as torchrun would launch it
# ---------------------------------------------------------------------------------------------------------------
def _train_cfg(tmp, steps=3, batch=1):
    import irc_b200 as R
    cfg = R.Config()
    cfg.mode, cfg.device, cfg.img_size, cfg.batch_size, cfg.epochs, cfg.synthetic_steps = "train", "cpu", 32, batch, 1, steps
    cfg.save_dir = os.path.join(tmp, "ckpt"); cfg.output_dir = os.path.join(tmp, "out")
    return cfg


def _kaist_worker(rank, world, port, tmp):
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    # exactly the environment torchrun provides; train_kaist joins the group itself
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    import irc_b200 as R
    from irc_b200 import layout as L, modules as M
    from irc_b200.train import train_kaist
    from ref_backend import RefBackend
    L.ACT_DTYPE = torch.float32
    M.set_backend(RefBackend())
    torch.manual_seed(100 + rank)            # every rank draws DIFFERENT initial weights: the broadcast must fix that
    cfg = _train_cfg(os.path.join(tmp, f"r{rank}"))
    hist = train_kaist(cfg, use_graph=False)
    ts = train_kaist.last_step
    assert dist.is_initialized() and dist.get_world_size() == world and cfg.world_size == world
    flats = [ts.G.arena.flat, ts.D2.arena.flat, ts.G.arena.m, ts.D2.arena.v]
    same = True
    for f in flats:
        parts = [torch.zeros_like(f) for _ in range(world)]
        dist.all_gather(parts, f.clone())
        same = same and all(torch.equal(parts[0], q) for q in parts)
    wrote = os.path.isdir(cfg.save_dir) and any(n.endswith(".pth") for n in os.listdir(cfg.save_dir))
    torch.save(dict(same=same, hist=hist, flatG=ts.G.arena.flat.clone(), flatD=ts.D2.arena.flat.clone(), wrote=wrote, steps=ts.optG.t,
                    step_dev=int(ts.optG.step_dev.item())), os.path.join(tmp, f"res{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_train_kaist_data_parallel_equals_single_process_on_the_global_batch(tmp_path):
    """2 ranks x batch 1 through train_kaist (process group joined from the torchrun environment, rank-0 weights broadcast,
    one global synthetic batch per step sliced across the ranks) == 1 process x batch 2 from rank 0's initial weights"""
    tmp = str(tmp_path)
    mp.spawn(_kaist_worker, args=(2, _free_port(), tmp), nprocs=2, join=True)
    r0, r1 = torch.load(os.path.join(tmp, "res0.pt")), torch.load(os.path.join(tmp, "res1.pt"))
    assert r0["same"] and r1["same"], "replicas diverged"
    assert r0["wrote"] and not r1["wrote"], "only rank 0 writes checkpoints"
    assert r0["steps"] == 3 and r0["step_dev"] == 3
    assert r0["hist"] == r1["hist"]                       # epoch means are all-reduced
    # single process, batch 2, same initial weights as rank 0
    sys.path.insert(0, ROOT)
    import irc_b200 as R
    from irc_b200 import layout as L, modules as M
    from irc_b200.train import train_kaist
    from ref_backend import RefBackend
    old, oldbe = L.ACT_DTYPE, M._BACKEND
    L.ACT_DTYPE = torch.float32
    M.set_backend(RefBackend())
    try:
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            os.environ.pop(k, None)
        torch.manual_seed(100)
        cfg = _train_cfg(os.path.join(tmp, "single"), batch=2)
        hist = train_kaist(cfg, use_graph=False)
        ts = train_kaist.last_step
    finally:
        L.ACT_DTYPE = old
        M.set_backend(oldbe)
    # three Adam steps of size lr = 2e-4: identical up to elements whose gradient sign is numerically fragile
    dG = (ts.G.arena.flat - r0["flatG"]).abs(); dD = (ts.D2.arena.flat - r0["flatD"]).abs()
    assert (dG > 1e-4).float().mean() < 2e-3 and (dD > 1e-4).float().mean() < 2e-3, ((dG > 1e-4).float().mean(), (dD > 1e-4).float().mean())
    for a, b in zip(hist[0], r0["hist"][0]):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (hist, r0["hist"])
