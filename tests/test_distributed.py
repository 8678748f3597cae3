"""world_size-2 gloo test of the multi-rank host logic (sharding + scalar gather); the GPU path uses the same
functions over NCCL (bench.py)."""
import os

import numpy as np
import pytest

import ddpb_loader

ddpb_loader.load()  # (the package directory has a hyphen: loaded as ddp_pinocchio_b200)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import ddpb_loader
    ddpb_loader.load()
    from ddp_pinocchio_b200 import distributed as D, synthetic as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 7
    lo, hi = D.shard_range(total, rank, world)
    w = S.workload("c2", B=total, T=6)
    g_all = w.generate()
    g_loc = w.generate(B=hi - lo, first_problem=lo)
    assert np.array_equal(g_loc["x0"], g_all["x0"][lo:hi])        # shards of one global batch
    local = torch.tensor(np.stack([np.arange(lo, hi, dtype=float) + 100 * f for f in range(3)]))
    out = D.gather_scalars(local, world, rank, dist, max_per_rank=4)
    out_default = D.gather_scalars(local, world, rank, dist)      # ragged shards (4 + 3), pad width agreed collectively
    assert torch.equal(out, out_default)
    ok = D.all_converged(torch.tensor([1.0] * (hi - lo)) if rank == 0 else torch.tensor([0.0] * (hi - lo)), dist, world)
    q.put((rank, out.numpy(), ok))
    dist.destroy_process_group()


def test_shard_and_gather_gloo():
    import torch.multiprocessing as mp
    from ddp_pinocchio_b200 import distributed as D
    assert [D.shard_range(7, r, 2) for r in range(2)] == [(0, 4), (4, 7)]
    assert [D.shard_range(8192, r, 8) for r in range(8)][-1] == (7168, 8192)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, out, ok in res:
        assert out.shape == (3, 7)
        assert np.array_equal(out[0], np.arange(7.0)) and np.array_equal(out[2], np.arange(7.0) + 200)
        assert ok is False


def _solve_shard(w, g, B, emu):
    """The kernels' CTA bodies on the CPU (tests/hostemu) for one shard: 2 DDP iterations; returns scalars [fields][B] and x_T."""
    import numpy as np
    bt = emu.EmuBatch.from_workload(w, B=B)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    sc = bt.solve(2, w.mu)
    xs, _ = bt.get_trajectory()
    from ddp_pinocchio_b200 import capi
    return np.stack([sc[f] for f in capi.SCALAR_FIELDS[:10]]), xs


def _solve_worker(rank, world, port, q):
    import sys
    import torch
    import torch.distributed as dist
    import ddpb_loader
    ddpb_loader.load()
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import hostemu_lib as emu
    from ddp_pinocchio_b200 import distributed as D, synthetic as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 5
    lo, hi = D.shard_range(total, rank, world)
    w = S.workload("c2", B=total, T=6)
    sc, _ = _solve_shard(w, w.generate(B=hi - lo, first_problem=lo), hi - lo, emu)
    out = D.gather_scalars(torch.tensor(sc), world, rank, dist)
    q.put((rank, out.numpy()))
    dist.destroy_process_group()


def test_sharded_solve_equals_single_process_solve():
    """The multi-GPU path end to end on the CPU: two ranks (gloo) each solve their shard of one batch -- the kernels' CTA
    bodies through the host emulator -- and gather the per-problem scalars; every rank ends up with exactly what one
    process computes for the whole batch (problems are independent, shards ragged: 3 + 2)."""
    import sys
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import hostemu_lib as emu
    from ddp_pinocchio_b200 import synthetic as S
    w = S.workload("c2", B=5, T=6)
    ref, _ = _solve_shard(w, w.generate(), 5, emu)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + os.getpid() % 500
    ps = [ctx.Process(target=_solve_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, out in res:
        assert out.shape == ref.shape
        assert np.array_equal(out, ref), rank
