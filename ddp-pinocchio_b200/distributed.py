"""Multi-GPU host logic: the problem batch shards across ranks (one process per GPU, independent problems,
no data-path collective); the only exchange is a gather of the per-problem scalars {cost, merit, opt_obj,
opt_constr, step, status, ...}.  torch.distributed is plumbing only: NCCL on GPUs,
gloo in the CPU tests."""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of problems owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_scalars(local, world, rank, dist=None, max_per_rank=None):
    """All-gather the field-major scalar block [F][B_local] of every rank into [F][B_total] (problem order =
    rank order).  `local` is a torch tensor on the rank's device; ragged shards are padded to `max_per_rank`."""
    import torch
    if world == 1:
        return local
    F, Bl = local.shape
    # shard_range produces ragged shards (ranks differ by one problem): the pad width must be the same on every
    # rank, so it is agreed on collectively (max of the per-rank counts) unless the caller passes an upper bound
    counts = torch.tensor([Bl], dtype=torch.int64, device=local.device)
    cl = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(cl, counts)
    sizes = [int(c.item()) for c in cl]
    m = max(sizes) if max_per_rank is None else int(max_per_rank)
    if max(sizes) > m:
        raise ValueError(f"gather_scalars: a shard holds {max(sizes)} problems, more than max_per_rank={m}")
    pad = torch.zeros((F, m), dtype=local.dtype, device=local.device)
    pad[:, :Bl] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:, :c] for o, c in zip(out, sizes)], dim=1)


def all_converged(status_local, dist=None, world=1):
    """ncclAllReduce(min) of 'every problem of this shard has status != 0' for early exit."""
    import torch
    flag = torch.tensor([1 if bool((status_local != 0).all()) else 0], dtype=torch.int32, device=status_local.device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())
