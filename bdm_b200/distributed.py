"""Sharding of a sampling job over ranks: one process per GPU, shapes partitioned contiguously, no
per-step communication; one all-gather of the finished clouds and one all-reduce of metric partials
at the end.  The reference samples whatever shard its dataloader hands each
Accelerate process and never communicates during sampling (main.py:111-120, :454-599); its only
collective is a 2-element all-reduce for meters (training_utils.py:130-141).

Works on any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_range(total, world_size=None, rank=None):
    """Contiguous [lo, hi) slice of `total` shapes owned by `rank`; remainders go to the first ranks."""
    if world_size is None or rank is None:
        world_size, rank = world()
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(seed, rank=None):
    """Per-rank generator seed, the reference's rule (training_utils.py:373-379: seed + rank)."""
    if rank is None:
        rank = world()[1]
    return seed + rank


def gather_samples(local, total):
    """local f32[B_local,N,3] on every rank -> f32[total,N,3] on every rank (ragged shards padded)."""
    ws, _ = world()
    if ws == 1:
        return local
    sizes = [shard_range(total, ws, r) for r in range(ws)]
    cap = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((cap,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    bucket = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bucket, pad)
    return torch.cat([bucket[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def reduce_metrics(cd, f1):
    """Per-shape CD and F-score of the local shard -> global means (sum CD, sum F, count all-reduced)."""
    part = torch.stack([cd.double().sum(), f1.double().sum(), cd.new_tensor(float(cd.numel())).double()])
    if world()[0] > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
    return float(part[0] / part[2]), float(part[1] / part[2]), int(part[2].item())
