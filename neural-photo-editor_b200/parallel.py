"""Data-parallel plumbing of the IAN hot path (one process per GPU, torch.distributed).

Inference BatchNorm keeps samples independent (reference IAN_simple.py graph, deterministic=True), so the path
shards by batch with replicated weights and has exactly one exchange step: an all-gather of the decoded images
(BASELINE north_star).  This module is backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of a batch of n samples for `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_images(local, n_total: int, group=None):
    """All-gather per-rank image shards (n_r,3,64,64) into the full (n_total,3,64,64) batch, in rank order.
    Equal shards use one all_gather_into_tensor; ragged shards pad to the largest shard."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d samples, expected %d" % (rank, local.shape[0], sizes[rank]))
    if len(set(sizes)) == 1:
        out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    m = max(sizes)
    padded = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)


def sharded_reconstruct(reconstruct_fn, x, group=None):
    """x: the FULL batch (same on every rank).  Each rank runs `reconstruct_fn` on its shard only, then the
    decoded shards are all-gathered.  Returns the full reconstructed batch on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return gather_images(reconstruct_fn(x[lo:hi]), x.shape[0], group)


def as_cuda_tensor(ptr: int, shape, device):
    """zero-copy torch view of a float32 device buffer owned by the library (the gather buffer)."""
    import torch

    class _Ptr:
        __cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Ptr(), device=device)
