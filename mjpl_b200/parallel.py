"""Sharding of independent validity work across the GPUs of one box.

The path shards by independent units (rows / edges / planning queries) with no data-path
collective: rank r owns the contiguous range
``[r*n/G, (r+1)*n/G)``.  A collective is used only when ONE logical batch was split and the
caller wants the whole mask back: an ``all_gather`` of the per-rank uint8 masks
(NCCL over NVLink for CUDA tensors, gloo for the CPU tests), plus scalar ``all_reduce`` of
counters.
"""

from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous ``[lo, hi)`` of ``n`` units owned by ``rank`` (sizes differ by at most 1)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, rem = divmod(int(n), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_masks(local_mask, n_total: int, group=None):
    """all_gather ragged per-rank masks into the full ``(n_total,)`` mask on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = local_mask if torch.is_tensor(local_mask) else torch.from_numpy(np.ascontiguousarray(local_mask))
    t = t.to(torch.uint8)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros(width, dtype=torch.uint8, device=t.device)
    padded[: t.numel()] = t
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)])


def reduce_counts(values, group=None, device=None):
    """Sum a small vector of counters over ranks."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(values, dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def sharded_sweep(engine, seed: int, n_total: int, flags: int, rank: int, world: int, gather: bool = False, group=None):
    """Validity sweep of ``n_total`` device-generated rows split over ranks by global row id."""
    lo, hi = shard_range(n_total, rank, world)
    mask = engine.sweep(seed, lo, hi - lo, flags)
    if gather and world > 1:
        return gather_masks(mask, n_total, group)
    return mask
