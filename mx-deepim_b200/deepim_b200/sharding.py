"""Batch sharding for multi-GPU inference: object instances are independent, so a batch
is split into contiguous per-rank slices and only the resulting poses are gathered.  One process per
GPU; no collective on the data path.  torch.distributed is used for the final gather only (backend
nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of n instances owned by `rank` (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def chunks(lo: int, hi: int, max_batch: int):
    """Split an owned slice into device batches of at most max_batch."""
    out = []
    while lo < hi:
        out.append((lo, min(hi, lo + max_batch)))
        lo = out[-1][1]
    return out


def gather_results(local: np.ndarray, n_total: int, axis: int = 0, dist=None, device=None):
    """All-gather per-rank result arrays (sharded along `axis` by shard_range) into the full array on
    every rank.  `dist` is torch.distributed (already initialised) or None for single-process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    loc = np.moveaxis(local, axis, 0)
    pad = np.zeros((mx,) + loc.shape[1:], loc.dtype)
    pad[: loc.shape[0]] = loc
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = np.concatenate([o.cpu().numpy()[: sizes[r]] for r, o in enumerate(outs)], axis=0)
    return np.moveaxis(full, 0, axis)
