"""Partitioning of independent units (particles / poses) across ranks -- SURVEY.md section 8e.

No arithmetic crosses particles or poses on this path (PCDSensorUpdaterEmbree.cpp:330-341), so multi-GPU is a contiguous
split with the map replicated per GPU and NO data-path collective; the only exchange is gathering the shards' results.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous [begin, end) of unit `rank` of `world`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_records(local: np.ndarray, dist, dst: int = 0):
    """Gather a structured numpy array (any record dtype) from all ranks onto `dst` with torch.distributed (gloo or nccl)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    raw = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([raw.size], dtype=torch.int64))
    m = int(max(s.item() for s in sizes))
    buf = torch.zeros(m, dtype=torch.uint8)
    buf[: raw.size] = torch.from_numpy(raw.copy())
    outs = [torch.zeros(m, dtype=torch.uint8) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst)
    if rank != dst:
        return None
    parts = [outs[r][: int(sizes[r].item())].numpy().view(local.dtype) for r in range(world)]
    return np.concatenate(parts)
