"""Partitioning of independent units (particles / poses) across ranks -- SURVEY.md section 8e.

No arithmetic crosses particles or poses on this path (PCDSensorUpdaterEmbree.cpp:330-341), so multi-GPU is a contiguous
split with the map replicated per GPU and NO data-path collective; the only exchange is gathering the shards' results.
The one stage of the particle-filter cycle with a real exchange is resampling (opponents are drawn from ALL particles,
resampling.cu:137): `gladiator_resample_sharded` all-gathers the particle set first.
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


def gladiator_resample_sharded(resample_fn, poses_local, attrs_local, dist, sync=None):
    """Gladiator resampling with particles sharded over ranks in equal contiguous slices (rank order = global order).

    poses_local (n, 8) / attrs_local (n, 9): float32 torch tensors (CUDA with nccl, CPU with gloo).  All-gathers the particle set
    (68 B per particle -- the exchange step), then calls `resample_fn(P_all, A_all, first, n_local) -> (P_new, A_new)` for this rank's
    champions.  The draws are keyed by the global particle index, so the concatenation over ranks equals the unsharded result."""
    import torch
    ws, rank = dist.get_world_size(), dist.get_rank()
    n_local = poses_local.shape[0]
    counts = torch.zeros(ws, dtype=torch.int64, device=poses_local.device)
    counts[rank] = n_local
    dist.all_reduce(counts)
    counts = [int(c) for c in counts.tolist()]
    if len(set(counts)) != 1:
        raise ValueError(f"gladiator_resample_sharded needs equal shard sizes, got {counts}")
    if sync is not None:
        sync()                      # whatever produced the local particles (on the handle's stream) must be complete before the collective reads them
    P_all = torch.empty((ws * n_local, 8), dtype=torch.float32, device=poses_local.device)
    A_all = torch.empty((ws * n_local, 9), dtype=torch.float32, device=poses_local.device)
    dist.all_gather_into_tensor(P_all, poses_local.contiguous())
    dist.all_gather_into_tensor(A_all, attrs_local.contiguous())
    if sync is not None:
        sync()                      # the collective ran on torch's stream; the resampling kernel runs on the handle's
    return resample_fn(P_all, A_all, rank * n_local, n_local)
