"""Multi-GPU plumbing for the MSM: contiguous sharding + a one-point-per-rank gather.

sum_i k_i P_i splits over any partition of the index set, so the path shards with NO data-path
collective: rank g owns the contiguous slice [g*n/G, (g+1)*n/G) of both vectors, runs the whole
single-GPU pipeline on it and produces one partial point (4 coordinates, <= 192 bytes).  The only
exchange is an all-gather of those G points (torch.distributed: NCCL over NVLink on GPUs, gloo in the
CPU tests), after which rank 0 adds them (zkb200_sum_points) and converts to affine.
SURVEY.md section 8e; the reference itself has no multi-device path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(n_global: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice of rank `rank` out of `world`; slices differ by at most one element."""
    return n_global * rank // world, n_global * (rank + 1) // world


def all_gather_partials(partial: np.ndarray, device: str = "cuda") -> np.ndarray:
    """Gather one fixed-size uint64 record per rank -> (world, words) uint64 on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.ascontiguousarray(partial, dtype=np.uint64).reshape(1, -1)
    mine = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.int64).copy()).to(device)
    bufs = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(bufs, mine)
    return torch.stack(bufs).cpu().numpy().view(np.uint64)


def msm_sharded(curve: str, scalars_shard, points_shard, npoints: Optional[int] = None, mont: bool = True,
                resident: bool = False, window: int = 0) -> Optional[np.ndarray]:
    """This rank's slice -> partial MSM on this rank's GPU -> gather -> affine result on rank 0 (None elsewhere).
    `resident`: the shard arguments are raw device pointers (inputs already in HBM), else host arrays."""
    import torch.distributed as dist

    import zikkurat_algebra_b200 as zk

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mode = "affine" if world == 1 else "xyzz"
    if resident:
        part = zk.msm_device(curve, scalars_shard, points_shard, npoints, mont=mont, out=mode, window=window)[0]
    else:
        part = zk.msm(curve, scalars_shard, points_shard, mont=mont, out=mode, window=window)
    if world == 1:
        return part
    allp = all_gather_partials(part, device="cuda")
    if rank == 0:
        return zk.sum_points(curve, allp, in_repr="xyzz", out="affine")
    return None
