"""Multi-GPU plumbing for the MSM: contiguous sharding + a one-point-per-rank gather, all on the devices.

sum_i k_i P_i splits over any partition of the index set, so the path shards with NO data-path
collective: rank g owns the contiguous slice [g*n/G, (g+1)*n/G) of both vectors, runs the whole
single-GPU pipeline on it and produces one partial point (XYZZ: 4 coordinates, <= 192 bytes) that STAYS in
its GPU's memory (zkb200_msm_ex, out_loc = DEVICE).  The only exchange is one all-gather of those G records
(torch.distributed: NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests) into a persistent device buffer,
after which rank 0 adds them with one kernel launch (zkb200_sum_points_ex reading the device buffer) and
returns the canonical affine point -- the only bytes that cross PCIe are the 64/96-byte answer.
A batch of MSMs over one shared point array (KZG commitments over one SRS) is dealt out MSM-wise instead
(`batch_range`) and the per-rank result records are all-gathered the same way.
SURVEY.md section 8e; the reference itself has no multi-device path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(n_global: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice of rank `rank` out of `world`; slices differ by at most one element."""
    return n_global * rank // world, n_global * (rank + 1) // world


batch_range = shard_range   # whole MSMs of a batch are dealt to the ranks the same way


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def all_gather_partials(partial: np.ndarray, device: str = "cuda") -> np.ndarray:
    """Gather one fixed-size uint64 record per rank -> (world, words) uint64 on every rank (host arrays in and
    out; the CPU/gloo tests and small control-plane exchanges use this, the data path uses `Combiner`)."""
    import torch
    import torch.distributed as dist

    world, _ = _world()
    if world == 1:
        return np.ascontiguousarray(partial, dtype=np.uint64).reshape(1, -1)
    mine = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.int64).copy()).to(device)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine)
    return torch.stack(bufs).cpu().numpy().view(np.uint64)


class Combiner:
    """Persistent device buffers for the combine step of one (curve, record shape): this rank's record(s) and the
    all-gather output.  Nothing is allocated per step."""

    def __init__(self, curve: str, records_per_rank: int = 1, coords: int = 4):
        import torch

        import zikkurat_algebra_b200 as zk
        self.curve = curve
        self.world, self.rank = _world()
        self.words = records_per_rank * coords * zk.CURVES[curve]["nlimbs_p"]
        self.records_per_rank = records_per_rank
        self.mine = torch.zeros(self.words, dtype=torch.int64, device="cuda")
        self.all = torch.zeros(self.world * self.words, dtype=torch.int64, device="cuda") if self.world > 1 else self.mine

    def gather(self):
        """all-gather `mine` into `all` (every rank); the caller's stream is synchronised so that the library's own
        stream may read the buffer afterwards."""
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.all_gather_into_tensor(self.all, self.mine)
            torch.cuda.current_stream().synchronize()
        return self.all


_combiners = {}


def _combiner(curve: str, records_per_rank: int, coords: int) -> Combiner:
    key = (curve, records_per_rank, coords, _world())
    if key not in _combiners:
        _combiners[key] = Combiner(curve, records_per_rank, coords)
    return _combiners[key]


def msm_sharded(curve: str, scalars_shard, points_shard, npoints: Optional[int] = None, mont: bool = True,
                resident: bool = False, window: int = 0) -> Optional[np.ndarray]:
    """This rank's slice -> partial MSM on this rank's GPU -> device all-gather -> affine result on rank 0 (None
    elsewhere).  `resident`: the shard arguments are raw device pointers (inputs already in HBM), else host arrays."""
    import zikkurat_algebra_b200 as zk

    world, rank = _world()
    if world == 1:
        if resident:
            return zk.msm_device(curve, scalars_shard, points_shard, npoints, mont=mont, out="affine", window=window)[0]
        return zk.msm(curve, scalars_shard, points_shard, mont=mont, out="affine", window=window)
    cb = _combiner(curve, 1, 4)
    if npoints is None:
        npoints = int(np.asarray(scalars_shard).size // 4)
    zk.msm_to_device(curve, scalars_shard, points_shard, npoints, cb.mine.data_ptr(), mont=mont, out="xyzz", window=window,
                     resident=resident)
    allp = cb.gather()
    if rank == 0:
        return zk.sum_points_device(curve, allp.data_ptr(), world, in_repr="xyzz", out="affine")
    return None


def msm_batch_dealt(curve: str, scalars_mine, points, npoints: int, nmsm_global: int, mont: bool = True,
                    resident: bool = False, window: int = 0) -> Optional[np.ndarray]:
    """A batch of `nmsm_global` MSMs over ONE shared point array, whole MSMs dealt to the ranks (`batch_range`): this
    rank computes its MSMs (scalars_mine: its (nmsm_mine, n, 4) block; points: the shared array, replicated on every
    GPU), the canonical affine records are all-gathered on the devices, rank 0 returns (nmsm_global, 2L) uint64.
    Requires nmsm_global to be a multiple of the world size (equal record counts per rank)."""
    import zikkurat_algebra_b200 as zk

    world, rank = _world()
    lo, hi = batch_range(nmsm_global, world, rank)
    mine = hi - lo
    if world == 1:
        if resident:
            return zk.msm_device(curve, scalars_mine, points, npoints, nmsm=mine, mont=mont, out="affine", window=window)
        return zk.msm_batch(curve, scalars_mine, points, mont=mont, out="affine", window=window)
    if nmsm_global % world:
        raise ValueError("msm_batch_dealt: the batch size must be a multiple of the number of ranks")
    cb = _combiner(curve, mine, 2)
    zk.msm_to_device(curve, scalars_mine, points, npoints, cb.mine.data_ptr(), nmsm=mine, mont=mont, out="affine", window=window,
                     resident=resident)
    allp = cb.gather()
    if rank == 0:
        return allp.cpu().numpy().view(np.uint64).reshape(nmsm_global, -1)
    return None
