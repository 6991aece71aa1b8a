"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous sharding + gather of one partial point per
rank + final sum == unsharded answer.  No GPU here, so the partials come from the CPU oracle; the plumbing
(zikkurat_algebra_b200.distributed) is the code under test."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, curve, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from tests import refs
    from zikkurat_algebra_b200.distributed import all_gather_partials, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = refs.CURVE_LIMBS[curve]
    pts = refs.chain_points(curve, n)
    sc = refs.random_scalars(curve, n, seed=31)
    lo, hi = shard_range(n, world, rank)
    part = refs.call_msm(refs.oracle(), f"zko_{curve}_G1_proj_MSM_mont_coeff_proj_out", sc[lo:hi].ravel(), pts[lo:hi].ravel(), 3 * L, n=hi - lo)
    allp = all_gather_partials(part, device="cpu")
    assert allp.shape == (world, 3 * L)
    assert allp[rank].tobytes() == part.tobytes()
    if rank == 0:
        acc = allp[0].copy()
        for k in range(1, world):
            acc = refs.call3(refs.oracle(), f"zko_{curve}_G1_proj_add", acc, allp[k].copy(), 3 * L)
        got = refs.call2(refs.oracle(), f"zko_{curve}_G1_proj_to_affine", acc, 2 * L)
        want = refs.call_msm(refs.oracle(), f"zko_{curve}_G1_proj_MSM_mont_coeff_affine_out", sc.ravel(), pts.ravel(), 2 * L, n=n)
        q.put(got.tobytes() == want.tobytes())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("curve,n", [("bn128", 301), ("bls12_381", 200)])
def test_two_rank_gather_equals_unsharded(curve, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + (0 if curve == "bn128" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, curve, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_ranges_partition():
    from zikkurat_algebra_b200.distributed import shard_range
    for n in (0, 1, 7, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 4, 8):
            edges = [shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
