"""CPU checks of the reproducible workloads (tests/workloads.py) and of the committed full-size goldens: the generators
are pure functions of the global index (so every rank of a multi-GPU run regenerates its slice), every BASELINE
config has its golden, and the cheap goldens are recomputed here with the reference C."""
import os

import numpy as np

from tests import pyec, refs, workloads


def _splitmix64(x):
    M = (1 << 64) - 1
    z = (x * 0x9E3779B97F4A7C15 + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


def test_counter_scalars_are_a_function_of_the_global_index():
    a = refs.counter_scalars(9, 0, 5000)
    assert a.shape == (5000, 4) and a.dtype == np.uint64
    for start, cnt in ((0, 10), (123, 77), (4990, 10)):
        assert np.array_equal(refs.counter_scalars(9, start, cnt), a[start:start + cnt])
    for i, j in ((0, 0), (7, 2), (4999, 1)):
        assert int(a[i, j]) == _splitmix64((9 << 32) + 4 * i + j)
    assert int(a[11, 3]) == _splitmix64((9 << 32) + 47) & ((1 << 61) - 1)
    assert int(a[:, 3].max()) < 1 << 61          # < 2^253 < r: valid as a standard integer and as a Montgomery word
    assert not np.array_equal(refs.counter_scalars(10, 0, 16), a[:16])


def test_chain_points_slices_and_threads_agree():
    for curve in ("bn128", "bls12_381"):
        cv = pyec.CURVES[curve]
        whole = refs.chain_points(curve, 9000)
        assert np.array_equal(refs.chain_points(curve, 9000, nthreads=4), whole)
        assert np.array_equal(refs.chain_points(curve, 100, start=8900), whole[8900:])
        for i in (0, 1, 8999):      # against the independent Python model
            assert whole[i].tobytes() == cv.affine_to_bytes(cv.mul(0x1234567 + i * 0x7654321, cv.gen))


def test_every_config_has_its_golden():
    vec = workloads.load_big_golden()
    for name, c in workloads.CONFIGS.items():
        sizes = [g << c["logn"] for g in (1, 2, 4, 8)] if c["weak"] else [1 << c["logn"]]
        L = refs.CURVE_LIMBS[c["curve"]]
        for n in sizes:
            blob = workloads.golden_bytes(c["curve"], n, c["form"], c["seed"], c["nmsm"])
            assert blob is not None, (name, n)
            assert len(blob) == c["nmsm"] * 16 * L
            assert blob != b"\xff" * len(blob)
    assert len(vec) >= 11


def test_cheap_goldens_recomputed_with_the_reference():
    """BN254 2^20 (3-4 s on 8 threads) and the first MSMs of the batched-KZG config, recomputed here."""
    if not refs.have_ref():
        import pytest
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    T = os.cpu_count() or 4
    c = workloads.CONFIGS["bn20"]
    n = 1 << c["logn"]
    got = refs.ref_msm_threads(c["curve"], refs.counter_scalars(c["seed"], 0, n), refs.chain_points(c["curve"], n, nthreads=T), mont=True, nthreads=T)
    assert got.tobytes() == workloads.golden_bytes(c["curve"], n, c["form"], c["seed"])
    k = workloads.CONFIGS["kzg"]
    n = 1 << k["logn"]
    srs = refs.chain_points(k["curve"], n)
    blob = workloads.golden_bytes(k["curve"], n, k["form"], k["seed"], k["nmsm"])
    for m in (0, 255):
        sc = refs.counter_scalars(k["seed"] + m, 0, n)
        one = refs.call_msm(refs.ref(), "bn128_G1_proj_MSM_mont_coeff_affine_out", sc.ravel(), srs.ravel(), 8, n=n)
        assert one.tobytes() == blob[m * 64:(m + 1) * 64]
    # the restatement (oracle/msm_oracle.c) agrees on the same batch entries
    one = refs.call_msm(refs.oracle(), "zko_bn128_G1_proj_MSM_mont_coeff_affine_out", sc.ravel(), srs.ravel(), 8, n=n)
    assert one.tobytes() == blob[255 * 64:256 * 64]


def test_batch_ranges_deal_whole_msms():
    from zikkurat_algebra_b200.distributed import batch_range
    for world in (1, 2, 4, 8):
        got = [batch_range(256, world, r) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == 256 and all(b - a == 256 // world for a, b in got)
