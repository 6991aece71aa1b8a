"""BASELINE.json's configs at their FULL sizes on the GPU, against the golden bytes of the unmodified reference C
(tests/golden/big_golden.json, generated once in the build container by tests/golden/make_big_golden.py), plus a direct
run of the threaded reference at 2^20, the resident point-array cache, the device-side combine and the in-library
multi-GPU path.  Everything goes through the C ABI.
"""
import os

import numpy as np
import pytest

from tests import refs, workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zk():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import zikkurat_algebra_b200 as z
    from zikkurat_algebra_b200 import build
    build.build()
    z.lib()
    return z


def _device_workload(zk, curve, n, seed, start=0):
    """(scalars host array, points device tensor) of the global workload slice [start, start + n)."""
    import torch
    L = zk.CURVES[curve]["nlimbs_p"]
    p0, d = refs.chain_base(curve)
    d_pts = torch.empty((n, 2 * L), dtype=torch.int64, device="cuda")
    zk.gen_chain(curve, n, p0, d, start=start, device_ptr=d_pts.data_ptr())
    return refs.counter_scalars(seed, start, n), d_pts


def test_device_chain_equals_oracle_chain_including_last_block(zk):
    """The GPU point generator of bench.py / these tests must produce the bytes the golden generator used:
    first block, an interior block and the LAST block of the biggest workloads."""
    for curve, logn in (("bn128", 24), ("bls12_381", 26)):
        n = 1 << logn
        for start in (0, 12345, n // 2 - 7, n - 4096):
            cnt = min(4096, n - start)
            p0, d = refs.chain_base(curve)
            got = zk.gen_chain(curve, cnt, p0, d, start=start)
            want = refs.chain_points(curve, cnt, start=start)
            assert got.tobytes() == want.tobytes(), (curve, start)


@pytest.mark.parametrize("name", ["bls20", "bn20"])
def test_config_2_pow_20_direct_reference_and_golden(zk, name):
    """BASELINE configs[1] (BLS12-381 2^20) and the BN254 twin: the reference C itself, run here on all host threads
    (1-7 s), the committed golden and the CUDA path through the reference-named symbol -- all three byte-equal."""
    c = workloads.CONFIGS[name]
    curve, n, seed = c["curve"], 1 << c["logn"], c["seed"]
    pts = refs.chain_points(curve, n, nthreads=os.cpu_count() or 4)
    sc = refs.counter_scalars(seed, 0, n)
    want = refs.ref_msm_threads(curve, sc, pts, mont=True, nthreads=os.cpu_count() or 4)
    golden = workloads.golden_bytes(curve, n, c["form"], seed)
    assert golden is not None and want.tobytes() == golden
    for rep in ("proj", "jac"):
        got = zk.call_reference_symbol(f"{curve}_G1_{rep}_MSM_mont_coeff_affine_out", sc, pts)
        assert got.tobytes() == golden, rep


@pytest.mark.parametrize("name,g", [("bls20", 2), ("bls20", 8), ("bn24", 1), ("bls26", 1)])
def test_full_size_configs_match_reference_golden(zk, name, g):
    """Global workloads of the multi-GPU configs computed on ONE GPU (device-resident inputs): BLS12-381 2^21 / 2^23
    (the weak-scaling sizes of 2 / 8 GPUs), BN254 2^24 (config 3), BLS12-381 2^26 with std scalars (config 4)."""
    import torch
    c = workloads.CONFIGS[name]
    curve, seed, form = c["curve"], c["seed"], c["form"]
    n = (g << c["logn"]) if c["weak"] else (1 << c["logn"])
    golden = workloads.golden_bytes(curve, n, form, seed)
    assert golden is not None, "tests/golden/big_golden.json lacks this config: run tests/golden/make_big_golden.py"
    sc, d_pts = _device_workload(zk, curve, n, seed)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    got = zk.msm_device(curve, d_sc.data_ptr(), d_pts.data_ptr(), n, mont=(form == "mont"), out="affine")[0]
    assert got.tobytes() == golden
    # the same workload as 8 contiguous shards whose XYZZ partials never leave the device until the final sum
    L = zk.CURVES[curve]["nlimbs_p"]
    parts = torch.zeros((8, 4 * L), dtype=torch.int64, device="cuda")
    for k in range(8):
        lo, hi = n * k // 8, n * (k + 1) // 8
        zk.msm_to_device(curve, d_sc[lo:hi].data_ptr(), d_pts[lo:hi].data_ptr(), hi - lo, parts[k].data_ptr(), mont=(form == "mont"),
                         out="xyzz")
    assert zk.sum_points_device(curve, parts.data_ptr(), 8, "xyzz", "affine").tobytes() == golden
    del d_sc, d_pts
    torch.cuda.empty_cache()
    zk.release_workspaces()


def test_batched_kzg_config_matches_reference_golden(zk):
    """config 5: 256 independent BN254 MSMs of 2^14 points over one shared SRS -- one batched call, 256 single calls through
    the reference-named symbol (second call onwards served by the resident copy of the SRS), and the golden."""
    c = workloads.CONFIGS["kzg"]
    curve, n, seed, nmsm = c["curve"], 1 << c["logn"], c["seed"], c["nmsm"]
    golden = workloads.golden_bytes(curve, n, c["form"], seed, nmsm)
    assert golden is not None
    srs = refs.chain_points(curve, n)
    sc = workloads.batch_scalars(seed, nmsm, n)
    got = zk.msm_batch(curve, sc, srs, mont=True, out="affine")
    assert got.tobytes() == golden
    zk.srs_cache_drop()
    L = zk.CURVES[curve]["nlimbs_p"]
    hits = 0
    for m in range(0, nmsm, 17):
        one = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_mont_coeff_affine_out", sc[m], srs)
        assert one.tobytes() == golden[m * 16 * L:(m + 1) * 16 * L], m
        hits += zk.last_srs_hit()
    assert hits == len(range(0, nmsm, 17)) - 1      # every call but the first found the SRS on the device


@pytest.mark.parametrize("curve", ["bn128", "bls12_381"])
def test_resident_copy_cache_is_keyed_by_content(zk, curve):
    """The cache must never serve stale points: same host buffer, new content -> new answer (fingerprint mismatch);
    same content -> hit; a different buffer of the same content -> its own entry; disabled budget -> never a hit."""
    n = 1 << 13
    L = zk.CURVES[curve]["nlimbs_p"]
    sym = f"{curve}_G1_proj_MSM_mont_coeff_affine_out"
    lib, pre = (refs.ref(), "") if refs.have_ref() else (refs.oracle(), "zko_")
    cpu = lambda sc, pts: refs.call_msm(lib, pre + sym, sc.ravel(), pts.ravel(), 2 * L, n=n)
    a = refs.chain_points(curve, n, s0=0x51, s1=0x77)
    b = refs.chain_points(curve, n, s0=0x99, s1=0x31)
    sc = refs.counter_scalars(77, 0, n)
    buf = a.copy()
    zk.srs_cache_drop()
    r1 = zk.call_reference_symbol(sym, sc, buf); h1 = zk.last_srs_hit()
    r2 = zk.call_reference_symbol(sym, sc, buf); h2 = zk.last_srs_hit()
    assert (h1, h2) == (False, True)
    assert r1.tobytes() == r2.tobytes() == cpu(sc, a).tobytes()
    buf[:] = b                                     # same address, same size, other points
    r3 = zk.call_reference_symbol(sym, sc, buf); h3 = zk.last_srs_hit()
    assert h3 is False and r3.tobytes() == cpu(sc, b).tobytes()
    r4 = zk.call_reference_symbol(sym, sc, buf)
    assert zk.last_srs_hit() and r4.tobytes() == r3.tobytes()
    # a point changed at an unsampled position of a LARGE array is the documented limit of the fingerprint; edge words and
    # the strided sample are covered: the first and the last record always are
    buf[0] = a[0]
    want = cpu(sc, buf)
    assert zk.call_reference_symbol(sym, sc, buf).tobytes() == want.tobytes() and not zk.last_srs_hit()
    buf[n - 1] = a[n - 1]
    want = cpu(sc, buf)
    assert zk.call_reference_symbol(sym, sc, buf).tobytes() == want.tobytes() and not zk.last_srs_hit()
    # batch entry and other output representations go through the same cache
    sc2 = np.stack([refs.counter_scalars(80 + i, 0, n) for i in range(3)])
    got = zk.msm_batch(curve, sc2, buf, mont=True, out="affine")
    assert zk.last_srs_hit()
    for i in range(3):
        assert got[i].tobytes() == cpu(sc2[i], buf).tobytes()
    zk.release_workspaces()
    assert zk.call_reference_symbol(sym, sc, buf).tobytes() == want.tobytes() and not zk.last_srs_hit()


def test_in_library_multi_gpu_matches_golden(zk):
    """ZKB200_DEVICES / zkb200_set_devices: ONE process hands the whole host arrays to the reference-named symbol and the
    library shards them over the GPUs of the box.  BN254 2^22 prefix checks + the BLS12-381 2^21 golden.  Needs >= 2 GPUs."""
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs 2 GPUs")
    devs = list(range(min(ng, 8)))
    c = workloads.CONFIGS["bls20"]
    curve, seed = c["curve"], c["seed"]
    n = 2 << c["logn"]
    golden = workloads.golden_bytes(curve, n, c["form"], seed)
    pts = refs.chain_points(curve, n, nthreads=os.cpu_count() or 4)
    sc = refs.counter_scalars(seed, 0, n)
    try:
        zk.set_devices(devs)
        for _ in range(2):       # second call: every device serves its slice of the points from its resident copy
            got = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_mont_coeff_affine_out", sc, pts)
            assert got.tobytes() == golden
        k = workloads.CONFIGS["kzg"]
        srs = refs.chain_points(k["curve"], 1 << k["logn"])
        bsc = workloads.batch_scalars(k["seed"], k["nmsm"], 1 << k["logn"])
        got = zk.msm_batch(k["curve"], bsc, srs, mont=True, out="affine")
        assert got.tobytes() == workloads.golden_bytes(k["curve"], 1 << k["logn"], k["form"], k["seed"], k["nmsm"])
    finally:
        zk.set_devices([])


def test_glv_precondition_and_switch(zk):
    """The endomorphism split assumes points of the prime-order subgroup.  BN254 G1 has cofactor 1 (nothing to assume);
    BLS12-381 G1 has curve points outside the subgroup, which the reference multiplies like any other point
    (it validates nothing: lib/cbits/curves/g1/proj/bn128_G1_proj.c:549-561).  With the split OFF the library reproduces
    the reference for such inputs bit for bit; with it ON it reproduces the reference for subgroup points (every other test)."""
    from tests import pyec
    curve = "bls12_381"
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    # curve points with small x: on y^2 = x^3 + 4 but (with overwhelming probability) not in the subgroup of order r
    pts = []
    x = 1
    while len(pts) < 64:
        x += 1
        rhs = (x * x * x + 4) % cv.p
        y = pow(rhs, (cv.p + 1) // 4, cv.p)          # p = 3 mod 4
        if y * y % cv.p == rhs:
            pts.append((x, y))
    assert any(cv.mul(cv.r, P) is not None for P in pts[:4])          # really outside the subgroup
    arr = np.frombuffer(cv.points_to_bytes(pts), dtype=np.uint64).copy().reshape(len(pts), 2 * L)
    sc = refs.counter_scalars(91, 0, len(pts))
    lib, pre = (refs.ref(), "") if refs.have_ref() else (refs.oracle(), "zko_")
    sym = f"{curve}_G1_proj_MSM_std_coeff_affine_out"
    want = refs.call_msm(lib, pre + sym, sc.ravel(), arr.ravel(), 2 * L, n=len(pts))
    try:
        zk.set_glv(False)
        got_off = zk.call_reference_symbol(sym, sc, arr)
    finally:
        zk.set_glv(True)
    assert got_off.tobytes() == want.tobytes()
    # subgroup points: identical with and without the split
    sub = refs.chain_points(curve, 64, s0=5, s1=9)
    want = refs.call_msm(lib, pre + sym, sc.ravel(), sub.ravel(), 2 * L, n=64)
    on = zk.call_reference_symbol(sym, sc, sub)
    try:
        zk.set_glv(False)
        off = zk.call_reference_symbol(sym, sc, sub)
    finally:
        zk.set_glv(True)
    assert on.tobytes() == off.tobytes() == want.tobytes()
