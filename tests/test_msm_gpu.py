"""Parity tests proper (B200): the CUDA path, called through the C-ABI, against the CPU oracles.

Bar: canonical affine output BIT-EXACT with the reference C (oracle/_ref, falling back to the pinned
restatement oracle/libzk_oracle.so); proj/jac outputs equal as group elements after the reference's own
*_to_affine.  Edge cases are the reference's: empty / tiny inputs, infinity inputs,
P and -P, repeated points, un-reduced 256-bit scalars, zero scalars (section 8b, 8d).
"""
import ctypes
import json
import os
import random

import numpy as np
import pytest

from tests import pyec, refs

pytestmark = pytest.mark.gpu

CURVES = ["bn128", "bls12_381"]
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "msm_golden.json")))["vectors"]


@pytest.fixture(scope="module")
def zk():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import zikkurat_algebra_b200 as z
    from zikkurat_algebra_b200 import build
    build.build()
    z.lib()
    return z


def cpu_lib():
    return (refs.ref(), "") if refs.have_ref() else (refs.oracle(), "zko_")


def cpu_affine(curve, sc, pts, form="std", rep="proj", n=None):
    lib, pre = cpu_lib()
    L = refs.CURVE_LIMBS[curve]
    n = sc.size // 4 if n is None else n
    if n == 0:  # the shipped reference evaluates log2(0); infinity is the defined answer
        return np.full(2 * L, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    return refs.call_msm(lib, f"{pre}{curve}_G1_{rep}_MSM_{form}_coeff_affine_out", sc.ravel(), pts.ravel(), 2 * L, n=n)


def to_affine_cpu(curve, rep, buf):
    lib, pre = cpu_lib()
    return refs.call2(lib, f"{pre}{curve}_G1_{rep}_to_affine", buf, 2 * refs.CURVE_LIMBS[curve])


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 123, 1000, 5000])
def test_all_reference_symbols(zk, curve, n):
    L = refs.CURVE_LIMBS[curve]
    pts = refs.chain_points(curve, n)
    for form in ("std", "mont"):
        sc = refs.random_scalars(curve, n, seed=n + 17, reduce=(form == "mont"))
        want = cpu_affine(curve, sc, pts, form, n=n)
        for rep in ("proj", "jac"):
            got = zk.call_reference_symbol(f"{curve}_G1_{rep}_MSM_{form}_coeff_affine_out", sc, pts, npoints=n)
            assert got.tobytes() == want.tobytes(), (curve, n, form, rep, "affine_out")
            out = zk.call_reference_symbol(f"{curve}_G1_{rep}_MSM_{form}_coeff_{rep}_out", sc, pts, npoints=n)
            assert to_affine_cpu(curve, rep, out).tobytes() == want.tobytes(), (curve, n, form, rep, "rep_out")


@pytest.mark.parametrize("curve", CURVES)
def test_infinity_representations(zk, curve):
    """n = 0 and all-zero scalars: proj (0,R,0), jac (R,R,0), affine 0xFF."""
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    one = np.frombuffer((cv.R % cv.p).to_bytes(cv.fp_bytes, "little"), dtype=np.uint64)
    zero = np.zeros(L, np.uint64)
    pts = refs.chain_points(curve, 4)
    for n, sc in ((0, np.zeros((0, 4), np.uint64)), (4, np.zeros((4, 4), np.uint64))):
        p = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_proj_out", sc, pts[:n], npoints=n)
        assert p.tobytes() == np.concatenate([zero, one, zero]).tobytes()
        j = zk.call_reference_symbol(f"{curve}_G1_jac_MSM_std_coeff_jac_out", sc, pts[:n], npoints=n)
        assert j.tobytes() == np.concatenate([one, one, zero]).tobytes()
        a = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_affine_out", sc, pts[:n], npoints=n)
        assert a.tobytes() == b"\xff" * (16 * L)


@pytest.mark.parametrize("curve", CURVES)
def test_adversarial_inputs(zk, curve):
    cv = pyec.CURVES[curve]
    G = cv.gen
    P = cv.mul(7, G)
    Q = cv.mul(1234567, G)
    top = (1 << 256) - 1
    sets = {
        "p_minus_p": ([5, 5], [P, cv.neg(P)]),
        "p_minus_p_then_more": ([5, 5, 9], [P, cv.neg(P), Q]),
        "repeated": ([3, 3, 3, 3, 3], [P, P, P, P, P]),
        "repeated_mixed": ([3, 3, 3, 8, 8], [P, P, Q, Q, Q]),
        "inf_input": ([9, 2, 1, 4], [None, P, None, Q]),
        "all_inf": ([9, 2], [None, None]),
        "full256": ([top, (1 << 255) + 12345, top - 1], [G, P, Q]),
        "topbit_lowbit": ([1 << 255, 1, 1 << 254, 2], [G, P, Q, G]),
        "small_scalars": (list(range(40)), pyec.chain_points(cv, 40, 3, 5)),
        "same_scalar_many_points": ([0xABCDEF] * 64, pyec.chain_points(cv, 64, 9, 11)),
        "neg_pairs_same_bucket": ([77] * 6, [P, Q, cv.neg(P), cv.neg(Q), P, cv.neg(P)]),
    }
    for name, (ks, pts) in sets.items():
        want = cv.affine_to_bytes(cv.msm(ks, pts))
        Pb = np.frombuffer(cv.points_to_bytes(pts), dtype=np.uint64).copy()
        S = np.frombuffer(b"".join(cv.scalar_std_bytes(k) for k in ks), dtype=np.uint64).copy()
        got = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_affine_out", S, Pb)
        assert got.tobytes() == want, name
        assert cpu_affine(curve, S, Pb).tobytes() == want, name  # and the oracle agrees with Python
        if all(k < cv.r for k in ks):
            Sm = np.frombuffer(b"".join(cv.scalar_mont_bytes(k) for k in ks), dtype=np.uint64).copy()
            got = zk.call_reference_symbol(f"{curve}_G1_jac_MSM_mont_coeff_affine_out", Sm, Pb)
            assert got.tobytes() == want, name + " (mont)"


@pytest.mark.parametrize("curve", CURVES)
def test_window_widths_and_long_runs(zk, curve):
    """every window width through the *_variable symbol; few distinct scalars => runs far longer than a chunk."""
    L = refs.CURVE_LIMBS[curve]
    n = 3000
    pts = refs.chain_points(curve, n)
    sc = refs.random_scalars(curve, n, seed=5, reduce=False)
    want = cpu_affine(curve, sc, pts)
    for c in (1, 2, 3, 5, 8, 9, 12, 16, 17):
        out = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_proj_out_variable", sc, pts, window_size=c)
        assert to_affine_cpu(curve, "proj", out).tobytes() == want.tobytes(), c
    few = sc[:3]
    sc2 = np.ascontiguousarray(few[np.arange(n) % 3])
    want2 = cpu_affine(curve, sc2, pts)
    for c in (4, 11):
        out = zk.call_reference_symbol(f"{curve}_G1_jac_MSM_std_coeff_jac_out_variable", sc2, pts, window_size=c)
        assert to_affine_cpu(curve, "jac", out).tobytes() == want2.tobytes(), c


@pytest.mark.parametrize("curve", CURVES)
def test_short_scalars_expo_nlimbs(zk, curve):
    L = refs.CURVE_LIMBS[curve]
    n = 200
    pts = refs.chain_points(curve, n)
    lib, pre = cpu_lib()
    for nl in (1, 2, 3):
        sc = refs.random_scalars(curve, n, seed=nl, reduce=False)[:, :nl].copy()
        want = refs.call_msm(lib, f"{pre}{curve}_G1_proj_MSM_std_coeff_affine_out", sc.ravel(), pts.ravel(), 2 * L, n=n, nlimbs=nl)
        got = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_affine_out", sc, pts, npoints=n, expo_nlimbs=nl)
        assert got.tobytes() == want.tobytes(), nl


@pytest.mark.parametrize("v", GOLDEN, ids=lambda g: f"{g['curve']}-{g['n']}-{g['form']}")
def test_golden_vectors(zk, v):
    """committed fixtures generated from the unmodified reference C (tests/golden/make_golden.py);
    includes BASELINE configs[0]: BN254 2^16."""
    curve, n, form = v["curve"], v["n"], v["form"]
    pts = refs.chain_points(curve, n)
    sc = refs.random_scalars(curve, n, seed=v["seed"], reduce=(form == "mont"))
    got = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_{form}_coeff_affine_out", sc, pts, npoints=n)
    assert got.tobytes().hex() == v["affine_hex"]


@pytest.mark.parametrize("curve", CURVES)
def test_batch_and_device_entry(zk, curve):
    import torch
    L = refs.CURVE_LIMBS[curve]
    n, nmsm = 1 << 12, 5
    pts = refs.chain_points(curve, n)
    sc = np.stack([refs.random_scalars(curve, n, seed=50 + i) for i in range(nmsm)])
    single = np.stack([zk.msm(curve, sc[i], pts, mont=True, out="affine") for i in range(nmsm)])
    batch = zk.msm_batch(curve, sc, pts, mont=True, out="affine")
    assert batch.tobytes() == single.tobytes()
    for i in range(nmsm):
        assert single[i].tobytes() == cpu_affine(curve, sc[i], pts, "mont").tobytes()
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_pts = torch.from_numpy(pts.view(np.int64)).cuda()
    torch.cuda.synchronize()
    dev = zk.msm_device(curve, d_sc.data_ptr(), d_pts.data_ptr(), n, nmsm=nmsm, mont=True, out="affine")
    assert dev.tobytes() == single.tobytes()
    st = zk.last_stats()
    # pairs per window: n, or 2n with the GLV split (P_i and phi(P_i))
    assert st["insertions"] in (nmsm * n * st["nwindows"], 2 * nmsm * n * st["nwindows"]) and st["phase_ms"]["accumulate"] > 0


@pytest.mark.parametrize("curve", CURVES)
def test_sharded_partials_combine(zk, curve):
    """contiguous shards -> partial proj/xyzz results -> zkb200_sum_points == unsharded answer (multi-GPU combine, K8)."""
    n = 6000
    pts = refs.chain_points(curve, n)
    sc = refs.random_scalars(curve, n, seed=77)
    want = cpu_affine(curve, sc, pts, "mont")
    for rep in ("proj", "jac", "xyzz"):
        parts = []
        for k in range(4):
            lo, hi = n * k // 4, n * (k + 1) // 4
            parts.append(zk.msm(curve, sc[lo:hi], pts[lo:hi], mont=True, out=rep))
        got = zk.sum_points(curve, np.stack(parts), in_repr=rep, out="affine")
        assert got.tobytes() == want.tobytes(), rep
    # a partial that is infinity, and opposite partials
    cv = pyec.CURVES[curve]
    inf = zk.msm(curve, np.zeros((2, 4), np.uint64), pts[:2], mont=False, out="proj")
    assert zk.sum_points(curve, np.stack([inf, inf]), "proj", "affine").tobytes() == b"\xff" * cv.affine_bytes


@pytest.mark.parametrize("curve,logn", [("bn128", 18), ("bls12_381", 17)])
def test_mid_size_vs_threaded_reference(zk, curve, logn):
    n = 1 << logn
    pts = refs.chain_points(curve, n)
    sc = refs.random_scalars(curve, n, seed=logn)
    want = refs.ref_msm_threads(curve, sc, pts, mont=True, nthreads=os.cpu_count() or 4)
    got = zk.msm(curve, sc, pts, mont=True, out="affine")
    assert got.tobytes() == want.tobytes()


def test_full_size_properties_bls12_381_2_20(zk):
    """BASELINE configs[1] size (BLS12-381, 2^20): size-independent properties instead of a CPU run.
    (1) split: MSM(all) == MSM(first half) + MSM(second half);  (2) linearity in the scalars:
    MSM(k) + MSM(k') == MSM(k + k' mod 2^256 as integers when no overflow)."""
    curve = "bls12_381"
    n = 1 << 20
    pts = refs.chain_points(curve, n)
    sc = refs.random_scalars(curve, n, seed=2020)          # < 2^253
    whole = zk.msm(curve, sc, pts, mont=False, out="affine")
    h = n // 2
    a = zk.msm(curve, sc[:h], pts[:h], mont=False, out="proj")
    b = zk.msm(curve, sc[h:], pts[h:], mont=False, out="proj")
    assert zk.sum_points(curve, np.stack([a, b]), "proj", "affine").tobytes() == whole.tobytes()
    sc2 = refs.random_scalars(curve, n, seed=2021)
    # limb-wise big-int addition with carries (values < 2^253, so the sum fits in 256 bits)
    s = np.zeros_like(sc)
    carry = np.zeros(n, dtype=np.uint64)
    for j in range(4):
        t = sc[:, j] + sc2[:, j]
        c1 = (t < sc[:, j]).astype(np.uint64)
        t2 = t + carry
        c2 = (t2 < t).astype(np.uint64)
        s[:, j] = t2
        carry = c1 + c2
    assert not carry.any()
    x = zk.msm(curve, sc, pts, mont=False, out="jac")
    y = zk.msm(curve, sc2, pts, mont=False, out="jac")
    z = zk.msm(curve, s, pts, mont=False, out="affine")
    assert zk.sum_points(curve, np.stack([x, y]), "jac", "affine").tobytes() == z.tobytes()
    # and a 2^16 prefix of the same inputs against the CPU reference
    m = 1 << 16
    assert zk.msm(curve, sc[:m], pts[:m], mont=False).tobytes() == cpu_affine(curve, sc[:m], pts[:m]).tobytes()


def test_imad_probe_runs(zk):
    v = zk.imad_peak(0, 200)
    assert v > 1e11


def test_in_library_multi_gpu_sharding(zk):
    """$ZKB200_DEVICES / zkb200_set_devices: the C-ABI call itself shards over the GPUs of the box (one host
    thread per device) -- same bytes as the single-device call.  Needs >= 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    devs = list(range(min(torch.cuda.device_count(), 8)))
    for curve, logn in (("bn128", 18), ("bls12_381", 17)):
        n = (1 << logn) + 37
        pts = refs.chain_points(curve, n)
        sc = refs.random_scalars(curve, n, seed=logn + 3)
        single = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_mont_coeff_affine_out", sc, pts)
        batch_sc = np.stack([refs.random_scalars(curve, 4096, seed=900 + i) for i in range(len(devs) + 1)])
        batch_single = zk.msm_batch(curve, batch_sc, pts[:4096], mont=True, out="affine")
        try:
            zk.set_devices(devs)
            multi = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_mont_coeff_affine_out", sc, pts)
            multi_proj = zk.call_reference_symbol(f"{curve}_G1_jac_MSM_mont_coeff_jac_out", sc, pts)
            batch_multi = zk.msm_batch(curve, batch_sc, pts[:4096], mont=True, out="affine")
        finally:
            zk.set_devices([])
        assert multi.tobytes() == single.tobytes()
        assert to_affine_cpu(curve, "jac", multi_proj).tobytes() == single.tobytes()
        assert batch_multi.tobytes() == batch_single.tobytes()


@pytest.mark.parametrize("curve", CURVES)
def test_batch_conversions_next_row(zk, curve):
    """<curve>_G1_{proj,jac}_batch_{to,from}_affine, bit-identical to the reference C
    (which does N separate inversions: bn128_G1_proj.c:158-166)."""
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    rng = random.Random(77)
    n = 203
    aff_pts = pyec.chain_points(cv, n, 11, 13)
    for k in (0, 5, 100, 202):
        aff_pts[k] = None
    lib, pre = cpu_lib()
    for rep in ("proj", "jac"):
        recs = []
        for P in aff_pts:
            if P is None:
                z = 0
                X, Y = (rng.randrange(1, cv.p), rng.randrange(1, cv.p))
            else:
                z = rng.randrange(1, cv.p)
                X = P[0] * (z if rep == "proj" else z * z) % cv.p
                Y = P[1] * (z if rep == "proj" else z * z * z) % cv.p
            recs.append(cv.fp_to_bytes(X) + cv.fp_to_bytes(Y) + cv.fp_to_bytes(z))
        src = np.frombuffer(b"".join(recs), dtype=np.uint64).copy().reshape(n, 3 * L)
        got = zk.batch_to_affine(curve, src, rep)
        assert got.tobytes() == cv.points_to_bytes(aff_pts)
        f = getattr(lib, f"{pre}{curve}_G1_{rep}_to_affine")
        for i in (0, 1, 5, 77, 202):
            assert refs.call2(lib, f"{pre}{curve}_G1_{rep}_to_affine", src[i].copy(), 2 * L).tobytes() == got[i].tobytes()
        # from_affine: exact records, infinity encodings included
        aff = np.frombuffer(cv.points_to_bytes(aff_pts), dtype=np.uint64).copy().reshape(n, 2 * L)
        back = zk.batch_from_affine(curve, aff, rep)
        if refs.have_ref():
            g = getattr(refs.ref(), f"{curve}_G1_{rep}_batch_from_affine")
            g.argtypes = [ctypes.c_int, refs.U64P, refs.U64P]
            g.restype = None
            want = np.zeros((n, 3 * L), np.uint64)
            g(n, refs.ptr(aff.ravel()), refs.ptr(want.ravel()))
            assert back.tobytes() == want.tobytes()
        assert zk.batch_to_affine(curve, back, rep).tobytes() == aff.tobytes()
    # a larger round trip through MSM-independent data: 2^16 chain points -> proj -> affine
    big = refs.chain_points(curve, 1 << 16)
    assert zk.batch_to_affine(curve, zk.batch_from_affine(curve, big, "jac"), "jac").tobytes() == big.tobytes()
    # large arrays take the one-inversion-per-call path (batch inversion tree): random denominators, some Z = 0, against
    # the reference's own per-point <curve>_G1_{proj,jac}_to_affine applied to every record
    from tests.test_device_primitives import cpu_map2
    n2 = 9001
    base = refs.chain_points(curve, n2, s0=0x4242, s1=0x99)
    for rep in ("proj", "jac"):
        recs = np.zeros((n2, 3 * L), np.uint64)
        for i in range(n2):
            x, y = cv.affine_from_bytes(base[i].tobytes())
            z = 0 if i % 1000 == 7 else rng.randrange(1, cv.p)
            X = x * (z if rep == "proj" else z * z) % cv.p if z else rng.randrange(1, cv.p)
            Y = y * (z if rep == "proj" else z * z * z) % cv.p if z else rng.randrange(1, cv.p)
            recs[i] = np.frombuffer(cv.fp_to_bytes(X) + cv.fp_to_bytes(Y) + cv.fp_to_bytes(z), dtype=np.uint64)
        want = cpu_map2(f"{curve}_G1_{rep}_to_affine", recs, 2 * L)
        assert zk.batch_to_affine(curve, recs, rep).tobytes() == want.tobytes(), rep


def test_concurrent_callers_are_serialised_correctly(zk):
    """The reference is re-entrant (threading); here concurrent host threads share one device
    context guarded by a mutex.  Four threads, different inputs, interleaved calls: every result must match."""
    import threading
    curve = "bn128"
    n = 1 << 13
    pts = refs.chain_points(curve, n)
    scs = [refs.random_scalars(curve, n, seed=300 + i) for i in range(4)]
    want = [cpu_affine(curve, sc, pts, "mont").tobytes() for sc in scs]
    got = [[None] * 3 for _ in range(4)]

    def work(i):
        for r in range(3):
            got[i][r] = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_mont_coeff_affine_out", scs[i], pts).tobytes()

    ths = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for i in range(4):
        assert all(g == want[i] for g in got[i])


@pytest.mark.parametrize("curve", CURVES)
def test_input_slices_give_identical_bytes(zk, curve):
    """$ZKB200_SLICES: host inputs cut into K slices that are transferred / sorted / accumulated one after the
    other into K bucket arrays (H2D overlap).  Any K must give the same bytes, including ragged n and n < K."""
    pts_all = refs.chain_points(curve, 5003)
    old = os.environ.get("ZKB200_SLICES")
    try:
        for n in (1, 2, 3, 7, 1000, 5003):
            pts, sc = pts_all[:n], refs.random_scalars(curve, n, seed=n, reduce=False)
            os.environ["ZKB200_SLICES"] = "1"
            want = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_affine_out", sc, pts)
            assert want.tobytes() == cpu_affine(curve, sc, pts).tobytes()
            for K in (2, 3, 8):
                os.environ["ZKB200_SLICES"] = str(K)
                got = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_affine_out", sc, pts)
                assert got.tobytes() == want.tobytes(), (n, K)
    finally:
        if old is None:
            os.environ.pop("ZKB200_SLICES", None)
        else:
            os.environ["ZKB200_SLICES"] = old


def _ntt_gen(cv, m):
    """generator of the order-2^m subgroup of Fr^* (5 / 7 generate Fr^* for BN254 / BLS12-381)."""
    g0 = 5 if cv.name == "bn128" else 7
    s = ((cv.r - 1) & -(cv.r - 1)).bit_length() - 1
    assert m <= s
    g = pow(g0, (cv.r - 1) >> m, cv.r)
    assert pow(g, 1 << m, cv.r) == 1 and (m == 0 or pow(g, 1 << (m - 1), cv.r) == cv.r - 1)
    return g


@pytest.mark.parametrize("curve", CURVES)
def test_fr_ntt_next_row(zk, curve):
    """<curve>_poly_mont_ntt_forward / _inverse, bit-identical to the reference C
    (lib/cbits/curves/poly/mont/bn128_poly_mont.c:418-525) and to a direct Python DFT at small sizes."""
    cv = pyec.CURVES[curve]
    rng = random.Random(5)
    mont = lambda x: np.frombuffer(((x * cv.Rr) % cv.r).to_bytes(32, "little"), dtype=np.uint64)
    unmont = lambda row: int.from_bytes(row.tobytes(), "little") * pow(cv.Rr, -1, cv.r) % cv.r
    for m in (0, 1, 2, 3, 5, 8, 9, 10, 11, 14, 18):
        N = 1 << m
        g = _ntt_gen(cv, m)
        xs = [rng.randrange(cv.r) for _ in range(N)] if m <= 11 else None
        src = (np.stack([mont(x) for x in xs]) if xs is not None
               else refs.random_scalars(curve, N, seed=m))               # any value < r is a valid Montgomery residue
        gen = mont(g).copy()
        fwd = zk.ntt(curve, m, gen, src)
        if m <= 8:
            want = [sum(xs[j] * pow(g, j * k, cv.r) for j in range(N)) % cv.r for k in range(N)]
            assert [unmont(fwd[k]) for k in range(N)] == want
        if refs.have_ref():
            f = getattr(refs.ref(), f"{curve}_poly_mont_ntt_forward")
            f.argtypes = [ctypes.c_int, refs.U64P, refs.U64P, refs.U64P]
            f.restype = None
            ref_out = np.zeros_like(src)
            f(m, refs.ptr(gen), refs.ptr(np.ascontiguousarray(src).ravel()), refs.ptr(ref_out.ravel()))
            assert fwd.tobytes() == ref_out.tobytes(), m
            fi = getattr(refs.ref(), f"{curve}_poly_mont_ntt_inverse")
            fi.argtypes = f.argtypes
            fi.restype = None
            ref_inv = np.zeros_like(src)
            fi(m, refs.ptr(gen), refs.ptr(np.ascontiguousarray(src).ravel()), refs.ptr(ref_inv.ravel()))
            assert zk.ntt(curve, m, gen, src, inverse=True).tobytes() == ref_inv.tobytes(), m
        back = zk.ntt(curve, m, gen, fwd, inverse=True)
        assert back.tobytes() == np.ascontiguousarray(src).tobytes(), m


def test_kzg_commit_from_values_stays_on_device(zk):
    """examples/KZG.hs:90-97 (commitInterpolate): coefficients = inverse NTT of the evaluations, commitment =
    MSM of the coefficients over the SRS.  Device-resident chain (zkb200_ntt -> zkb200_msm with device pointers)
    against the same two steps through host memory and against the CPU reference."""
    import torch
    curve, m = "bn128", 12
    cv = pyec.CURVES[curve]
    N = 1 << m
    gen = np.frombuffer(((_ntt_gen(cv, m) * cv.Rr) % cv.r).to_bytes(32, "little"), dtype=np.uint64).copy()
    values = refs.random_scalars(curve, N, seed=4242)
    srs = refs.chain_points(curve, N)
    coeffs_host = zk.ntt(curve, m, gen, values, inverse=True)
    want = zk.msm(curve, coeffs_host, srs, mont=True, out="affine")
    assert want.tobytes() == cpu_affine(curve, coeffs_host, srs, "mont").tobytes()
    d_vals = torch.from_numpy(values.view(np.int64)).cuda()
    d_coef = torch.empty_like(d_vals)
    d_srs = torch.from_numpy(srs.view(np.int64)).cuda()
    torch.cuda.synchronize()
    zk.ntt_device(curve, m, gen, d_vals.data_ptr(), d_coef.data_ptr(), inverse=True)
    assert d_coef.cpu().numpy().view(np.uint64).tobytes() == coeffs_host.tobytes()
    got = zk.msm_device(curve, d_coef.data_ptr(), d_srs.data_ptr(), N, mont=True, out="affine")[0]
    assert got.tobytes() == want.tobytes()


def _g2_ref_chain(curve, n, step=3):
    """n G2 points G, G+D, G+2D, ... (D = step*G) in affine Montgomery bytes, built with the reference's own
    affine addition (CPU); (n, 4*nlimbs) uint64."""
    lib = refs.ref()
    W = 4 * refs.CURVE_LIMBS[curve]           # u64 words per affine G2 point (2 coordinates x 2 x nlimbs)
    gen = (ctypes.c_uint64 * W).in_dll(lib, f"{curve}_G2_affine_gen_G2")
    G = np.frombuffer(bytes(gen), dtype=np.uint64).copy()
    D = G.copy()
    for _ in range(step - 1):
        D = refs.call3(lib, f"{curve}_G2_affine_add", D, G, W)
    out = np.zeros((n, W), np.uint64)
    P = G.copy()
    for i in range(n):
        out[i] = P
        P = refs.call3(lib, f"{curve}_G2_affine_add", P, D, W)
    return out


@pytest.mark.parametrize("curve", CURVES)
def test_g2_msm_next_row(zk, curve):
    """<curve>_G2_proj_MSM_{std,mont}_coeff_{proj,affine}_out against the reference C
    (lib/cbits/curves/g2/proj/bn128_G2_proj.c:498-660): affine bytes identical, proj output equal after the
    reference's G2_proj_to_affine; edge cases as for G1."""
    if not refs.have_ref():
        pytest.skip("needs oracle/_ref")
    lib = refs.ref()
    nl = refs.CURVE_LIMBS[curve]
    W = 4 * nl
    g2 = curve + "_g2"
    pts_all = _g2_ref_chain(curve, 700)
    for n in (0, 1, 2, 3, 33, 700):
        pts = pts_all[:n]
        for form in ("std", "mont"):
            sc = refs.random_scalars(curve, n, seed=n + 5, reduce=(form == "mont"))
            got = zk.call_reference_symbol(f"{curve}_G2_proj_MSM_{form}_coeff_affine_out", sc, pts, npoints=n)
            if n == 0:
                assert got.tobytes() == b"\xff" * (8 * W)
                continue
            want = refs.call_msm(lib, f"{curve}_G2_proj_MSM_{form}_coeff_affine_out", sc.ravel(), pts.ravel(), W, n=n)
            assert got.tobytes() == want.tobytes(), (n, form)
            proj = zk.call_reference_symbol(f"{curve}_G2_proj_MSM_{form}_coeff_proj_out", sc, pts, npoints=n)
            assert refs.call2(lib, f"{curve}_G2_proj_to_affine", proj, W).tobytes() == want.tobytes(), (n, form)
    # exceptional cases: P and -P, repeated points, infinity inputs, zero scalars
    P, Q = pts_all[5].copy(), pts_all[9].copy()
    negP = P.copy()
    cvp = pyec.CURVES[curve].p
    for off in (2 * nl, 3 * nl):               # y = (c0, c1): negate both components
        y = int.from_bytes(P[off:off + nl].tobytes(), "little")
        negP[off:off + nl] = np.frombuffer(((cvp - y) % cvp).to_bytes(8 * nl, "little"), dtype=np.uint64)
    inf = np.full(W, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    cases = {
        "p_minus_p": ([5, 5], [P, negP]),
        "p_minus_p_more": ([5, 5, 9], [P, negP, Q]),
        "repeated": ([3, 3, 3, 3], [P, P, P, P]),
        "inf_inputs": ([9, 2, 1], [inf, P, inf]),
        "zeros": ([0, 0], [P, Q]),
        "full256": ([(1 << 256) - 1, 12345], [P, Q]),
    }
    cv = pyec.CURVES[curve]
    for name, (ks, ps) in cases.items():
        S = np.frombuffer(b"".join(cv.scalar_std_bytes(k) for k in ks), dtype=np.uint64).copy()
        Pb = np.ascontiguousarray(np.stack(ps))
        want = refs.call_msm(lib, f"{curve}_G2_proj_MSM_std_coeff_affine_out", S, Pb.ravel(), W)
        got = zk.call_reference_symbol(f"{curve}_G2_proj_MSM_std_coeff_affine_out", S, Pb)
        assert got.tobytes() == want.tobytes(), name
        for R in (1, 2):   # the same shapes inside the batched-affine pre-reduction (Fp2 coordinates)
            with _Env(ZKB200_AFFINE=R, ZKB200_WINDOW=3):
                got = zk.call_reference_symbol(f"{curve}_G2_proj_MSM_std_coeff_affine_out", S, Pb)
            assert got.tobytes() == want.tobytes(), (name, R)
    # the affine pre-reduction over Fp2, every lane mode
    for n in (33, 257, 700):
        pts, sc = pts_all[:n], refs.random_scalars(curve, n, seed=n + 5, reduce=True)
        want = refs.call_msm(lib, f"{curve}_G2_proj_MSM_mont_coeff_affine_out", sc.ravel(), pts.ravel(), W, n=n)
        for R, mode in ((1, {}), (3, {}), (5, dict(ZKB200_STAGGER=0)), (4, dict(ZKB200_STAGGER=4)), (2, dict(ZKB200_STAGGER=0, ZKB200_AFF_GROUPS=1))):
            with _Env(ZKB200_AFFINE=R, **mode):
                got = zk.call_reference_symbol(f"{curve}_G2_proj_MSM_mont_coeff_affine_out", sc, pts, npoints=n)
            assert got.tobytes() == want.tobytes(), (n, R, mode)
    # wide windows (the row / column bucket reduction of kernels_red.cuh takes over from c = 7) and its switch, over Fp2
    n = 700
    pts, sc = pts_all[:n], refs.random_scalars(curve, n, seed=4242, reduce=True)
    want = refs.call_msm(lib, f"{curve}_G2_proj_MSM_mont_coeff_affine_out", sc.ravel(), pts.ravel(), W, n=n)
    for c in (7, 10, 13):
        for red in (0, 1):
            with _Env(ZKB200_WINDOW=c, ZKB200_RED2D=red):
                got = zk.call_reference_symbol(f"{curve}_G2_proj_MSM_mont_coeff_affine_out", sc, pts, npoints=n)
            assert got.tobytes() == want.tobytes(), (c, red)
    # GPU chain generator for G2 == reference chain; larger size by the split property
    D = refs.call3(lib, f"{curve}_G2_affine_add", refs.call3(lib, f"{curve}_G2_affine_add", pts_all[0].copy(), pts_all[0].copy(), W), pts_all[0].copy(), W)
    assert zk.gen_chain(g2, 50, pts_all[0], D).tobytes() == pts_all[:50].tobytes()
    n = 1 << 14
    big = zk.gen_chain(g2, n, pts_all[0], D)
    sc = refs.random_scalars(curve, n, seed=99)
    whole = zk.msm(g2, sc, big, mont=True, out="affine")
    parts = [zk.msm(g2, sc[i * (n // 4):(i + 1) * (n // 4)], big[i * (n // 4):(i + 1) * (n // 4)], mont=True, out="proj") for i in range(4)]
    assert zk.sum_points(g2, np.stack(parts), "proj", "affine").tobytes() == whole.tobytes()
    m = 1 << 11
    want = refs.call_msm(lib, f"{curve}_G2_proj_MSM_mont_coeff_affine_out", sc[:m].ravel(), big[:m].ravel(), W, n=m)
    assert zk.msm(g2, sc[:m], big[:m], mont=True, out="affine").tobytes() == want.tobytes()


@pytest.mark.parametrize("curve", CURVES)
def test_group_fft_next_row(zk, curve):
    """<curve>_G1_proj_fft_forward / _inverse, bit-identical to the reference C (normalised
    projective output, lib/cbits/curves/g1/proj/bn128_G1_proj.c:678-789), including infinity entries, repeated
    points and non-normalised inputs; forward then inverse is the identity on normalised points."""
    if not refs.have_ref():
        pytest.skip("needs oracle/_ref")
    lib = refs.ref()
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    rng = random.Random(12)
    for m in (0, 1, 2, 3, 5, 7):
        N = 1 << m
        gen = np.frombuffer(((_ntt_gen(cv, m) * cv.Rr) % cv.r).to_bytes(32, "little"), dtype=np.uint64).copy()
        aff = refs.chain_points(curve, N, s0=77, s1=5)
        if N >= 4:
            aff[1] = aff[0]                       # repeated point
            aff[3] = np.uint64(0xFFFFFFFFFFFFFFFF)  # infinity
        proj = zk.batch_from_affine(curve, aff, "proj")
        if N >= 8:                                # a non-normalised representative: (xz, yz, z)
            z = rng.randrange(2, cv.p)
            P = cv.affine_from_bytes(aff[5].tobytes())
            proj[5] = np.frombuffer(cv.fp_to_bytes(P[0] * z % cv.p) + cv.fp_to_bytes(P[1] * z % cv.p) + cv.fp_to_bytes(z), dtype=np.uint64)
        for inverse in (False, True):
            f = getattr(lib, f"{curve}_G1_proj_fft_{'inverse' if inverse else 'forward'}")
            f.argtypes = [ctypes.c_int, refs.U64P, refs.U64P, refs.U64P]
            f.restype = None
            want = np.zeros_like(proj)
            f(m, refs.ptr(gen), refs.ptr(np.ascontiguousarray(proj).ravel()), refs.ptr(want.ravel()))
            got = zk.group_fft(curve, m, gen, proj, inverse=inverse)
            assert got.tobytes() == want.tobytes(), (m, inverse)
            zk.set_glv(0)                       # twiddle products without the endomorphism split: same bytes
            try:
                got = zk.group_fft(curve, m, gen, proj, inverse=inverse)
            finally:
                zk.set_glv(1)
            assert got.tobytes() == want.tobytes(), (m, inverse, "no glv")
        fwd = zk.group_fft(curve, m, gen, proj)
        back = zk.group_fft(curve, m, gen, fwd, inverse=True)
        assert zk.batch_to_affine(curve, back, "proj").tobytes() == aff.tobytes(), m


def test_very_large_batch_is_split(zk):
    """nmsm * W exceeds the 65535 limit of a grid dimension: the batch is processed in sub-batches."""
    curve, n, nmsm = "bn128", 8, 3000          # n = 8 -> c = 2..3 -> W >= 43 (GLV halves): 3000 * W > 65535
    pts = refs.chain_points(curve, n)
    sc = np.stack([refs.random_scalars(curve, n, seed=7000 + i) for i in range(nmsm)])
    got = zk.msm_batch(curve, sc, pts, mont=False, out="affine")
    for i in (0, 1, 511, 512, 1499, 1500, 2999):
        assert got[i].tobytes() == cpu_affine(curve, sc[i], pts).tobytes(), i
    assert zk.last_stats()["nwindows"] * nmsm > 65535


@pytest.mark.parametrize("curve", CURVES)
def test_unreduced_montgomery_scalars(zk, curve):
    """mont_coeff inputs >= r (not canonical): the reference's REDC still decodes them; so must we."""
    n = 64
    pts = refs.chain_points(curve, n)
    rng = np.random.Generator(np.random.PCG64(5))
    sc = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64, endpoint=False)   # full 256-bit values, most >= r
    sc[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    cv = pyec.CURVES[curve]
    sc[1] = np.frombuffer(cv.r.to_bytes(32, "little"), dtype=np.uint64)            # exactly r  -> 0
    sc[2] = np.frombuffer((cv.r + 1).to_bytes(32, "little"), dtype=np.uint64)
    want = cpu_affine(curve, sc, pts, "mont")
    for rep in ("proj", "jac"):
        got = zk.call_reference_symbol(f"{curve}_G1_{rep}_MSM_mont_coeff_affine_out", sc, pts)
        assert got.tobytes() == want.tobytes(), rep


def test_resident_srs_repeated_commits(zk):
    """KZG prover shape: the SRS is uploaded once (zkb200_device_upload), every commitment sends only scalars."""
    curve, n = "bn128", 1 << 12
    srs = refs.chain_points(curve, n)
    res = zk.ResidentPoints(curve, srs)
    try:
        for i in range(3):
            sc = refs.random_scalars(curve, n, seed=600 + i)
            assert res.msm(sc, mont=True).tobytes() == cpu_affine(curve, sc, srs, "mont").tobytes()
        batch = np.stack([refs.random_scalars(curve, n, seed=700 + i) for i in range(4)])
        got = res.msm(batch, mont=True)
        for i in range(4):
            assert got[i].tobytes() == cpu_affine(curve, batch[i], srs, "mont").tobytes()
    finally:
        res.close()


@pytest.mark.parametrize("curve", CURVES)
def test_remaining_twins_of_the_next_rows(zk, curve):
    """Jacobian-input and G2 group FFTs, G2 batch conversions and the *_slow_reference MSM symbols: the rest of the
    symbol families around scope rows 8f.1/8f.3/8f.4, each against the reference C."""
    if not refs.have_ref():
        pytest.skip("needs oracle/_ref")
    lib = refs.ref()
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    m = 4
    N = 1 << m
    gen = np.frombuffer(((_ntt_gen(cv, m) * cv.Rr) % cv.r).to_bytes(32, "little"), dtype=np.uint64).copy()

    def ref_fft(sym, arr):
        f = getattr(lib, sym)
        f.argtypes = [ctypes.c_int, refs.U64P, refs.U64P, refs.U64P]
        f.restype = None
        out = np.zeros_like(arr)
        f(m, refs.ptr(gen), refs.ptr(np.ascontiguousarray(arr).ravel()), refs.ptr(out.ravel()))
        return out

    # G1, Jacobian input
    aff = refs.chain_points(curve, N, s0=31, s1=7)
    aff[2] = np.uint64(0xFFFFFFFFFFFFFFFF)
    jac = zk.batch_from_affine(curve, aff, "jac")
    for inverse in (False, True):
        want = ref_fft(f"{curve}_G1_jac_fft_{'inverse' if inverse else 'forward'}", jac)
        assert zk.group_fft(curve, m, gen, jac, inverse=inverse, group="G1_jac").tobytes() == want.tobytes(), inverse
    # G2: conversions and FFT
    W = 4 * L
    g2aff = _g2_ref_chain(curve, N)
    g2aff[1] = np.uint64(0xFFFFFFFFFFFFFFFF)
    conv = getattr(lib, f"{curve}_G2_proj_batch_from_affine")
    conv.argtypes = [ctypes.c_int, refs.U64P, refs.U64P]
    conv.restype = None
    g2proj = np.zeros((N, 6 * L), np.uint64)
    conv(N, refs.ptr(g2aff.ravel()), refs.ptr(g2proj.ravel()))
    mine = np.zeros_like(g2proj)
    f = getattr(zk.lib(), f"{curve}_G2_proj_batch_from_affine")
    f(N, refs.ptr(g2aff.ravel()), refs.ptr(mine.ravel()))
    assert mine.tobytes() == g2proj.tobytes()
    back = np.zeros_like(g2aff)
    getattr(zk.lib(), f"{curve}_G2_proj_batch_to_affine")(N, refs.ptr(g2proj.ravel()), refs.ptr(back.ravel()))
    assert back.tobytes() == g2aff.tobytes()
    for inverse in (False, True):
        want = ref_fft(f"{curve}_G2_proj_fft_{'inverse' if inverse else 'forward'}", g2proj)
        assert zk.group_fft(curve, m, gen, g2proj, inverse=inverse, group="G2_proj").tobytes() == want.tobytes(), inverse
    # slow_reference symbols: same group element as the reference's slow routine
    n = 25
    pts = refs.chain_points(curve, n)
    sc = refs.random_scalars(curve, n, seed=3, reduce=False)
    want_aff = cpu_affine(curve, sc, pts)
    for rep in ("proj", "jac"):
        sym = f"{curve}_G1_{rep}_MSM_std_coeff_{rep}_out_slow_reference"
        ref_out = refs.call_msm(lib, sym, sc.ravel(), pts.ravel(), 3 * L, n=n)
        assert refs.call2(lib, f"{curve}_G1_{rep}_to_affine", ref_out, 2 * L).tobytes() == want_aff.tobytes()
        got = zk.call_reference_symbol(sym, sc, pts)
        assert refs.call2(lib, f"{curve}_G1_{rep}_to_affine", got, 2 * L).tobytes() == want_aff.tobytes()
    g2pts = _g2_ref_chain(curve, n)
    sym = f"{curve}_G2_proj_MSM_std_coeff_proj_out_slow_reference"
    ref_out = refs.call_msm(lib, sym, sc.ravel(), g2pts.ravel(), 6 * L, n=n)
    got = zk.call_reference_symbol(sym, sc, g2pts)
    assert (refs.call2(lib, f"{curve}_G2_proj_to_affine", got, W).tobytes()
            == refs.call2(lib, f"{curve}_G2_proj_to_affine", ref_out, W).tobytes())


class _Env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("curve", CURVES)
def test_affine_pre_reduction_levels(zk, curve):
    """$ZKB200_AFFINE = R: R levels of pairwise batched-affine sums in front of the XYZZ accumulation
    (kernels_aff.cuh).  Every R must give the reference's bytes: ragged n, narrow windows (long runs), wide
    windows (no runs at all), repeated points, P/-P pairs, infinity inputs, both scalar forms, input slices."""
    cv = pyec.CURVES[curve]
    pts_all = refs.chain_points(curve, 6001)
    for n in (2, 3, 5, 64, 257, 1000, 6001):
        pts, sc = pts_all[:n], refs.random_scalars(curve, n, seed=100 + n, reduce=False)
        want = cpu_affine(curve, sc, pts).tobytes()
        for R in (1, 2, 3, 5, 9):
            with _Env(ZKB200_AFFINE=R):
                got = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_affine_out", sc, pts)
            assert got.tobytes() == want, (n, R)
    # lanes are concurrent streams sharing workspaces: repeat small problems with many windows in every lane mode
    for mode in (dict(ZKB200_STAGGER=0), dict(ZKB200_STAGGER=3), dict(ZKB200_STAGGER=4), dict(ZKB200_STAGGER=0, ZKB200_WGROUPS=4)):
        for rep in range(6):
            n = (64, 100, 257)[rep % 3]
            pts, sc = pts_all[:n], refs.random_scalars(curve, n, seed=300 + rep, reduce=False)
            want = cpu_affine(curve, sc, pts).tobytes()
            with _Env(ZKB200_AFFINE=5, **mode):
                got = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_affine_out", sc, pts)
            assert got.tobytes() == want, (mode, n, rep)
    # window widths: c = 2 gives runs of ~n/2, c = 16 gives almost no equal keys
    n = 3000
    pts, sc = pts_all[:n], refs.random_scalars(curve, n, seed=7, reduce=False)
    want = cpu_affine(curve, sc, pts).tobytes()
    for c in (1, 2, 5, 9, 16):
        for R in (1, 4):
            with _Env(ZKB200_AFFINE=R):
                out = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_proj_out_variable", sc, pts, window_size=c)
            assert to_affine_cpu(curve, "proj", out).tobytes() == want, (c, R)
    # exceptional shapes inside the batched additions
    G = cv.gen
    P, Q = cv.mul(7, G), cv.mul(1234567, G)
    sets = {
        "repeated": ([3] * 9, [P] * 9),
        "repeated_mixed": ([3, 3, 3, 8, 8, 8, 8], [P, P, Q, Q, Q, P, P]),
        "neg_pairs": ([77] * 8, [P, Q, cv.neg(P), cv.neg(Q), P, cv.neg(P), Q, Q]),
        "inf_inputs": ([9, 9, 9, 9, 2, 2], [None, P, None, None, Q, None]),
        "same_scalar_many_points": ([0xABCDEF] * 200, pyec.chain_points(cv, 200, 9, 11)),
        "all_cancel": ([5] * 4, [P, cv.neg(P), Q, cv.neg(Q)]),
    }
    for name, (ks, pl) in sets.items():
        want = cv.affine_to_bytes(cv.msm(ks, pl))
        Pb = np.frombuffer(cv.points_to_bytes(pl), dtype=np.uint64).copy()
        Sm = np.frombuffer(b"".join(cv.scalar_mont_bytes(k) for k in ks), dtype=np.uint64).copy()
        for R in (1, 2, 3):
            for c in (0, 3):
                with _Env(ZKB200_AFFINE=R, ZKB200_WINDOW=c):
                    got = zk.call_reference_symbol(f"{curve}_G1_jac_MSM_mont_coeff_affine_out", Sm, Pb)
                assert got.tobytes() == want, (name, R, c)
    # input slices + batch entry point
    n = 5003
    pts, sc = pts_all[:n], refs.random_scalars(curve, n, seed=3, reduce=True)
    want = cpu_affine(curve, sc, pts, "mont").tobytes()
    for K in (1, 3):
        with _Env(ZKB200_AFFINE=3, ZKB200_SLICES=K):
            got = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_mont_coeff_affine_out", sc, pts)
        assert got.tobytes() == want, K
    nm = 3
    scb = np.concatenate([refs.random_scalars(curve, 1000, seed=40 + i, reduce=True) for i in range(nm)])
    with _Env(ZKB200_AFFINE=2):
        outs = zk.msm_batch(curve, scb.reshape(nm, 1000, 4), pts_all[:1000], mont=True)
    for i in range(nm):
        w = cpu_affine(curve, scb.reshape(nm, 1000, 4)[i], pts_all[:1000], "mont").tobytes()
        assert np.asarray(outs[i]).tobytes() == w, i


@pytest.mark.parametrize("curve", CURVES)
def test_reduction_and_handover_switches(zk, curve):
    """Two code paths behind switches must give the reference's bytes whichever way the switch is set:
    $ZKB200_RED2D (row / column form of the bucket reduction, kernels_red.cuh K5'; windows narrower than 7 bits always
    take the level recurrence) and $ZKB200_AFF_NEXT (a tree level hands the next one its denominators,
    kernels_aff.cuh; needs an even number of merges per segment, so ragged sizes mix both forms)."""
    cv = pyec.CURVES[curve]
    pts_all = refs.chain_points(curve, 6001)
    n = 3000
    pts, sc = pts_all[:n], refs.random_scalars(curve, n, seed=11, reduce=False)
    want = cpu_affine(curve, sc, pts).tobytes()
    for red in (0, 1):
        for c in (6, 7, 8, 10, 13, 16):
            with _Env(ZKB200_RED2D=red):
                out = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_proj_out_variable", sc, pts, window_size=c)
            assert to_affine_cpu(curve, "proj", out).tobytes() == want, (red, c)
        zk.set_glv(0)      # wide windows with enough of them for window groups: 255-bit scalars
        try:
            for c in (19, 21):
                with _Env(ZKB200_RED2D=red):
                    out = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_proj_out_variable", sc, pts, window_size=c)
                assert to_affine_cpu(curve, "proj", out).tobytes() == want, (red, c, "no glv")
        finally:
            zk.set_glv(1)
        for K in (1, 3):   # input slices: several bucket arrays summed on the fly by the first reduction step
            with _Env(ZKB200_RED2D=red, ZKB200_SLICES=K):
                got = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_affine_out", sc, pts)
            assert got.tobytes() == want, (red, K)
    for n in (64, 256, 1000, 4096, 6001):
        pts, sc = pts_all[:n], refs.random_scalars(curve, n, seed=500 + n, reduce=False)
        want = cpu_affine(curve, sc, pts).tobytes()
        for R in (2, 3, 5):
            for glv in (0, 1):
                zk.set_glv(glv)
                try:
                    with _Env(ZKB200_AFF_NEXT=1, ZKB200_AFFINE=R):
                        got = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_std_coeff_affine_out", sc, pts)
                finally:
                    zk.set_glv(1)
                assert got.tobytes() == want, (n, R, glv)
    # the hand-over classifies P + P, P - P and infinity operands a level early
    G = cv.gen
    P, Q = cv.mul(7, G), cv.mul(1234567, G)
    sets = {
        "repeated": ([3] * 16, [P] * 16),
        "repeated_mixed": ([3, 3, 3, 3, 8, 8, 8, 8], [P, P, Q, Q, Q, P, P, P]),
        "neg_pairs": ([77] * 8, [P, Q, cv.neg(P), cv.neg(Q), P, cv.neg(P), Q, Q]),
        "inf_inputs": ([9, 9, 9, 9, 2, 2, 2, 2], [None, P, None, None, Q, None, P, P]),
        "all_cancel": ([5] * 8, [P, cv.neg(P), Q, cv.neg(Q), P, cv.neg(P), Q, cv.neg(Q)]),
    }
    for name, (ks, pl) in sets.items():
        want = cv.affine_to_bytes(cv.msm(ks, pl))
        Pb = np.frombuffer(cv.points_to_bytes(pl), dtype=np.uint64).copy()
        Sm = np.frombuffer(b"".join(cv.scalar_mont_bytes(k) for k in ks), dtype=np.uint64).copy()
        for R in (2, 3):
            for c in (0, 3):
                for glv in (0, 1):
                    zk.set_glv(glv)
                    try:
                        with _Env(ZKB200_AFF_NEXT=1, ZKB200_AFFINE=R, ZKB200_WINDOW=c):
                            got = zk.call_reference_symbol(f"{curve}_G1_jac_MSM_mont_coeff_affine_out", Sm, Pb)
                    finally:
                        zk.set_glv(1)
                    assert got.tobytes() == want, (name, R, c, glv)


def test_affine_pre_reduction_mid_size_vs_threaded_reference(zk):
    curve, n = "bls12_381", 1 << 17
    pts = refs.chain_points(curve, n)
    sc = refs.random_scalars(curve, n, seed=23)
    want = refs.ref_msm_threads(curve, sc, pts, mont=True, nthreads=os.cpu_count() or 4).tobytes()
    for R in (2, 4):
        with _Env(ZKB200_AFFINE=R):
            got = zk.msm(curve, sc, pts, mont=True, out="affine")
        assert got.tobytes() == want, R


@pytest.mark.parametrize("curve", CURVES)
def test_repeated_large_calls_are_stable(zk, curve):
    """The default large-n path runs several streams at once (lanes of the affine pre-reduction, per-group reduction
    and window combination).  A missing dependency between them would show up as an occasional wrong answer:
    repeat the same 2^19-point MSM, resident and from host buffers, and compare with the single-stream XYZZ path."""
    import torch
    n = 1 << 19
    cv = pyec.CURVES[curve]
    p0 = np.frombuffer(cv.affine_to_bytes(cv.mul(0x1234567, cv.gen)), dtype=np.uint64).copy()
    d = np.frombuffer(cv.affine_to_bytes(cv.mul(0x7654321, cv.gen)), dtype=np.uint64).copy()
    pts = torch.empty((n, 2 * cv.nlimbs_p), dtype=torch.int64, device="cuda")
    zk.gen_chain(curve, n, p0, d, device_ptr=pts.data_ptr())
    sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda")
    sc[:, 3] &= (1 << 61) - 1
    torch.cuda.synchronize()
    with _Env(ZKB200_AFFINE=0, ZKB200_WGROUPS=1, ZKB200_AFF_GROUPS=1):
        want = zk.msm_device(curve, sc.data_ptr(), pts.data_ptr(), n, mont=True, out="affine")[0].tobytes()
    h_pts, h_sc = pts.cpu().numpy().view(np.uint64), sc.cpu().numpy().view(np.uint64)
    for mode in (dict(ZKB200_AFFINE=3), dict(ZKB200_AFFINE=3, ZKB200_STAGGER=4), dict(ZKB200_AFFINE=4, ZKB200_STAGGER=0),
                 dict(ZKB200_AFFINE=0)):
        with _Env(**mode):
            for rep in range(8):
                got = zk.msm_device(curve, sc.data_ptr(), pts.data_ptr(), n, mont=True, out="affine")[0].tobytes()
                assert got == want, (mode, rep)
            for rep in range(3):
                assert zk.msm(curve, h_sc, h_pts, mont=True, out="affine").tobytes() == want, (mode, "host", rep)
