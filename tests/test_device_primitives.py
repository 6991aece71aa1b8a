"""Device-level differential tests (B200): the field and group primitives AS COMPILED BY PTXAS -- the inline-PTX carry
chains of fp.cuh, the XYZZ group law of ec.cuh -- run element-wise over arrays of operands (zkb200_selftest_field /
zkb200_selftest_group) and are compared, operand by operand, with the unmodified reference C (oracle/_ref; the pinned
restatement when it is absent).  Bit-exact: every field result is canonical, every group result is a canonical affine
record.

Mirrors the reference's fast-vs-reference field suite (test/src/ZK/Test/Field/AgainstRef.hs:25-60) and the group-law
cases of test/src/ZK/Test/Curve/Properties.hs:425-483 (mixed add vs add, doubling, left / right unit, inverse), for
  <curve>_Fp_mont_{mul,sqr,add,sub,neg,inv}   lib/cbits/curves/fields/mont/bn128_Fp_mont.c:44-109,177-204
  <curve>_Fr_mont_{mul,sqr,to_std}            lib/cbits/curves/fields/mont/bn128_Fr_mont.c:177-199,330-335
  <curve>_G1_proj_{madd_proj_aff,add,dbl}     lib/cbits/curves/g1/proj/bn128_G1_proj.c:230-373
"""
import ctypes

import numpy as np
import pytest

from tests import pyec, refs

pytestmark = pytest.mark.gpu
CURVES = ["bn128", "bls12_381"]
U64P = refs.U64P


@pytest.fixture(scope="module")
def zk():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import zikkurat_algebra_b200 as z
    from zikkurat_algebra_b200 import build
    build.build()
    z.lib()
    return z


def _cpu():
    return (refs.ref(), "") if refs.have_ref() else (refs.oracle(), "zko_")


def _fn(name):
    lib, pre = _cpu()
    return ctypes.cast(getattr(lib, pre + name), ctypes.c_void_p)


def cpu_map3(name, a, b, out_limbs):
    o = refs.oracle()
    o.zko_map3.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, U64P, U64P, U64P]
    o.zko_map3.restype = None
    n = a.shape[0]
    out = np.zeros((n, out_limbs), np.uint64)
    o.zko_map3(_fn(name), n, a.shape[1], b.shape[1], out_limbs, refs.ptr(a), refs.ptr(b), refs.ptr(out))
    return out


def cpu_map2(name, a, out_limbs):
    o = refs.oracle()
    o.zko_map2.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, U64P, U64P]
    o.zko_map2.restype = None
    n = a.shape[0]
    out = np.zeros((n, out_limbs), np.uint64)
    o.zko_map2(_fn(name), n, a.shape[1], out_limbs, refs.ptr(a), refs.ptr(out))
    return out


def field_operands(mod, limbs, n, seed):
    """n canonical residues: every pair of edge values first (0, 1, 2, p-1, p-2, R mod p, 2^k patterns, (p-1)/2, all-ones
    limbs below p), then uniform random ones."""
    R = 1 << (64 * limbs)
    edge = [0, 1, 2, mod - 1, mod - 2, R % mod, (R * R) % mod, (mod - 1) // 2, (mod + 1) // 2, (1 << 32) - 1, 1 << 32, (1 << 64) - 1,
            1 << 64, (1 << (64 * limbs - 3)) % mod, ((1 << (64 * (limbs - 1))) - 1) % mod, mod - (1 << 32), mod - (1 << 64) + 1,
            0xFFFFFFFF00000000FFFFFFFF00000000 % mod]
    pa = [x for x in edge for _ in edge]
    pb = [y for _ in edge for y in edge]
    rng = np.random.Generator(np.random.PCG64(seed))
    raw = rng.integers(0, 1 << 64, size=(2, n, limbs), dtype=np.uint64, endpoint=False)
    top = mod >> (64 * (limbs - 1))
    raw[:, :, limbs - 1] %= np.uint64(top)          # < p, and still covering (almost) the whole range
    a, b = raw[0].copy(), raw[1].copy()
    k = min(len(pa), n)
    for i in range(k):
        a[i] = np.frombuffer(pa[i].to_bytes(8 * limbs, "little"), dtype=np.uint64)
        b[i] = np.frombuffer(pb[i].to_bytes(8 * limbs, "little"), dtype=np.uint64)
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


@pytest.mark.parametrize("curve", CURVES)
def test_fp_ops_device_vs_reference(zk, curve):
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    n = 1_000_000
    a, b = field_operands(cv.p, L, n, seed=101)
    want_mul = cpu_map3(f"{curve}_Fp_mont_mul", a, b, L)
    for op in ("mul", "mul_call"):
        assert np.array_equal(zk.selftest_field(curve, "Fp", op, a, b), want_mul), op
    want_sqr = cpu_map2(f"{curve}_Fp_mont_sqr", a, L)
    for op in ("sqr", "sqr_call"):
        assert np.array_equal(zk.selftest_field(curve, "Fp", op, a), want_sqr), op
    m = 200_000
    a2, b2, c2, d2 = a[:m], b[:m], np.ascontiguousarray(a[::-1][:m]), np.ascontiguousarray(b[::-1][:m])
    want_add = cpu_map3(f"{curve}_Fp_mont_add", a2, b2, L)
    want_sub = cpu_map3(f"{curve}_Fp_mont_sub", a2, b2, L)
    assert np.array_equal(zk.selftest_field(curve, "Fp", "add", a2, b2), want_add)
    assert np.array_equal(zk.selftest_field(curve, "Fp", "sub", a2, b2), want_sub)
    assert np.array_equal(zk.selftest_field(curve, "Fp", "neg", a2), cpu_map2(f"{curve}_Fp_mont_neg", a2, L))
    assert np.array_equal(zk.selftest_field(curve, "Fp", "dbl", a2), cpu_map3(f"{curve}_Fp_mont_add", a2, a2, L))
    # fused a*b + c*d with one reduction == add(mul, mul) of the reference
    want_mul2 = cpu_map3(f"{curve}_Fp_mont_add", np.ascontiguousarray(want_mul[:m]), cpu_map3(f"{curve}_Fp_mont_mul", c2, d2, L), L)
    for op in ("mul2", "mul2_call"):
        assert np.array_equal(zk.selftest_field(curve, "Fp", op, a2, b2, c2, d2), want_mul2), op
    # paired products a*b, a*c with their rows taken in turns (the peel-off step of the batched-affine additions)
    assert np.array_equal(zk.selftest_field(curve, "Fp", "pair_first", a2, b2, c2), np.ascontiguousarray(want_mul[:m]))
    assert np.array_equal(zk.selftest_field(curve, "Fp", "pair_second", a2, b2, c2), cpu_map3(f"{curve}_Fp_mont_mul", a2, c2, L))
    # inversion (Kaliski on the device, binary Euclid in the reference): same canonical value; 0 is skipped by both callers
    k = 20_000
    ai = a[:k].copy()
    zero = ~ai.any(axis=1)
    ai[zero, 0] = 1
    assert np.array_equal(zk.selftest_field(curve, "Fp", "inv", ai), cpu_map2(f"{curve}_Fp_mont_inv", ai, L))


@pytest.mark.parametrize("curve", CURVES)
def test_fr_ops_device_vs_reference(zk, curve):
    cv = pyec.CURVES[curve]
    n = 300_000
    a, b = field_operands(cv.r, 4, n, seed=202)
    assert np.array_equal(zk.selftest_field(curve, "Fr", "mul", a, b), cpu_map3(f"{curve}_Fr_mont_mul", a, b, 4))
    assert np.array_equal(zk.selftest_field(curve, "Fr", "sqr", a), cpu_map2(f"{curve}_Fr_mont_sqr", a, 4))
    assert np.array_equal(zk.selftest_field(curve, "Fr", "from_mont", a), cpu_map2(f"{curve}_Fr_mont_to_std", a, 4))
    # the scalar path accepts non-canonical Montgomery words too: full 256-bit inputs through REDC
    rng = np.random.Generator(np.random.PCG64(7))
    raw = rng.integers(0, 1 << 64, size=(50_000, 4), dtype=np.uint64, endpoint=False)
    assert np.array_equal(zk.selftest_field(curve, "Fr", "from_mont", raw), cpu_map2(f"{curve}_Fr_mont_to_std", raw, 4))


def _affine_bytes(cv, P):
    return np.frombuffer(cv.affine_to_bytes(P) if P is not None else b"\xff" * cv.affine_bytes, dtype=np.uint64)


@pytest.mark.parametrize("curve", CURVES)
def test_group_ops_device_vs_reference(zk, curve):
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    lib, pre = _cpu()
    n = 20_000
    p1 = refs.chain_points(curve, n, s0=0x1111, s1=0x2222)
    p2 = refs.chain_points(curve, n, s0=0x777, s1=0x2222)            # generic: p2 != +-p1
    INF = np.full(2 * L, 0xFFFFFFFFFFFFFFFF, np.uint64)
    # exceptional shapes (Properties.hs:425-430, bn128_G1_proj.c:334-358): same point, opposite point, units
    neg = cpu_map2(f"{curve}_G1_affine_neg", p1[:64].copy(), 2 * L)
    p2[0:16] = p1[0:16]                # P + P  -> doubling branch
    p2[16:32] = neg[16:32]             # P + (-P) -> infinity
    p1[32:40] = INF                    # inf + P
    p2[40:48] = INF                    # P + inf
    p1[48:52] = INF; p2[48:52] = INF   # inf + inf
    z1, z2 = field_operands(cv.p, L, n, seed=303)
    for z in (z1, z2):                 # scales must be invertible
        z[~z.any(axis=1), 0] = 1
    z1[100:200] = np.frombuffer((cv.R % cv.p).to_bytes(8 * L, "little"), dtype=np.uint64)   # Z = 1: the "fresh bucket" shape
    # reference: lift both operands to projective, use the reference's own operation, convert back
    q1 = cpu_map2(f"{curve}_G1_proj_from_affine", p1, 3 * L)
    q2 = cpu_map2(f"{curve}_G1_proj_from_affine", p2, 3 * L)
    to_aff = lambda q: cpu_map2(f"{curve}_G1_proj_to_affine", q, 2 * L)
    want_add = to_aff(cpu_map3(f"{curve}_G1_proj_add", q1, q2, 3 * L))
    want_dbl = to_aff(cpu_map2(f"{curve}_G1_proj_dbl", q1, 3 * L))
    for op in ("add", "add_calls"):
        got = zk.selftest_group(curve, op, p1, z1, p2, z2)
        assert np.array_equal(got, want_add), op
    for op in ("madd", "madd_calls"):
        # the bucket insertion skips infinity operands on the right (as k_accumulate does) -> result = left operand
        got = zk.selftest_group(curve, op, p1, z1, p2, z2)
        assert np.array_equal(got, want_add), op
    # mixed addition of the reference itself where its preconditions hold (finite right operand): same element
    fin = ~(p2 == INF).all(axis=1)
    want_madd = to_aff(cpu_map3(f"{curve}_G1_proj_madd_proj_aff", np.ascontiguousarray(q1[fin]), np.ascontiguousarray(p2[fin]), 3 * L))
    assert np.array_equal(want_madd, want_add[fin])
    assert np.array_equal(zk.selftest_group(curve, "dbl", p1, z1, p2, z2), want_dbl)
    assert np.array_equal(zk.selftest_group(curve, "dbl_affine", p1, z1, p2, z2), want_dbl)
    # spot-check of the reference against the independent Python model (pins the checker itself)
    for i in (0, 5, 20, 33, 41, 50, 1000):
        a = cv.affine_from_bytes(p1[i].tobytes()) if not (p1[i] == INF).all() else None
        b = cv.affine_from_bytes(p2[i].tobytes()) if not (p2[i] == INF).all() else None
        assert np.array_equal(want_add[i], _affine_bytes(cv, cv.add(a, b)))
