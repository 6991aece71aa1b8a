"""Pin the CPU oracles (no GPU): restatement == unmodified reference C == independent Python EC.

Mirrors the reference's own strategy: fast-C vs slow reference comparisons
(test/src/ZK/Test/Field/AgainstRef.hs:25-60), the curve-law edge cases of
test/src/ZK/Test/Curve/Properties.hs:425-483, and the examples/MSM.hs:65-77 three-way MSM equality.
"""
import random

import numpy as np
import pytest

from tests import pyec, refs

CURVES = ["bn128", "bls12_381"]


def _rand_fp(cv, rng):
    return rng.randrange(cv.p)


def _fp_arr(cv, x):
    return np.frombuffer(cv.fp_to_bytes(x), dtype=np.uint64).copy()


@pytest.mark.parametrize("curve", CURVES)
def test_field_ops_three_way(curve):
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    rng = random.Random(1)
    edge = [0, 1, cv.p - 1, cv.R % cv.p, (cv.p - 1) // 2]
    vals = [(a, b) for a in edge for b in edge] + [(_rand_fp(cv, rng), _rand_fp(cv, rng)) for _ in range(300)]
    libs = [(refs.oracle(), f"zko_{curve}_Fp_mont_")]
    if refs.have_ref():
        libs.append((refs.ref(), f"{curve}_Fp_mont_"))
    for a, b in vals:
        A, B = _fp_arr(cv, a), _fp_arr(cv, b)
        for lib, pre in libs:
            assert cv.fp_from_bytes(refs.call3(lib, pre + "mul", A, B, L).tobytes()) == a * b % cv.p
            assert cv.fp_from_bytes(refs.call3(lib, pre + "add", A, B, L).tobytes()) == (a + b) % cv.p
            assert cv.fp_from_bytes(refs.call3(lib, pre + "sub", A, B, L).tobytes()) == (a - b) % cv.p
            assert cv.fp_from_bytes(refs.call2(lib, pre + "neg", A, L).tobytes()) == (-a) % cv.p
            if a:
                assert cv.fp_from_bytes(refs.call2(lib, pre + "inv", A, L).tobytes()) == pow(a, -1, cv.p)


@pytest.mark.parametrize("curve", CURVES)
def test_fr_to_std(curve):
    cv = pyec.CURVES[curve]
    rng = random.Random(2)
    for k in [0, 1, cv.r - 1] + [rng.randrange(cv.r) for _ in range(100)] + [rng.randrange(1 << 256) for _ in range(20)]:
        enc = np.frombuffer(k.to_bytes(32, "little"), dtype=np.uint64).copy()
        want = k * pow(cv.Rr, -1, cv.r) % cv.r
        got = int.from_bytes(refs.call2(refs.oracle(), f"zko_{curve}_Fr_mont_to_std", enc, 4).tobytes(), "little")
        assert got == want
        if refs.have_ref():
            got = int.from_bytes(refs.call2(refs.ref(), f"{curve}_Fr_mont_to_std", enc, 4).tobytes(), "little")
            assert got == want


@pytest.mark.parametrize("curve", CURVES)
def test_group_ops_edge_cases(curve):
    """mixed add vs Python EC incl. doubling / inverse / infinity operands (Curve/Properties.hs:467-483)."""
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    G = cv.gen
    P, Q = cv.mul(5, G), cv.mul(11, G)
    cases = [(P, Q), (P, P), (P, cv.neg(P)), (None, Q), (P, None), (None, None)]
    for repr_, from_bytes in (("proj", cv.proj_from_bytes), ("jac", cv.jac_from_bytes)):
        for A, B in cases:
            a_aff = np.frombuffer(cv.affine_to_bytes(A), dtype=np.uint64).copy()
            b_aff = np.frombuffer(cv.affine_to_bytes(B), dtype=np.uint64).copy()
            for lib, pre in [(refs.oracle(), f"zko_{curve}_G1_")] + ([(refs.ref(), f"{curve}_G1_")] if refs.have_ref() else []):
                if pre.startswith("zko_") and repr_ == "jac":
                    # restatement exposes jac_from_affine only through madd; build via madd onto infinity
                    inf = np.frombuffer(cv.affine_to_bytes(None), dtype=np.uint64).copy()
                    a_rep = _jac_from_affine(lib, pre, cv, a_aff)
                else:
                    a_rep = refs.call2(lib, pre + f"{repr_}_from_affine", a_aff, 3 * L)
                s = refs.call3(lib, pre + f"{repr_}_madd_{repr_}_aff", a_rep, b_aff, 3 * L)
                assert from_bytes(s.tobytes()) == cv.add(A, B), (repr_, pre, A is None, B is None)
                aff = refs.call2(lib, pre + f"{repr_}_to_affine", s, 2 * L)
                assert aff.tobytes() == cv.affine_to_bytes(cv.add(A, B))


def _jac_from_affine(lib, pre, cv, a_aff):
    L = cv.nlimbs_p
    one = np.frombuffer(((cv.R % cv.p).to_bytes(cv.fp_bytes, "little")), dtype=np.uint64)
    inf = np.concatenate([one, one, np.zeros(L, np.uint64)])
    return refs.call3(lib, pre + "jac_madd_jac_aff", np.ascontiguousarray(inf), a_aff, 3 * L)


def _inputs(cv, n, seed, full256=False):
    rng = random.Random(seed)
    pts = pyec.chain_points(cv, n, s0=rng.randrange(1, 1 << 40), s1=rng.randrange(1, 1 << 40))
    ks = [rng.randrange(1 << 256) if full256 else rng.randrange(cv.r) for _ in range(n)]
    return ks, pts


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("n", [0, 1, 2, 3, 40])
def test_msm_all_entry_points_vs_python(curve, n):
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    ks, pts = _inputs(cv, n, seed=100 + n)
    want = cv.affine_to_bytes(cv.msm(ks, pts))
    P = np.frombuffer(cv.points_to_bytes(pts), dtype=np.uint64).copy()
    S_std = np.frombuffer(b"".join(cv.scalar_std_bytes(k) for k in ks), dtype=np.uint64).copy()
    S_mont = np.frombuffer(b"".join(cv.scalar_mont_bytes(k) for k in ks), dtype=np.uint64).copy()
    libs = [(refs.oracle(), "zko_")] + ([(refs.ref(), "")] if refs.have_ref() else [])
    for lib, pre in libs:
        for rep in ("proj", "jac"):
            for form, S in (("std", S_std), ("mont", S_mont)):
                if n == 0 and pre == "":
                    continue  # the reference evaluates log2(0): not a defined case for the shipped C
                got = refs.call_msm(lib, f"{pre}{curve}_G1_{rep}_MSM_{form}_coeff_affine_out", S, P, 2 * L, n=n)
                assert got.tobytes() == want, (pre, rep, form)
                out = refs.call_msm(lib, f"{pre}{curve}_G1_{rep}_MSM_{form}_coeff_{rep}_out", S, P, 3 * L, n=n)
                dec = cv.proj_from_bytes if rep == "proj" else cv.jac_from_bytes
                assert cv.affine_to_bytes(dec(out.tobytes())) == want


@pytest.mark.parametrize("curve", CURVES)
def test_msm_semantics_observed_on_reference(curve):
    """SURVEY.md section 8b 'observed semantics': zero scalars, P/-P, repeated points, infinity inputs,
    un-reduced 256-bit std scalars."""
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    G = cv.gen
    P = cv.mul(7, G)
    sets = {
        "all_zero": ([0, 0, 0], [G, P, cv.mul(3, G)]),
        "p_minus_p": ([5, 5], [P, cv.neg(P)]),
        "repeated": ([3, 4, 5, 6], [P, P, P, P]),
        "inf_input": ([9, 2, 1], [None, P, None]),
        "full256": ([(1 << 256) - 1, (1 << 255) + 12345], [G, P]),
        "topbit_lowbit": ([1 << 255, 1], [G, P]),
    }
    libs = [(refs.oracle(), "zko_")] + ([(refs.ref(), "")] if refs.have_ref() else [])
    for name, (ks, pts) in sets.items():
        want = cv.affine_to_bytes(cv.msm(ks, pts))
        Pb = np.frombuffer(cv.points_to_bytes(pts), dtype=np.uint64).copy()
        S = np.frombuffer(b"".join(cv.scalar_std_bytes(k) for k in ks), dtype=np.uint64).copy()
        for lib, pre in libs:
            for rep in ("proj", "jac"):
                got = refs.call_msm(lib, f"{pre}{curve}_G1_{rep}_MSM_std_coeff_affine_out", S, Pb, 2 * L)
                assert got.tobytes() == want, (name, pre, rep)


@pytest.mark.parametrize("curve", CURVES)
def test_restatement_equals_reference_mid_size(curve):
    """n = 2^10: restatement, reference (1 thread), reference 4-shard split: identical affine bytes."""
    if not refs.have_ref():
        pytest.skip("oracle/_ref not built")
    L = refs.CURVE_LIMBS[curve]
    n = 1 << 10
    pts = refs.chain_points(curve, n)
    sc = refs.random_scalars(curve, n, seed=7)
    a = refs.call_msm(refs.oracle(), f"zko_{curve}_G1_proj_MSM_std_coeff_affine_out", sc.ravel(), pts.ravel(), 2 * L)
    b = refs.call_msm(refs.ref(), f"{curve}_G1_proj_MSM_std_coeff_affine_out", sc.ravel(), pts.ravel(), 2 * L)
    c = refs.call_msm(refs.ref(), f"{curve}_G1_jac_MSM_mont_coeff_affine_out", sc.ravel(), pts.ravel(), 2 * L)
    d = refs.ref_msm_threads(curve, sc, pts, mont=False, nthreads=4)
    e = refs.ref_msm_threads(curve, sc, pts, mont=False, nthreads=3, use_ref=False)
    assert a.tobytes() == b.tobytes() == d.tobytes() == e.tobytes()
    c2 = refs.call_msm(refs.oracle(), f"zko_{curve}_G1_jac_MSM_mont_coeff_affine_out", sc.ravel(), pts.ravel(), 2 * L)
    assert c.tobytes() == c2.tobytes()


@pytest.mark.parametrize("curve", CURVES)
def test_chain_generator_matches_python(curve):
    cv = pyec.CURVES[curve]
    want = cv.points_to_bytes(pyec.chain_points(cv, 17))
    assert refs.chain_points(curve, 17).tobytes() == want
