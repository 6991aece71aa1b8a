"""ctypes access to the CPU oracles + synthetic-input generators shared by tests and bench.

TEST INFRASTRUCTURE ONLY (see oracle/msm_oracle.c header).  Two oracles:

* ``ref()``    -> oracle/_ref/libzk_ref.so : the unmodified reference C (built here from
                  /root/reference by oracle/Makefile; the prebuilt .so travels to the GPU box).
* ``oracle()`` -> oracle/libzk_oracle.so   : our plain-C restatement (symbols prefixed ``zko_``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from functools import lru_cache

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libzk_ref.so")
ORACLE_SO = os.path.join(ORACLE_DIR, "libzk_oracle.so")

U64P = ctypes.POINTER(ctypes.c_uint64)

CURVE_LIMBS = {"bn128": 4, "bls12_381": 6}


def build_oracles() -> None:
    """Compile the oracles (idempotent).  `ref` is only rebuilt when /root/reference exists."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle", "ref"], check=True)


def ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(U64P)


@lru_cache(maxsize=None)
def oracle() -> ctypes.CDLL:
    if not os.path.exists(ORACLE_SO):
        build_oracles()
    return ctypes.CDLL(ORACLE_SO)


@lru_cache(maxsize=None)
def ref() -> ctypes.CDLL:
    if not os.path.exists(REF_SO):
        build_oracles()
    if not os.path.exists(REF_SO):
        raise FileNotFoundError(REF_SO)
    return ctypes.CDLL(REF_SO)


def have_ref() -> bool:
    try:
        ref()
        return True
    except (OSError, FileNotFoundError, subprocess.CalledProcessError):
        return False


MSM_ARGTYPES = [ctypes.c_int, U64P, U64P, U64P, ctypes.c_int]


def call_msm(lib: ctypes.CDLL, symbol: str, scalars: np.ndarray, points: np.ndarray, out_limbs: int,
             n: int | None = None, nlimbs: int = 4) -> np.ndarray:
    """Call a reference-ABI MSM symbol: void f(int n, const u64* expos, const u64* grps, u64* tgt, int nlimbs)."""
    f = getattr(lib, symbol)
    f.argtypes = MSM_ARGTYPES
    f.restype = None
    if n is None:
        n = scalars.size // nlimbs
    out = np.zeros(out_limbs, dtype=np.uint64)
    s = scalars if scalars.size else np.zeros(1, np.uint64)
    p = points if points.size else np.zeros(1, np.uint64)
    f(n, ptr(s), ptr(p), ptr(out), nlimbs)
    return out


def call2(lib, symbol: str, a: np.ndarray, out_limbs: int) -> np.ndarray:
    f = getattr(lib, symbol)
    f.argtypes = [U64P, U64P]
    f.restype = None
    out = np.zeros(out_limbs, np.uint64)
    f(ptr(a), ptr(out))
    return out


def call3(lib, symbol: str, a: np.ndarray, b: np.ndarray, out_limbs: int) -> np.ndarray:
    f = getattr(lib, symbol)
    f.argtypes = [U64P, U64P, U64P]
    f.restype = None
    out = np.zeros(out_limbs, np.uint64)
    f(ptr(a), ptr(b), ptr(out))
    return out


# ---------------------------------------------------------------------------------------------
# synthetic inputs (identical bytes go to the CUDA path and to the oracle)

R_MOD = {
    "bn128": 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    "bls12_381": 52435875175126190479447740508185965837690552500527637822603658699938581184513,
}


def random_scalars(curve: str, n: int, seed: int, mont: bool = False, reduce: bool = True) -> np.ndarray:
    """n scalars as (n,4) uint64 little-endian limbs.  `reduce`: uniform in [0,r) (approximately:
    a 256-bit draw conditionally reduced); otherwise arbitrary 256-bit integers.  `mont`: the
    *encoding* handed to a mont_coeff entry point -- for throughput inputs any value < r is as good
    as any other, so the draw itself is used as the Montgomery representative."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64, endpoint=False)
    if reduce:
        # force < r by clearing enough top bits (r > 2^253 for both curves)
        a[:, 3] &= np.uint64((1 << 61) - 1)
    return np.ascontiguousarray(a)


def counter_scalars(seed: int, start: int, n: int, bits: int = 253) -> np.ndarray:
    """Scalars of a GLOBAL workload as a pure function of (seed, global index): limb j of scalar i is
    splitmix64(seed * 2^32 + 4*i + j), the top limb cut so that the value is < 2^bits (253: < r for both
    curves, so the same words serve as a standard integer or as a Montgomery representative).  Any rank
    regenerates exactly its slice [start, start + n); tests/golden/make_big_golden.py generates the whole
    vector with the same function.  Returns (n, 4) uint64."""
    with np.errstate(over="ignore"):
        idx = (np.arange(start * 4, (start + n) * 4, dtype=np.uint64) + np.uint64((seed << 32) & 0xFFFFFFFFFFFFFFFF))
        z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    a = z.reshape(n, 4)
    if bits < 256:
        a[:, 3] &= np.uint64((1 << (bits - 192)) - 1)
    return np.ascontiguousarray(a)


def chain_base(curve: str, s0: int = 0x1234567, s1: int = 0x7654321, start: int = 0):
    """(P_start, D) of the global point chain P_i = (s0 + i*s1)*G as affine Montgomery limbs (Python big ints only)."""
    from . import pyec  # local import: tests package
    cv = pyec.CURVES[curve]
    p0 = np.frombuffer(cv.affine_to_bytes(cv.mul(s0 + start * s1, cv.gen)), dtype=np.uint64).copy()
    d = np.frombuffer(cv.affine_to_bytes(cv.mul(s1, cv.gen)), dtype=np.uint64).copy()
    return p0, d


def chain_points(curve: str, n: int, s0: int = 0x1234567, s1: int = 0x7654321, start: int = 0, nthreads: int = 1) -> np.ndarray:
    """(n, 2L) uint64 affine Montgomery points P_i = (s0 + (start + i)*s1)*G via the oracle's chain generator
    (`nthreads` > 1: contiguous blocks, each started from its own multiple of G)."""
    import threading
    L = CURVE_LIMBS[curve]
    out = np.zeros((n, 2 * L), dtype=np.uint64)
    f = getattr(oracle(), f"zko_{curve}_gen_chain")
    f.argtypes = [ctypes.c_long, U64P, U64P, U64P]
    f.restype = None
    if n == 0:
        return out
    nthreads = max(1, min(nthreads, n // 4096 or 1))
    bounds = [n * k // nthreads for k in range(nthreads + 1)]
    bases = [chain_base(curve, s0, s1, start + bounds[k]) for k in range(nthreads)]

    def work(k):
        lo, hi = bounds[k], bounds[k + 1]
        if hi > lo:
            f(hi - lo, ptr(bases[k][0]), ptr(bases[k][1]), out[lo:hi].ctypes.data_as(U64P))

    ths = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return out


def ref_msm_threads(curve: str, scalars: np.ndarray, points: np.ndarray, mont: bool, nthreads: int,
                    use_ref: bool = True) -> np.ndarray:
    """Large-n CPU answer: T contiguous shards in T threads (ctypes drops the GIL), partial proj
    results combined with the reference's proj_add, then proj_to_affine."""
    import threading
    L = CURVE_LIMBS[curve]
    n = scalars.shape[0]
    if not (use_ref and have_ref()):
        f = getattr(oracle(), f"zko_{curve}_msm_threads")
        f.argtypes = [ctypes.c_long, U64P, U64P, U64P, ctypes.c_int, ctypes.c_int]
        f.restype = None
        out = np.zeros(2 * L, np.uint64)
        f(n, ptr(scalars), ptr(points), ptr(out), int(mont), nthreads)
        return out
    lib = ref()
    sym = f"{curve}_G1_proj_MSM_{'mont' if mont else 'std'}_coeff_proj_out"
    parts = [None] * nthreads

    def work(k):
        lo, hi = n * k // nthreads, n * (k + 1) // nthreads
        parts[k] = call_msm(lib, sym, scalars[lo:hi], points[lo:hi], 3 * L, n=hi - lo)

    ths = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    acc = parts[0]
    for k in range(1, nthreads):
        acc = call3(lib, f"{curve}_G1_proj_add", acc, parts[k], 3 * L)
    return call2(lib, f"{curve}_G1_proj_to_affine", acc, 2 * L)
