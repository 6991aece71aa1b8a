"""zikkurat_algebra_b200 -- B200 (sm_100a) G1 multi-scalar multiplication for BN254 / BLS12-381.

Host-side mirror (Python, ctypes) of the reference's MSM interface over the C-ABI shared library
``lib/libzkmsm_b200.so`` (include/zk_msm_b200.h).  Mirrors the generated Haskell bindings

    ZK.Algebra.Curves.<Curve>.G1.Proj.msm / msmStd      lib/src/ZK/Algebra/Curves/BN128/G1/Proj.hs:228-263
    ZK.Algebra.Curves.<Curve>.G1.Affine.msm / msmStd    lib/src/ZK/Algebra/Curves/BN128/G1/Affine.hs:144-149

over flat little-endian uint64 arrays (the FlatArray layout of lib/src/ZK/Algebra/Class/Flat.hs:81-90).
There is no CPU implementation in this package: importing works anywhere (so that the library's exported
symbols can be inspected), every compute call needs a CUDA device and aborts without one.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

__all__ = [
    "CURVES", "lib", "lib_path", "msm", "msm_std", "msm_batch", "msm_device", "sum_points", "batch_to_affine", "batch_from_affine", "CONVERT_SYMBOLS", "ntt", "ntt_device", "NTT_SYMBOLS", "G2_SYMBOLS", "GFFT_SYMBOLS", "EXTRA_SYMBOLS", "group_fft", "call_reference_symbol",
    "last_stats", "imad_peak", "set_device", "set_devices", "gen_chain", "launch_count", "ResidentPoints", "REFERENCE_SYMBOLS", "EXTENSION_SYMBOLS",
    "msm_to_device", "sum_points_device", "last_srs_hit", "release_workspaces", "srs_cache_drop", "last_op_ms", "set_glv", "selftest_field", "selftest_group", "FIELD_OPS", "GROUP_OPS",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libzkmsm_b200.so")

# nlimbs_p = uint64 words per point COORDINATE (G2 coordinates are Fp2 elements: c0 || c1)
CURVES = {"bn128": dict(id=0, nlimbs_p=4), "bls12_381": dict(id=1, nlimbs_p=6),
          "bn128_g2": dict(id=2, nlimbs_p=8), "bls12_381_g2": dict(id=3, nlimbs_p=12)}
OUT_PROJ, OUT_JAC, OUT_AFFINE, OUT_XYZZ = 0, 1, 2, 3
_OUT = {"proj": OUT_PROJ, "jac": OUT_JAC, "affine": OUT_AFFINE, "xyzz": OUT_XYZZ}
_OUT_COORDS = {OUT_PROJ: 3, OUT_JAC: 3, OUT_AFFINE: 2, OUT_XYZZ: 4}
HOST, DEVICE = 0, 1

REFERENCE_SYMBOLS = [
    f"{c}_G1_{r}_MSM_{f}_coeff_{o}_out"
    for c in ("bn128", "bls12_381")
    for r in ("proj", "jac")
    for f in ("std", "mont")
    for o in (r, "affine")
] + [f"{c}_G1_{r}_MSM_std_coeff_{r}_out_variable" for c in ("bn128", "bls12_381") for r in ("proj", "jac")]
CONVERT_SYMBOLS = [f"{c}_G1_{r}_batch_{d}_affine" for c in ("bn128", "bls12_381") for r in ("proj", "jac") for d in ("to", "from")]
NTT_SYMBOLS = [f"{c}_poly_mont_ntt_{d}" for c in ("bn128", "bls12_381") for d in ("forward", "inverse")]
G2_SYMBOLS = [f"{c}_G2_proj_MSM_{f}_coeff_{o}_out" for c in ("bn128", "bls12_381") for f in ("std", "mont") for o in ("proj", "affine")]
GFFT_SYMBOLS = [f"{c}_{g}_fft_{d}" for c in ("bn128", "bls12_381") for g in ("G1_proj", "G1_jac", "G2_proj") for d in ("forward", "inverse")]
EXTRA_SYMBOLS = ([f"{c}_G2_proj_batch_{d}_affine" for c in ("bn128", "bls12_381") for d in ("to", "from")] +
                 [f"{c}_{g}_out_slow_reference" for c in ("bn128", "bls12_381")
                  for g in ("G1_proj_MSM_std_coeff_proj", "G1_jac_MSM_std_coeff_jac", "G2_proj_MSM_std_coeff_proj")])
EXTENSION_SYMBOLS = ["zkb200_msm", "zkb200_sum_points", "zkb200_set_device", "zkb200_last_stats", "zkb200_imad_peak",
                     "zkb200_version", "zkb200_gen_chain", "zkb200_launch_count", "zkb200_set_devices", "zkb200_ntt", "zkb200_device_upload", "zkb200_device_free", "zkb200_last_affine_levels",
                     "zkb200_msm_ex", "zkb200_sum_points_ex", "zkb200_last_srs_hit", "zkb200_srs_cache_drop", "zkb200_release_workspaces",
                     "zkb200_selftest_field", "zkb200_selftest_group", "zkb200_last_op_ms", "zkb200_set_glv"]
FIELD_IDS = {("bn128", "Fp"): 0, ("bls12_381", "Fp"): 1, ("bn128", "Fr"): 2, ("bls12_381", "Fr"): 3}
FIELD_OPS = {"mul": 0, "sqr": 1, "mul2": 2, "add": 3, "sub": 4, "neg": 5, "inv": 6, "mul_call": 7, "sqr_call": 8, "mul2_call": 9,
             "dbl": 10, "from_mont": 11, "pair_first": 12, "pair_second": 13}
GROUP_OPS = {"madd": 0, "madd_calls": 1, "add": 2, "add_calls": 3, "dbl": 4, "dbl_affine": 5}

_U64P = ctypes.POINTER(ctypes.c_uint64)
_lib: Optional[ctypes.CDLL] = None


def lib_path() -> str:
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    """The C-ABI library.  Fails loudly when it has not been built (python zikkurat_algebra_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python zikkurat_algebra_b200/build.py` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        L = ctypes.CDLL(_LIB_PATH)
        L.zkb200_msm.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _U64P]
        L.zkb200_msm.restype = None
        L.zkb200_sum_points.argtypes = [ctypes.c_int, ctypes.c_int, _U64P, ctypes.c_int, ctypes.c_int, _U64P]
        L.zkb200_sum_points.restype = None
        L.zkb200_msm_ex.argtypes = L.zkb200_msm.argtypes[:-1] + [ctypes.c_void_p, ctypes.c_int]
        L.zkb200_msm_ex.restype = None
        L.zkb200_sum_points_ex.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _U64P]
        L.zkb200_sum_points_ex.restype = None
        L.zkb200_last_srs_hit.restype = ctypes.c_int
        L.zkb200_release_workspaces.restype = None
        L.zkb200_selftest_field.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long] + [ctypes.c_void_p] * 4 + [_U64P]
        L.zkb200_selftest_field.restype = None
        L.zkb200_selftest_group.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long] + [ctypes.c_void_p] * 4 + [_U64P]
        L.zkb200_selftest_group.restype = None
        L.zkb200_set_device.argtypes = [ctypes.c_int]
        L.zkb200_set_device.restype = None
        L.zkb200_last_stats.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int),
                                        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong)]
        L.zkb200_last_stats.restype = None
        L.zkb200_imad_peak.argtypes = [ctypes.c_int, ctypes.c_int]
        L.zkb200_imad_peak.restype = ctypes.c_double
        L.zkb200_version.restype = ctypes.c_char_p
        L.zkb200_gen_chain.argtypes = [ctypes.c_int, ctypes.c_ulonglong, ctypes.c_long, _U64P, _U64P, ctypes.c_void_p, ctypes.c_int]
        L.zkb200_gen_chain.restype = None
        L.zkb200_launch_count.restype = ctypes.c_longlong
        L.zkb200_set_devices.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        L.zkb200_set_devices.restype = None
        L.zkb200_ntt.argtypes = [ctypes.c_int, ctypes.c_int, _U64P, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                 ctypes.c_int]
        L.zkb200_ntt.restype = None
        L.zkb200_device_upload.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.zkb200_device_upload.restype = ctypes.c_void_p
        L.zkb200_device_free.argtypes = [ctypes.c_void_p]
        L.zkb200_device_free.restype = None
        for name in NTT_SYMBOLS + GFFT_SYMBOLS:
            f = getattr(L, name)
            f.argtypes = [ctypes.c_int, _U64P, _U64P, _U64P]
            f.restype = None
        for name in CONVERT_SYMBOLS + EXTRA_SYMBOLS[:4]:
            f = getattr(L, name)
            f.argtypes = [ctypes.c_int, _U64P, _U64P]
            f.restype = None
        for name in REFERENCE_SYMBOLS + G2_SYMBOLS + EXTRA_SYMBOLS[4:]:
            f = getattr(L, name)
            f.argtypes = [ctypes.c_int, _U64P, _U64P, _U64P, ctypes.c_int] + ([ctypes.c_int] if name.endswith("_variable") else [])
            f.restype = None
        _lib = L
    return _lib


def _as_u64(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_U64P)


def set_device(device: int) -> None:
    lib().zkb200_set_device(int(device))


def _device_count() -> int:
    try:
        import torch
        return int(torch.cuda.device_count())
    except Exception:          # torch is optional plumbing: without it the library itself rejects bad indices
        return 1 << 30


def set_devices(devices) -> None:
    """Shard every subsequent host-buffer call over these GPUs from inside the library ([] = single device)."""
    devices = [int(d) for d in devices]
    cnt = _device_count()
    for d in devices:
        if d < 0 or d >= cnt:
            raise ValueError(f"device index {d} out of range (visible devices: {cnt})")
    arr = (ctypes.c_int * max(1, len(devices)))(*devices)
    lib().zkb200_set_devices(arr, len(devices))


def call_reference_symbol(name: str, scalars: np.ndarray, points: np.ndarray, npoints: Optional[int] = None,
                          expo_nlimbs: int = 4, window_size: Optional[int] = None) -> np.ndarray:
    """Call one of the reference-named entry points exactly as the reference's FFI would
    (void f(int npoints, const uint64_t* expos, const uint64_t* grps, uint64_t* tgt, int expo_nlimbs))."""
    curve = "bls12_381" if name.startswith("bls12_381") else "bn128"
    L = CURVES[curve + ("_g2" if "_G2_" in name else "")]["nlimbs_p"]
    coords = 2 if "affine_out" in name else 3
    s = _as_u64(scalars).ravel()
    p = _as_u64(points).ravel()
    n = s.size // expo_nlimbs if npoints is None else npoints
    out = np.zeros(coords * L, dtype=np.uint64)
    s_ = s if s.size else np.zeros(1, np.uint64)
    p_ = p if p.size else np.zeros(1, np.uint64)
    f = getattr(lib(), name)
    if name.endswith("_variable"):
        f(n, _ptr(s_), _ptr(p_), _ptr(out), expo_nlimbs, int(window_size or 0))
    else:
        f(n, _ptr(s_), _ptr(p_), _ptr(out), expo_nlimbs)
    return out


def msm_batch(curve: str, scalars: np.ndarray, points: np.ndarray, mont: bool = True, out: str = "affine",
              window: int = 0, expo_nlimbs: int = 4) -> np.ndarray:
    """nmsm independent MSMs over one shared point array.  scalars: (nmsm, n, expo_nlimbs) uint64,
    points: (n, 2L) uint64 (host arrays).  Returns (nmsm, coords*L) uint64."""
    cv = CURVES[curve]
    s = _as_u64(scalars)
    p = _as_u64(points)
    if s.ndim != 3 or s.shape[2] != expo_nlimbs:
        raise ValueError(f"scalars must have shape (nmsm, n, {expo_nlimbs}), got {s.shape}")
    nmsm, n = s.shape[0], s.shape[1]
    if p.size != n * 2 * cv["nlimbs_p"]:
        raise ValueError(f"points must hold n = {n} records of {2 * cv['nlimbs_p']} uint64 words, got {p.size} words")
    mode = _OUT[out]
    res = np.zeros((nmsm, _OUT_COORDS[mode] * cv["nlimbs_p"]), dtype=np.uint64)
    lib().zkb200_msm(cv["id"], nmsm, n, s.ctypes.data if s.size else None, HOST, p.ctypes.data if p.size else None, HOST,
                     expo_nlimbs, int(mont), mode, window, _ptr(res))
    return res


def msm(curve: str, scalars: np.ndarray, points: np.ndarray, mont: bool = True, out: str = "affine",
        window: int = 0, expo_nlimbs: int = 4) -> np.ndarray:
    """sum_i k_i * P_i.  `mont=True` = Curve.msm (Montgomery Fr coefficients), `mont=False` = msmStd."""
    s = _as_u64(scalars).reshape(1, -1, expo_nlimbs)
    return msm_batch(curve, s, points, mont=mont, out=out, window=window, expo_nlimbs=expo_nlimbs)[0]


def msm_std(curve: str, scalars: np.ndarray, points: np.ndarray, **kw) -> np.ndarray:
    return msm(curve, scalars, points, mont=False, **kw)


def msm_device(curve: str, scalars_ptr: int, points_ptr: int, npoints: int, nmsm: int = 1, mont: bool = True,
               out: str = "affine", window: int = 0, expo_nlimbs: int = 4) -> np.ndarray:
    """Same computation with inputs already resident in the current device's memory (raw device pointers,
    e.g. torch.Tensor.data_ptr()); only the result crosses PCIe."""
    cv = CURVES[curve]
    mode = _OUT[out]
    res = np.zeros((nmsm, _OUT_COORDS[mode] * cv["nlimbs_p"]), dtype=np.uint64)
    lib().zkb200_msm(cv["id"], nmsm, npoints, scalars_ptr, DEVICE, points_ptr, DEVICE, expo_nlimbs, int(mont), mode,
                     window, _ptr(res))
    return res


def msm_to_device(curve: str, scalars, points, npoints: int, out_ptr: int, nmsm: int = 1, mont: bool = True, out: str = "xyzz",
                  window: int = 0, resident: bool = True, expo_nlimbs: int = 4) -> None:
    """MSM whose result record(s) stay in DEVICE memory at `out_ptr` (multi-GPU combine without a host bounce).
    `resident`: scalars / points are raw device pointers; else host arrays."""
    cv = CURVES[curve]
    if resident:
        sp, pp, loc = scalars, points, DEVICE
    else:
        s, p = _as_u64(scalars), _as_u64(points)
        if s.size != nmsm * npoints * expo_nlimbs or p.size != npoints * 2 * cv["nlimbs_p"]:
            raise ValueError("scalars / points do not match npoints")
        sp, pp, loc = s.ctypes.data, p.ctypes.data, HOST
    lib().zkb200_msm_ex(cv["id"], nmsm, npoints, sp, loc, pp, loc, expo_nlimbs, int(mont), _OUT[out], window, out_ptr, DEVICE)


def sum_points_device(curve: str, in_ptr: int, k: int, in_repr: str = "xyzz", out: str = "affine") -> np.ndarray:
    """Sum of k group elements that already sit in the current device's memory (e.g. an NCCL all-gather output)."""
    cv = CURVES[curve]
    mode = _OUT[out]
    res = np.zeros(_OUT_COORDS[mode] * cv["nlimbs_p"], dtype=np.uint64)
    lib().zkb200_sum_points_ex(cv["id"], k, in_ptr, DEVICE, _OUT[in_repr], mode, _ptr(res))
    return res


def last_srs_hit() -> bool:
    """True when the most recent MSM on the current device took its points from the resident-copy cache."""
    return bool(lib().zkb200_last_srs_hit())


def release_workspaces() -> None:
    lib().zkb200_release_workspaces()


def set_glv(on: bool) -> None:
    """Switch the endomorphism (GLV) split of the scalars on / off (see zkb200_set_glv: subgroup precondition)."""
    L = lib()
    L.zkb200_set_glv.argtypes = [ctypes.c_int]
    L.zkb200_set_glv.restype = None
    L.zkb200_set_glv(int(bool(on)))


def last_op_ms() -> float:
    """Device time of the kernels of the last conversion / NTT / group FFT call (copies excluded)."""
    L = lib()
    L.zkb200_last_op_ms.restype = ctypes.c_float
    return float(L.zkb200_last_op_ms())


def srs_cache_drop() -> None:
    """Forget every resident point array (the next call over an array uploads it again)."""
    lib().zkb200_srs_cache_drop()


def selftest_field(curve: str, field: str, op: str, a: np.ndarray, b=None, c=None, d=None) -> np.ndarray:
    """Element-wise device field operation on (n, limbs) uint64 arrays (tests only; see zkb200_selftest_field)."""
    a = _as_u64(a)
    arrs = [a] + [None if x is None else _as_u64(x) for x in (b, c, d)]
    for x in arrs[1:]:
        if x is not None and x.shape != a.shape:
            raise ValueError("operand shapes differ")
    out = np.zeros_like(a)
    lib().zkb200_selftest_field(FIELD_IDS[(curve, field)], FIELD_OPS[op], a.shape[0],
                                *[None if x is None else x.ctypes.data for x in arrs], _ptr(out.ravel()))
    return out


def selftest_group(curve: str, op: str, p1: np.ndarray, z1: np.ndarray, p2: np.ndarray, z2: np.ndarray) -> np.ndarray:
    """Element-wise device group operation (tests only; see zkb200_selftest_group) -> (n, 2L) canonical affine."""
    p1, z1, p2, z2 = (_as_u64(x) for x in (p1, z1, p2, z2))
    n = p1.shape[0]
    L = CURVES[curve]["nlimbs_p"]
    if p1.shape != (n, 2 * L) or p2.shape != (n, 2 * L) or z1.shape != (n, L) or z2.shape != (n, L):
        raise ValueError("selftest_group: operand shapes")
    out = np.zeros((n, 2 * L), dtype=np.uint64)
    lib().zkb200_selftest_group(CURVES[curve]["id"], GROUP_OPS[op], n, p1.ctypes.data, z1.ctypes.data, p2.ctypes.data, z2.ctypes.data,
                                _ptr(out.ravel()))
    return out


def ntt(curve: str, m: int, gen: np.ndarray, src: np.ndarray, inverse: bool = False) -> np.ndarray:
    """Fr NTT of 2^m elements ((N, 4) uint64, canonical Montgomery form), the reference's
    <curve>_poly_mont_ntt_forward / _inverse (natural order in and out)."""
    a = _as_u64(src).reshape(-1, 4)
    assert a.shape[0] == 1 << m
    g = _as_u64(gen).ravel()
    out = np.zeros_like(a)
    getattr(lib(), f"{curve}_poly_mont_ntt_{'inverse' if inverse else 'forward'}")(m, _ptr(g), _ptr(a.ravel()), _ptr(out.ravel()))
    return out


def ntt_device(curve: str, m: int, gen: np.ndarray, src_ptr: int, dst_ptr: int, inverse: bool = False) -> None:
    """Same transform on device-resident buffers (raw device pointers, 2^m x 32 bytes each)."""
    g = _as_u64(gen).ravel()
    lib().zkb200_ntt(CURVES[curve]["id"], m, _ptr(g), src_ptr, DEVICE, dst_ptr, DEVICE, int(inverse))


def group_fft(curve: str, m: int, gen: np.ndarray, src: np.ndarray, inverse: bool = False, group: str = "G1_proj") -> np.ndarray:
    """FFT of 2^m group elements ((N, 3 coordinates) uint64) -> normalised points, the reference's
    <curve>_{G1_proj,G1_jac,G2_proj}_fft_forward / _inverse."""
    a = _as_u64(src)
    assert a.shape[0] == 1 << m
    g = _as_u64(gen).ravel()
    out = np.zeros_like(a)
    getattr(lib(), f"{curve}_{group}_fft_{'inverse' if inverse else 'forward'}")(m, _ptr(g), _ptr(a.ravel()), _ptr(out.ravel()))
    return out


def batch_to_affine(curve: str, pts: np.ndarray, repr: str = "proj") -> np.ndarray:
    """(N, 3L) projective / Jacobian points -> (N, 2L) canonical affine (the reference's batchToAffine)."""
    a = _as_u64(pts)
    n = a.shape[0]
    out = np.zeros((n, 2 * CURVES[curve]["nlimbs_p"]), dtype=np.uint64)
    if n:
        getattr(lib(), f"{curve}_G1_{repr}_batch_to_affine")(n, _ptr(a.ravel()), _ptr(out.ravel()))
    return out


def batch_from_affine(curve: str, pts: np.ndarray, repr: str = "proj") -> np.ndarray:
    a = _as_u64(pts)
    n = a.shape[0]
    out = np.zeros((n, 3 * CURVES[curve]["nlimbs_p"]), dtype=np.uint64)
    if n:
        getattr(lib(), f"{curve}_G1_{repr}_batch_from_affine")(n, _ptr(a.ravel()), _ptr(out.ravel()))
    return out


def sum_points(curve: str, pts: np.ndarray, in_repr: str = "proj", out: str = "affine") -> np.ndarray:
    """Sum of k group elements in a reference representation (the multi-GPU combine of partial MSMs)."""
    cv = CURVES[curve]
    a = _as_u64(pts)
    k = a.shape[0] if a.ndim == 2 else 1
    mode = _OUT[out]
    res = np.zeros(_OUT_COORDS[mode] * cv["nlimbs_p"], dtype=np.uint64)
    lib().zkb200_sum_points(cv["id"], k, _ptr(a.ravel()), _OUT[in_repr], mode, _ptr(res))
    return res


def gen_chain(curve: str, n: int, p0: np.ndarray, d: np.ndarray, start: int = 0, device_ptr: Optional[int] = None) -> Optional[np.ndarray]:
    """Synthetic points out[i] = P0 + (start+i)*D (affine Montgomery records), computed on the GPU.
    With `device_ptr` the points are written to that device buffer and None is returned."""
    cv = CURVES[curve]
    p0 = _as_u64(p0).ravel()
    d = _as_u64(d).ravel()
    if device_ptr is not None:
        lib().zkb200_gen_chain(cv["id"], start, n, _ptr(p0), _ptr(d), device_ptr, DEVICE)
        return None
    out = np.zeros((n, 2 * cv["nlimbs_p"]), dtype=np.uint64)
    lib().zkb200_gen_chain(cv["id"], start, n, _ptr(p0), _ptr(d), out.ctypes.data, HOST)
    return out


class ResidentPoints:
    """A point array kept on the device across calls (the SRS of a KZG prover): upload once, commit many times."""

    def __init__(self, curve: str, points: np.ndarray):
        p = _as_u64(points)
        words = 2 * CURVES[curve]["nlimbs_p"]
        if p.ndim != 2 or p.shape[1] != words:
            raise ValueError(f"points must have shape (n, {words}), got {p.shape}")
        self.curve, self.n = curve, p.shape[0]
        self.ptr = lib().zkb200_device_upload(p.ctypes.data, p.nbytes)

    def msm(self, scalars: np.ndarray, mont: bool = True, out: str = "affine", window: int = 0) -> np.ndarray:
        """scalars: (n, 4) or (nmsm, n, 4) host array -> one result per MSM."""
        cv = CURVES[self.curve]
        s = _as_u64(scalars)
        if s.shape[-1] != 4 or s.size % (self.n * 4) != 0:
            raise ValueError(f"scalars must have shape (n, 4) or (nmsm, n, 4) with n = {self.n}, got {s.shape}")
        batch = s.reshape(-1, self.n, 4)
        mode = _OUT[out]
        res = np.zeros((batch.shape[0], _OUT_COORDS[mode] * cv["nlimbs_p"]), dtype=np.uint64)
        lib().zkb200_msm(cv["id"], batch.shape[0], self.n, batch.ctypes.data, HOST, self.ptr, DEVICE, 4, int(mont), mode, window,
                         _ptr(res))
        return res if s.ndim == 3 else res[0]

    def close(self) -> None:
        if self.ptr:
            lib().zkb200_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def launch_count() -> int:
    return int(lib().zkb200_launch_count())


def last_stats() -> dict:
    ms = (ctypes.c_float * 9)()
    c, w, ins = ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong()
    lib().zkb200_last_stats(ms, ctypes.byref(c), ctypes.byref(w), ctypes.byref(ins))
    names = ["h2d_scalars", "recode", "sort", "wait_points", "accumulate", "fixup", "reduce", "tail_d2h", "total"]
    L = lib()
    L.zkb200_last_affine_levels.restype = ctypes.c_int
    return {"phase_ms": dict(zip(names, [float(x) for x in ms])), "window": c.value, "nwindows": w.value,
            "insertions": ins.value, "affine_levels": int(L.zkb200_last_affine_levels()), "srs_hit": bool(L.zkb200_last_srs_hit())}


def imad_peak(kind: int = 0, iters: int = 2000) -> float:
    """Measured 32x32-bit products per second of the whole GPU (see zkb200_imad_peak)."""
    return float(lib().zkb200_imad_peak(kind, iters))
