"""Group FFT timing (scope row 8f.4): GPU (host buffers) vs the reference C on one core."""
import ctypes, json, sys, time
import numpy as np
sys.path.insert(0, ".")
import zikkurat_algebra_b200 as zk
from tests import pyec, refs
out = {}
for curve, m in (("bn128", 10), ("bn128", 14), ("bls12_381", 10), ("bls12_381", 14)):
    cv = pyec.CURVES[curve]; N = 1 << m
    g0 = 5 if curve == "bn128" else 7
    gen = np.frombuffer(((pow(g0, (cv.r - 1) >> m, cv.r) * cv.Rr) % cv.r).to_bytes(32, "little"), dtype=np.uint64).copy()
    proj = zk.batch_from_affine(curve, refs.chain_points(curve, N), "proj")
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); got = zk.group_fft(curve, m, gen, proj); best = min(best, time.perf_counter() - t0)
    cpu = None
    if m <= 10:
        f = getattr(refs.ref(), f"{curve}_G1_proj_fft_forward"); f.argtypes = [ctypes.c_int, refs.U64P, refs.U64P, refs.U64P]; f.restype = None
        want = np.zeros_like(proj); t0 = time.perf_counter(); f(m, refs.ptr(gen), refs.ptr(proj.ravel()), refs.ptr(want.ravel())); cpu = time.perf_counter() - t0
        assert want.tobytes() == got.tobytes()
    out[f"{curve}_2^{m}"] = dict(gpu_ms=best * 1e3, reference_c_1core_ms=cpu * 1e3 if cpu else None)
    print(curve, m, out[f"{curve}_2^{m}"], flush=True)
json.dump(out, open("gpurun_out/gfft_bench.json", "w"), indent=1)
