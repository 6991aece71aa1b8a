"""G2 MSM timing (scope row 8f.3), inputs resident on the device."""
import ctypes, json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import zikkurat_algebra_b200 as zk
from tests import refs
out = {}
for curve, logn in (("bn128", 16), ("bn128", 20), ("bls12_381", 16), ("bls12_381", 20)):
    lib = refs.ref()
    nl = refs.CURVE_LIMBS[curve]; W = 4 * nl; g2 = curve + "_g2"; n = 1 << logn
    gen = (ctypes.c_uint64 * W).in_dll(lib, f"{curve}_G2_affine_gen_G2")
    G = np.frombuffer(bytes(gen), dtype=np.uint64).copy()
    D = refs.call3(lib, f"{curve}_G2_affine_add", G, G, W)
    pts = torch.empty((n, W), dtype=torch.int64, device="cuda")
    zk.gen_chain(g2, n, G, D, device_ptr=pts.data_ptr())
    sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda"); sc[:, 3] &= (1 << 61) - 1
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); zk.msm_device(g2, sc.data_ptr(), pts.data_ptr(), n, mont=True, out="affine"); best = min(best, time.perf_counter() - t0)
    st = zk.last_stats()
    cpu = None
    if logn <= 16:
        m = 1 << 12
        hp = pts[:m].cpu().numpy().view(np.uint64); hs = sc[:m].cpu().numpy().view(np.uint64)
        t0 = time.perf_counter(); refs.call_msm(lib, f"{curve}_G2_proj_MSM_mont_coeff_affine_out", hs.ravel(), hp.ravel(), W, n=m); cpu = (time.perf_counter() - t0)
    row = dict(affine_levels=st.get("affine_levels"), ms=best * 1e3, points_per_s=n / best, phase_ms=st["phase_ms"], window=st["window"], nwindows=st["nwindows"],
               reference_c_2p12_1core_ms=cpu * 1e3 if cpu else None)
    out[f"{g2}_2^{logn}"] = row
    print(g2, logn, json.dumps(row), flush=True)
import os
json.dump(out, open("gpurun_out/g2_bench%s.json" % ("_R" + os.environ["ZKB200_AFFINE"] if "ZKB200_AFFINE" in os.environ else ""), "w"), indent=1)
