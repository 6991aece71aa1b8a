"""Time the affine pre-reduction ($ZKB200_AFFINE = R levels) against the plain XYZZ accumulation, inputs resident.
usage: python tools/bench_affine.py [curve] [logn] [R list]   -> JSON lines in gpurun_out/affine_<curve>_<logn>.json"""
import sys, time, json, os
import numpy as np, torch
sys.path.insert(0, ".")
import zikkurat_algebra_b200 as zk
from tests import pyec
curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_381"
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
Rs = [x for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else ["0", "1", "2", "3", "4", "5"]   # "a" = library default
cv = pyec.CURVES[curve]; n = 1 << logn
p0 = np.frombuffer(cv.affine_to_bytes(cv.mul(0x1234567, cv.gen)), dtype=np.uint64).copy()
d = np.frombuffer(cv.affine_to_bytes(cv.mul(0x7654321, cv.gen)), dtype=np.uint64).copy()
pts = torch.empty((n, 2 * cv.nlimbs_p), dtype=torch.int64, device="cuda")
zk.gen_chain(curve, n, p0, d, device_ptr=pts.data_ptr())
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda"); sc[:, 3] &= (1 << 61) - 1
torch.cuda.synchronize()
res = []
ref = None
for R in Rs:
    os.environ.pop("ZKB200_AFFINE", None)
    if R != "a": os.environ["ZKB200_AFFINE"] = R
    best, stats = 1e9, None
    for rep in range(6):
        t0 = time.perf_counter(); out = zk.msm_device(curve, sc.data_ptr(), pts.data_ptr(), n, mont=True, out="affine")[0]; dt = time.perf_counter() - t0
        if rep and dt < best: best, stats = dt, zk.last_stats()
    if ref is None: ref = out.tobytes()
    line = {"curve": curve, "logn": logn, "R": R, "ms": round(best * 1e3, 3), "same_bytes": out.tobytes() == ref, "stats": stats}
    print(json.dumps(line), flush=True); res.append(line)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/affine_{curve}_{logn}.json", "w"), indent=1)
