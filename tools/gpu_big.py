"""Largest single-GPU case: BLS12-381 2^26 (BASELINE configs[3] on ONE GPU), inputs resident.
Consistency: MSM(all) == MSM(first half) + MSM(second half) (canonical affine bytes)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
import zikkurat_algebra_b200 as zk
from tests import pyec
curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_381"
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 26
cv = pyec.CURVES[curve]; n = 1 << logn
p0 = np.frombuffer(cv.affine_to_bytes(cv.mul(0x1234567, cv.gen)), dtype=np.uint64).copy()
d = np.frombuffer(cv.affine_to_bytes(cv.mul(0x7654321, cv.gen)), dtype=np.uint64).copy()
pts = torch.empty((n, 2 * cv.nlimbs_p), dtype=torch.int64, device="cuda")
t0 = time.perf_counter(); zk.gen_chain(curve, n, p0, d, device_ptr=pts.data_ptr()); print("gen_chain s", time.perf_counter() - t0, flush=True)
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda"); sc[:, 3] &= (1 << 61) - 1
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); whole = zk.msm_device(curve, sc.data_ptr(), pts.data_ptr(), n, mont=True, out="affine")[0]; dt = time.perf_counter() - t0
    print(f"{curve} 2^{logn}: {dt*1e3:.2f} ms", json.dumps(zk.last_stats()), flush=True)
h = n // 2
a = zk.msm_device(curve, sc.data_ptr(), pts.data_ptr(), h, mont=True, out="xyzz")[0]
b = zk.msm_device(curve, sc[h:].data_ptr(), pts[h:].data_ptr(), h, mont=True, out="xyzz")[0]
s = zk.sum_points(curve, np.stack([a, b]), "xyzz", "affine")
print("split consistent:", s.tobytes() == whole.tobytes(), "mem GB", torch.cuda.max_memory_allocated() / 1e9, flush=True)
