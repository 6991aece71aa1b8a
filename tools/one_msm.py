"""Exactly two device-resident MSM calls (warm-up + one) for profiler captures.  usage: one_msm.py [curve] [logn]"""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import zikkurat_algebra_b200 as zk
from tests import pyec
curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_381"
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cv = pyec.CURVES[curve]; n = 1 << logn
p0 = np.frombuffer(cv.affine_to_bytes(cv.mul(0x1234567, cv.gen)), dtype=np.uint64).copy()
d = np.frombuffer(cv.affine_to_bytes(cv.mul(0x7654321, cv.gen)), dtype=np.uint64).copy()
pts = torch.empty((n, 2 * cv.nlimbs_p), dtype=torch.int64, device="cuda")
zk.gen_chain(curve, n, p0, d, device_ptr=pts.data_ptr())
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda"); sc[:, 3] &= (1 << 61) - 1
torch.cuda.synchronize()
for rep in range(2):
    out = zk.msm_device(curve, sc.data_ptr(), pts.data_ptr(), n, mont=True, out="affine")[0]
    print(json.dumps(zk.last_stats()), flush=True)
