"""Phase breakdown of the host-buffer (end-to-end) call: pinned inputs, $ZKB200_SLICES / $ZKB200_SLICE0 variants."""
import sys, time, json, os
import numpy as np, torch
sys.path.insert(0, ".")
import zikkurat_algebra_b200 as zk
from tests import pyec
curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_381"
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cv = pyec.CURVES[curve]; n = 1 << logn; L = cv.nlimbs_p
p0 = np.frombuffer(cv.affine_to_bytes(cv.mul(0x1234567, cv.gen)), dtype=np.uint64).copy()
d = np.frombuffer(cv.affine_to_bytes(cv.mul(0x7654321, cv.gen)), dtype=np.uint64).copy()
d_pts = torch.empty((n, 2 * L), dtype=torch.int64, device="cuda")
zk.gen_chain(curve, n, p0, d, device_ptr=d_pts.data_ptr())
d_sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda"); d_sc[:, 3] &= (1 << 61) - 1
h_pts = torch.empty((n, 2 * L), dtype=torch.int64, pin_memory=True); h_sc = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
h_pts.copy_(d_pts); h_sc.copy_(d_sc); torch.cuda.synchronize()
np_pts = h_pts.numpy().view(np.uint64); np_sc = h_sc.numpy().view(np.uint64)
# raw copy rate
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); d_pts.copy_(h_pts, non_blocking=True); d_sc.copy_(h_sc, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("plain H2D of both arrays: %.3f ms" % e0.elapsed_time(e1), flush=True)
for cfg in sys.argv[3:] or ["1", "2"]:
    parts = cfg.split(":")
    os.environ["ZKB200_SLICES"] = parts[0]
    if len(parts) > 1: os.environ["ZKB200_SLICE0"] = parts[1]
    else: os.environ.pop("ZKB200_SLICE0", None)
    best, st = 1e9, None
    for rep in range(6):
        t0 = time.perf_counter(); zk.msm(curve, np_sc, np_pts, mont=True, out="affine"); dt = time.perf_counter() - t0
        if rep and dt < best: best, st = dt, zk.last_stats()
    print(cfg, "%.3f ms" % (best * 1e3), json.dumps({k: round(v, 3) for k, v in st["phase_ms"].items()}), "R", st["affine_levels"], flush=True)
