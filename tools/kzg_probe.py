"""Batched-KZG shape on one GPU: where the time of the host-buffer path goes (run under gpurun)."""
import os, sys, time
os.environ.setdefault("ZKB200_TRACE", "1")
sys.path.insert(0, ".")
import numpy as np
import torch
import zikkurat_algebra_b200 as zk
from tests import refs, workloads

c = workloads.CONFIGS["kzg"]
curve, n, nmsm = c["curve"], 1 << c["logn"], int(sys.argv[1]) if len(sys.argv) > 1 else 128
srs = refs.chain_points(curve, n)
sc = workloads.batch_scalars(c["seed"], nmsm, n)
h_sc = torch.from_numpy(sc.view(np.int64)).pin_memory().numpy().view(np.uint64)
d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
d_pts = torch.from_numpy(srs.view(np.int64)).cuda()
def t(fn, reps=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    return (time.perf_counter() - t0) / reps * 1e3, r
ms, r0 = t(lambda: zk.msm_device(curve, d_sc.data_ptr(), d_pts.data_ptr(), n, nmsm=nmsm, mont=True)); print("resident", round(ms, 3), zk.last_stats()["phase_ms"], flush=True)
ms, r1 = t(lambda: zk.msm_batch(curve, sc, srs, mont=True)); print("pageable", round(ms, 3), zk.last_srs_hit(), zk.last_stats()["phase_ms"], flush=True)
ms, r2 = t(lambda: zk.msm_batch(curve, h_sc, srs, mont=True)); print("pinned", round(ms, 3), zk.last_srs_hit(), zk.last_stats()["phase_ms"], flush=True)
assert r0.tobytes() == r1.tobytes() == r2.tobytes()
