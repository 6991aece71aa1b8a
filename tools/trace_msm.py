"""Timeline of one MSM's window groups ($ZKB200_TRACE) and the end-to-end effect of input slices (run under gpurun).
usage: python tools/trace_msm.py [curve] [logn]"""
import os
import sys
import time

os.environ["ZKB200_TRACE"] = "1"
sys.path.insert(0, ".")
import numpy as np
import torch

import zikkurat_algebra_b200 as zk
from tests import refs

curve = sys.argv[1] if len(sys.argv) > 1 else "bls12_381"
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = 1 << logn
L = zk.CURVES[curve]["nlimbs_p"]
p0, d = refs.chain_base(curve)
d_pts = torch.empty((n, 2 * L), dtype=torch.int64, device="cuda")
zk.gen_chain(curve, n, p0, d, device_ptr=d_pts.data_ptr())
sc = refs.counter_scalars(2, 0, n)
d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
pts = d_pts.cpu().numpy().view(np.uint64)
print("== resident", flush=True)
for _ in range(5):
    zk.msm_device(curve, d_sc.data_ptr(), d_pts.data_ptr(), n, mont=True)
h_sc = torch.from_numpy(sc.view(np.int64)).pin_memory().numpy().view(np.uint64)
for slices in ("1", "2"):
    for name, s in (("pageable", sc), ("pinned", h_sc)):
        os.environ["ZKB200_SLICES"] = slices
        for _ in range(3):
            zk.msm(curve, s, pts, mont=True)
        t0 = time.perf_counter()
        for _ in range(10):
            zk.msm(curve, s, pts, mont=True)
        sys.stderr.flush()
        print(f"== slices={slices} {name}: {(time.perf_counter() - t0) * 100:.3f} ms per call (srs hit {zk.last_srs_hit()})", flush=True)
