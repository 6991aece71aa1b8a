"""Resident-input MSM time over sizes (run under gpurun): best of 8 calls per (curve, log2 n)."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
import zikkurat_algebra_b200 as zk
from tests import refs
out = {}
for curve in ("bn128", "bls12_381"):
    L = zk.CURVES[curve]["nlimbs_p"]
    p0, d = refs.chain_base(curve)
    for logn in [int(a) for a in sys.argv[1:]] or [10, 12, 14, 16, 18, 19, 20, 21, 22]:
        n = 1 << logn
        d_pts = torch.empty((n, 2 * L), dtype=torch.int64, device="cuda")
        zk.gen_chain(curve, n, p0, d, device_ptr=d_pts.data_ptr())
        d_sc = torch.from_numpy(refs.counter_scalars(2, 0, n).view(np.int64)).cuda()
        ts = []
        for i in range(10):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            zk.msm_device(curve, d_sc.data_ptr(), d_pts.data_ptr(), n, mont=True)
            ts.append((time.perf_counter() - t0) * 1e3)
        st = zk.last_stats()
        ph = st["phase_ms"]
        out[f"{curve}_2^{logn}"] = dict(ms=min(ts[2:]), c=st["window"], W=st["nwindows"], R=st["affine_levels"], **{k: round(v, 3) for k, v in ph.items()})
        print(curve, logn, f"{min(ts[2:]):.3f} ms  c={st['window']} W={st['nwindows']} R={st['affine_levels']}  sort {ph['recode']+ph['sort']:.3f} acc {ph['accumulate']:.3f} reduce {ph['reduce']:.3f} tail {ph['tail_d2h']:.3f}", flush=True)
        del d_pts, d_sc
json.dump(out, open("gpurun_out/r2_size_sweep.json", "w"), indent=1)
