"""Resident BLS12-381 / BN254 MSM under a list of environment settings; prints best-of-N total and the last trace line.
usage: python tools/trace_env.py curve logn "VAR=val VAR2=val" "..." """
import os, subprocess, sys
curve, logn = sys.argv[1], sys.argv[2]
code = r'''
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import zikkurat_algebra_b200 as zk
from tests import refs
curve, logn = sys.argv[1], int(sys.argv[2]); n = 1 << logn
L = zk.CURVES[curve]["nlimbs_p"]
p0, d = refs.chain_base(curve)
d_pts = torch.empty((n, 2 * L), dtype=torch.int64, device="cuda")
zk.gen_chain(curve, n, p0, d, device_ptr=d_pts.data_ptr())
d_sc = torch.from_numpy(refs.counter_scalars(2, 0, n).view(np.int64)).cuda()
ts = []
for i in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    zk.msm_device(curve, d_sc.data_ptr(), d_pts.data_ptr(), n, mont=True)
    ts.append((time.perf_counter() - t0) * 1e3)
st = zk.last_stats()["phase_ms"]
print("best %.3f median %.3f | acc %.3f reduce %.3f tail %.3f" % (min(ts[2:]), sorted(ts[2:])[5], st["accumulate"], st["reduce"], st["tail_d2h"]))
'''
for setting in sys.argv[3:]:
    env = dict(os.environ)
    for kv in setting.split():
        if "=" in kv:
            k, v = kv.split("=", 1); env[k] = v
    env["ZKB200_TRACE"] = "1"
    r = subprocess.run([sys.executable, "-c", code, curve, logn], env=env, capture_output=True, text=True)
    tr = [l for l in r.stderr.splitlines() if "trace" in l]
    print("##", setting or "(default)", "->", r.stdout.strip(), flush=True)
    if tr: print("   ", tr[-1][len("[zkmsm_b200 trace] "):], flush=True)
    if r.returncode: print(r.stderr[-2000:])
