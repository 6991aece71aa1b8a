"""Host-buffer (pageable / pinned) MSM timing under staging settings (run under gpurun).  usage: e2e_probe.py [curve] [logn]"""
import os, subprocess, sys
code = r'''
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import zikkurat_algebra_b200 as zk
from tests import refs
curve, logn = sys.argv[1], int(sys.argv[2]); n = 1 << logn
L = zk.CURVES[curve]["nlimbs_p"]
p0, d = refs.chain_base(curve)
pts = zk.gen_chain(curve, n, p0, d)
sc = refs.counter_scalars(2, 0, n)
sym = f"{curve}_G1_proj_MSM_mont_coeff_affine_out"
h_sc = torch.from_numpy(sc.view(np.int64)).pin_memory().numpy().view(np.uint64)
def t(s):
    for _ in range(3): zk.call_reference_symbol(sym, s, pts)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); zk.call_reference_symbol(sym, s, pts); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), sorted(ts)[5]
print("pageable best %.3f median %.3f | pinned best %.3f median %.3f" % (*t(sc), *t(h_sc)))
'''
for setting in sys.argv[3:] or [""]:
    env = dict(os.environ)
    for kv in setting.split():
        k, v = kv.split("=", 1); env[k] = v
    r = subprocess.run([sys.executable, "-c", code, sys.argv[1], sys.argv[2]], env=env, capture_output=True, text=True)
    print("##", setting or "(default)", "->", r.stdout.strip(), r.stderr[-300:] if r.returncode else "", flush=True)
