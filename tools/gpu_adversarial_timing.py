"""Timing of adversarial scalar distributions (robustness of the chunked accumulation / fix-up tree)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
import zikkurat_algebra_b200 as zk
from tests import pyec

def gen(curve, n):
    cv = pyec.CURVES[curve]
    p0 = np.frombuffer(cv.affine_to_bytes(cv.mul(0x1234567, cv.gen)), dtype=np.uint64).copy()
    d = np.frombuffer(cv.affine_to_bytes(cv.mul(0x7654321, cv.gen)), dtype=np.uint64).copy()
    buf = torch.empty((n, 2 * cv.nlimbs_p), dtype=torch.int64, device="cuda")
    zk.gen_chain(curve, n, p0, d, device_ptr=buf.data_ptr())
    return buf

curve = sys.argv[1] if len(sys.argv) > 1 else "bn128"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
pts = gen(curve, n)
rnd = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda"); rnd[:, 3] &= (1 << 60) - 1
same_pt = pts[:1].repeat(n, 1).contiguous()      # every point equal: each batched addition is a doubling
cases = {
    "uniform": rnd,
    "all_equal": rnd[:1].repeat(n, 1).contiguous(),
    "all_zero": torch.zeros((n, 4), dtype=torch.int64, device="cuda"),
    "64bit_scalars": torch.cat([rnd[:, :1], torch.zeros((n, 3), dtype=torch.int64, device="cuda")], 1).contiguous(),
    "16_distinct": rnd[:16].repeat(n // 16, 1).contiguous(),
    "one_hot_bits": (torch.ones((n, 4), dtype=torch.int64, device="cuda") << (torch.arange(n, device="cuda") % 60).unsqueeze(1)).contiguous(),
}
runs = [(name, sc, pts) for name, sc in cases.items()] + [("uniform_same_point", rnd, same_pt), ("all_equal_same_point", cases["all_equal"], same_pt)]
for name, sc, P in runs:
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); r = zk.msm_device(curve, sc.data_ptr(), P.data_ptr(), n, mont=False, out="affine"); best = min(best, time.perf_counter() - t0)
    st = zk.last_stats()["phase_ms"]
    print(f"{name:16s} {best*1e3:8.3f} ms  acc={st['accumulate']:.3f} fix={st['fixup']:.3f} red={st['reduce']:.3f} tail={st['tail_d2h']:.3f} sort={st['sort']:.3f}", flush=True)
