"""First-contact GPU script: IMAD throughput probe + phase timings at a few sizes (run under gpurun)."""
import json
import sys
import time

import numpy as np
import torch

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


def main():
    out = {}
    props = torch.cuda.get_device_properties(0)
    out["gpu"] = props.name
    out["sms"] = props.multi_processor_count
    for kind, name in ((0, "lo_hi_chain"), (1, "mad_wide"), (2, "mad_lo")):
        v = zk.imad_peak(kind, 4000)
        out[f"imad_{name}_Gprod_s"] = v / 1e9
        print(name, v / 1e9, "Gproducts/s", flush=True)
    sizes = [("bls12_381", 20), ("bn128", 20), ("bn128", 22), ("bls12_381", 16), ("bn128", 16)]
    if len(sys.argv) > 1:
        sizes = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1:] if ":" in a]
    for curve, logn in sizes:
        n = 1 << logn
        pts = gen(curve, n)
        sc = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda")
        sc[:, 3] &= (1 << 60) - 1
        torch.cuda.synchronize()
        best = None
        for rep in range(4):
            t0 = time.perf_counter()
            r = zk.msm_device(curve, sc.data_ptr(), pts.data_ptr(), n, mont=True, out="affine")
            dt = time.perf_counter() - t0
            st = zk.last_stats()
            if best is None or dt < best[0]:
                best = (dt, st)
        dt, st = best
        print(curve, logn, f"{dt*1e3:.3f} ms", json.dumps(st), flush=True)
        out[f"{curve}_2^{logn}"] = dict(ms=dt * 1e3, **st)
    # batched KZG shape (BASELINE configs[4] per GPU): 32 MSMs x 2^14 points over one shared SRS
    if "--batch" in sys.argv or len(sys.argv) == 1:
        curve, n, nmsm = "bn128", 1 << 14, 32
        pts = gen(curve, n)
        sc = torch.randint(0, 2**62, (nmsm, n, 4), dtype=torch.int64, device="cuda")
        sc[:, :, 3] &= (1 << 60) - 1
        torch.cuda.synchronize()
        best = None
        for rep in range(4):
            t0 = time.perf_counter()
            zk.msm_device(curve, sc.data_ptr(), pts.data_ptr(), n, nmsm=nmsm, mont=True, out="affine")
            dt = time.perf_counter() - t0
            st = zk.last_stats()
            if best is None or dt < best[0]:
                best = (dt, st)
        print("batch32x2^14", curve, f"{best[0]*1e3:.3f} ms", json.dumps(best[1]), flush=True)
        t0 = time.perf_counter()
        for i in range(nmsm):
            zk.msm_device(curve, sc[i].data_ptr(), pts.data_ptr(), n, mont=True, out="affine")
        print("same as 32 single calls:", f"{(time.perf_counter()-t0)*1e3:.3f} ms", flush=True)
        out["batch32x2^14_bn128"] = dict(ms=best[0] * 1e3, **best[1])
    json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
