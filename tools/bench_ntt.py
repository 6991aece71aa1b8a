"""Fr NTT timing on the GPU box (scope row 8f.2): device-resident, host-buffer and the reference C on one core."""
import ctypes
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import zikkurat_algebra_b200 as zk
from tests import pyec, refs


def gen_of(cv, m):
    g0 = 5 if cv.name == "bn128" else 7
    return pow(g0, (cv.r - 1) >> m, cv.r)


def main():
    out = {}
    for curve, m in (("bn128", 20), ("bls12_381", 20), ("bn128", 24)):
        cv = pyec.CURVES[curve]
        N = 1 << m
        gen = np.frombuffer(((gen_of(cv, m) * cv.Rr) % cv.r).to_bytes(32, "little"), dtype=np.uint64).copy()
        src = refs.random_scalars(curve, N, seed=m)
        d_src = torch.from_numpy(src.view(np.int64)).cuda()
        d_dst = torch.empty_like(d_src)
        torch.cuda.synchronize()
        best_dev = best_host = 1e9
        for _ in range(5):
            t0 = time.perf_counter(); zk.ntt_device(curve, m, gen, d_src.data_ptr(), d_dst.data_ptr()); best_dev = min(best_dev, time.perf_counter() - t0)
        for _ in range(3):
            t0 = time.perf_counter(); res = zk.ntt(curve, m, gen, src); best_host = min(best_host, time.perf_counter() - t0)
        assert d_dst.cpu().numpy().view(np.uint64).tobytes() == res.tobytes()
        cpu = None
        if m <= 20 and refs.have_ref():
            f = getattr(refs.ref(), f"{curve}_poly_mont_ntt_forward")
            f.argtypes = [ctypes.c_int, refs.U64P, refs.U64P, refs.U64P]
            f.restype = None
            o = np.zeros_like(src)
            t0 = time.perf_counter(); f(m, refs.ptr(gen), refs.ptr(src.ravel()), refs.ptr(o.ravel())); cpu = time.perf_counter() - t0
            assert o.tobytes() == res.tobytes()
        passes = (m + 8) // 9
        gbs = passes * 64 * N / best_dev / 1e9
        row = dict(device_ms=best_dev * 1e3, host_buffers_ms=best_host * 1e3, reference_c_1core_ms=cpu * 1e3 if cpu else None,
                   passes=passes, algorithmic_GBps=gbs)
        out[f"{curve}_2^{m}"] = row
        print(curve, m, json.dumps(row), flush=True)
    json.dump(out, open("gpurun_out/ntt_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
