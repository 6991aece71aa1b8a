"""bench.py --row f1|f2|f3|f4: one JSON line per "next" row of the scope table, same contract as
the MSM line (value = device-resident where the row has such an entry point, e2e = host buffers through the
reference-named symbol, roofline of the dominant kernel from CUDA-event kernel time, cpu_baseline = the reference C on a
bounded sample).  One GPU.

  f1  <curve>_G1_proj_batch_to_affine        2^20 BLS12-381 projective points   (reference: bn128_G1_proj.c:147-166)
  f2  <curve>_poly_mont_ntt_forward          2^22 BLS12-381 Fr elements         (reference: bn128_poly_mont.c:418-525)
  f3  <curve>_G2_proj_MSM_mont_coeff_affine_out  2^18 BLS12-381 G2 points       (reference: bn128_G2_proj.c:498-660)
  f4  <curve>_G1_proj_fft_forward            2^14 BLS12-381 projective points   (reference: bn128_G1_proj.c:678-789)
"""
import ctypes
import json
import os
import time

import numpy as np


def _gen_of(cv, m):
    g0 = 5 if cv.name == "bn128" else 7
    return np.frombuffer(((pow(g0, (cv.r - 1) >> m, cv.r) * cv.Rr) % cv.r).to_bytes(32, "little"), dtype=np.uint64).copy()


def _time(fn, steps, warmup):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps, r


def run(row, steps, warmup, sampler, want_cpu=True):
    import torch

    import zikkurat_algebra_b200 as zk
    from tests import pyec, refs
    curve = "bls12_381"
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p
    props = torch.cuda.get_device_properties(0)
    launches0 = zk.launch_count()
    warmup = max(warmup, 3)
    cpu = None
    extra = {}
    if row == "f1":
        n = 1 << 20
        unit, metric = "points/s", "G1 batch_to_affine throughput"
        aff = refs.chain_points(curve, n, nthreads=os.cpu_count() or 4)
        proj = zk.batch_from_affine(curve, aff, "proj")
        # (Z = 1 records: the work per point -- inversion tree + two multiplications -- does not depend on the value of Z;
        #  non-trivial denominators are what tests/test_msm_gpu.py::test_batch_conversions_next_row checks)
        sampler.mark("t0")
        dt, got = _time(lambda: zk.batch_to_affine(curve, proj, "proj"), steps, warmup)
        sampler.mark("t1")
        assert got.tobytes() == aff.tobytes()
        k_ms = zk.last_op_ms()
        work = n * 5 * 300           # 3 (batch inversion) + 2 (x, y) Fp multiplications per point, 300 products each
        value = e2e = n / dt
        h2d, d2h = proj.nbytes, got.nbytes
        workload = "BLS12-381 G1: 2^20 projective points -> canonical affine, host buffers (the symbol has no device-resident form)"
        if want_cpu and refs.have_ref():
            ns = 1 << 14
            f = getattr(refs.ref(), f"{curve}_G1_proj_batch_to_affine")
            f.argtypes = [ctypes.c_int, refs.U64P, refs.U64P]
            f.restype = None
            o = np.zeros((ns, 2 * L), np.uint64)
            t0 = time.perf_counter(); f(ns, refs.ptr(proj[:ns].ravel()), refs.ptr(o.ravel())); c = time.perf_counter() - t0
            assert o.tobytes() == aff[:ns].tobytes()
            cpu = dict(value=ns / c, unit=unit, cores=1, kind="reference", sample=f"{curve}_G1_proj_batch_to_affine on the first 2^14 points, one core")
    elif row == "f2":
        m = 22
        n = 1 << m
        unit, metric = "elements/s", "Fr NTT throughput"
        gen = _gen_of(cv, m)
        src = refs.counter_scalars(11, 0, n)
        d_src = torch.from_numpy(src.view(np.int64)).cuda()
        d_dst = torch.empty_like(d_src)
        sampler.mark("t0")
        dt_dev, _ = _time(lambda: zk.ntt_device(curve, m, gen, d_src.data_ptr(), d_dst.data_ptr()), steps, warmup)
        sampler.mark("t1")
        k_ms = zk.last_op_ms()
        dt, got = _time(lambda: zk.ntt(curve, m, gen, src), max(2, steps // 2), 1)
        assert d_dst.cpu().numpy().view(np.uint64).tobytes() == got.tobytes()
        work = (n // 2) * m * 136    # butterflies x one Fr multiplication (8 limbs: 2*64 + 8 products)
        value, e2e = n / dt_dev, n / dt
        h2d = d2h = src.nbytes
        workload = "BLS12-381 Fr NTT, 2^22 elements (natural order in and out)"
        extra["hbm_passes"] = (m + 8) // 9
        if want_cpu and refs.have_ref():
            ms_ = 18
            g2 = _gen_of(cv, ms_)
            f = getattr(refs.ref(), f"{curve}_poly_mont_ntt_forward")
            f.argtypes = [ctypes.c_int, refs.U64P, refs.U64P, refs.U64P]
            f.restype = None
            s2 = np.ascontiguousarray(src[:1 << ms_])
            o = np.zeros_like(s2)
            t0 = time.perf_counter(); f(ms_, refs.ptr(g2), refs.ptr(s2.ravel()), refs.ptr(o.ravel())); c = time.perf_counter() - t0
            assert o.tobytes() == zk.ntt(curve, ms_, g2, s2).tobytes()
            cpu = dict(value=(1 << ms_) / c, unit=unit, cores=1, kind="reference", sample=f"{curve}_poly_mont_ntt_forward at 2^18 elements, one core")
    elif row == "f3":
        logn = 18
        n = 1 << logn
        unit, metric = "points/s", "G2 MSM throughput"
        g2c = curve + "_g2"
        L2 = zk.CURVES[g2c]["nlimbs_p"]
        # G2 chain points: generator multiples built with the library's own chain generator from two reference-made points
        lib = refs.ref() if refs.have_ref() else None
        assert lib is not None, "row f3 needs oracle/_ref for the two seed points of the G2 chain"
        from tests.test_msm_gpu import _g2_ref_chain
        seed_pts = _g2_ref_chain(curve, 2)
        d_pts = torch.empty((n, 2 * L2), dtype=torch.int64, device="cuda")
        zk.gen_chain(g2c, n, seed_pts[0], seed_pts[1], device_ptr=d_pts.data_ptr())
        sc = refs.counter_scalars(12, 0, n)
        d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
        pts = d_pts.cpu().numpy().view(np.uint64)
        sampler.mark("t0")
        dt_dev, res = _time(lambda: zk.msm_device(g2c, d_sc.data_ptr(), d_pts.data_ptr(), n, mont=True), steps, warmup)
        sampler.mark("t1")
        st = zk.last_stats()
        k_ms = st["phase_ms"]["accumulate"]
        dt, got = _time(lambda: zk.call_reference_symbol(f"{curve}_G2_proj_MSM_mont_coeff_affine_out", sc, pts), max(2, steps // 2), 2)
        assert got.tobytes() == res[0].tobytes()
        work = st["insertions"] * 10 * 3 * 300      # 10 Fp2 multiplications per insertion x 3 Fp multiplications (Karatsuba) x 300
        value, e2e = n / dt_dev, n / dt
        h2d, d2h = sc.nbytes, got.nbytes
        workload = "BLS12-381 G2 MSM, 2^18 points, Montgomery scalars"
        extra.update(window_c=st["window"], nwindows=st["nwindows"], insertions=st["insertions"], srs_cache_hit=st["srs_hit"])
        if want_cpu:
            ns = 1 << 12
            t0 = time.perf_counter()
            want = refs.call_msm(lib, f"{curve}_G2_proj_MSM_mont_coeff_affine_out", sc[:ns].ravel(), np.ascontiguousarray(pts[:ns]).ravel(), 2 * L2, n=ns)
            c = time.perf_counter() - t0
            assert want.tobytes() == zk.call_reference_symbol(f"{curve}_G2_proj_MSM_mont_coeff_affine_out", sc[:ns], np.ascontiguousarray(pts[:ns])).tobytes()
            cpu = dict(value=ns / c, unit=unit, cores=1, kind="reference", sample=f"{curve}_G2_proj_MSM_mont_coeff_affine_out on the first 2^12 points, one core")
    elif row == "f4":
        m = 14
        n = 1 << m
        unit, metric = "points/s", "G1 group FFT throughput"
        gen = _gen_of(cv, m)
        proj = zk.batch_from_affine(curve, refs.chain_points(curve, n), "proj")
        sampler.mark("t0")
        dt, got = _time(lambda: zk.group_fft(curve, m, gen, proj), steps, warmup)
        sampler.mark("t1")
        k_ms = zk.last_op_ms()
        # a butterfly = one 255-bit scalar multiplication (4-bit fixed window: 255 doublings + 64 additions) + 2 additions
        work = (n // 2) * m * (255 * 9 + 66 * 14) * 300
        value = e2e = n / dt
        h2d = d2h = proj.nbytes
        workload = "BLS12-381 G1 group FFT, 2^14 projective points, host buffers (the symbol has no device-resident form)"
        if want_cpu and refs.have_ref():
            ms_ = 10
            g2 = _gen_of(cv, ms_)
            f = getattr(refs.ref(), f"{curve}_G1_proj_fft_forward")
            f.argtypes = [ctypes.c_int, refs.U64P, refs.U64P, refs.U64P]
            f.restype = None
            p2 = np.ascontiguousarray(proj[:1 << ms_])
            o = np.zeros_like(p2)
            t0 = time.perf_counter(); f(ms_, refs.ptr(g2), refs.ptr(p2.ravel()), refs.ptr(o.ravel())); c = time.perf_counter() - t0
            assert o.tobytes() == zk.group_fft(curve, ms_, g2, p2).tobytes()
            cpu = dict(value=(1 << ms_) / c, unit=unit, cores=1, kind="reference", sample=f"{curve}_G1_proj_fft_forward at 2^10 points, one core")
    else:
        raise SystemExit(f"unknown row {row}")
    clocks = sampler.stop()
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    peak = props.multi_processor_count * 32 * sm_mhz * 1e6
    achieved = work / (k_ms * 1e-3) if k_ms else None
    line = {"metric": metric, "row": row, "value": value, "unit": unit, "n_gpus": 1, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * (n / value), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": dict(workload=workload, curve=curve, **extra), "clocks": clocks,
            "parity": "bit-identical to the reference C on the cpu_baseline sample; device-resident and host-buffer results identical",
            "e2e": {"value": e2e, "unit": unit, "ms_per_step": 1e3 * n / e2e, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "host_memory": "ordinary (pageable) numpy arrays through the reference-named symbol"},
            "gpu_launches": int(zk.launch_count() - launches0),
            "roofline": {"bound": "imad", "kernel": "all kernels of the call (CUDA events around them, copies excluded)",
                         "achieved": achieved / 1e9 if achieved else None, "peak": peak / 1e9,
                         "unit": "Gproducts/s (32x32->64-bit multiply-adds)", "frac": achieved / peak if achieved else None,
                         "per_launch": {"algorithmic_products": work, "avg_ms": k_ms}, "traffic": None}}
    if cpu:
        line["cpu_baseline"] = cpu
    return line
