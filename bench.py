#!/usr/bin/env python3
"""bench.py -- G1 MSM throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--curve bls12_381|bn128] [--logn L] [--impl ours|reference]

A "step" is one complete MSM over synthetic inputs: per GPU 2^L points/scalars (default: BLS12-381,
2^20 -- BASELINE.json configs[1]); with N GPUs the global vectors hold N*2^L elements, sharded
contiguously, one process per GPU (torchrun), each computing a partial MSM, the N partial points
all-gathered over NCCL and summed on rank 0 ("scaling": "weak").

* `value`   points/s with inputs already resident in HBM (zkb200_msm with device pointers).
* `e2e`     the same through the reference-facing C-ABI call with HOST buffers (pinned), H2D of scalars
            and points and D2H of the result inside the timed region.
* `roofline` IMAD-pipe roofline of the bucket-accumulation kernel: achieved =
            n*W insertions x (1360 | 3000) 32x32-bit products / accumulate-kernel time (CUDA events
            on the launching stream, read from the library); peak = the same GPU's carry-chain
            mad.lo.cc/madc.hi.cc throughput measured live by zkb200_imad_peak.
* `cpu_baseline` the reference's own C MSM (oracle/_ref, unmodified sources) on this box's host cores,
            on a bounded sample of the same workload.
`--impl reference` times only that CPU arm and prints the same JSON shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRODUCTS_PER_INSERTION = {"bn128": 1360, "bls12_381": 3000}   # 10 Fp mul x (2L^2 + L), SURVEY.md 8d
# what k_accumulate really executes per insertion: 6 mul + 2 dedicated squarings + 1 fused (a*b + c*d):
# 6(2L^2+L) + 2(L(L+1)/2 + L^2 + L) + (3L^2 + L)
EXECUTED_PER_INSERTION = {"bn128": 6 * 136 + 2 * 108 + 200, "bls12_381": 6 * 300 + 2 * 234 + 444}
# one batched-affine addition of the pre-reduction tree (kernels_aff.cuh): 5 mul + 1 squaring
EXECUTED_PER_AFFINE_ADD = {"bn128": 5 * 136 + 108, "bls12_381": 5 * 300 + 234}
# dram__bytes_read.sum + dram__bytes_write.sum of the bucket-accumulation phase from an `ncu --set full` capture, keyed by
# (curve, log2 n, affine levels); None where no capture exists.
#   R = 0: ONE k_accumulate launch                      profiles/r1_e_ncu_k_accumulate_bls12381_2p20.txt
#   R = 3: all kernels of the phase summed (tree + records)   profiles/r1_g_ncu_accumulate_phase_bls12381_2p20.txt
NCU_TRAFFIC = {("bls12_381", 20, 0): 1.738798e9 + 0.172671e9,
               ("bls12_381", 20, 3): 7.393e9 + 2.057e9}
METRIC = "G1 MSM throughput"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--curve", default="bls12_381", choices=["bls12_381", "bn128"])
    ap.add_argument("--logn", type=int, default=20, help="log2(points per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--pageable", action="store_true", help="also time the host-buffer path with ordinary (pageable) numpy arrays")
    return ap.parse_args()


def workload_name(curve, logn, n_gpus):
    cname = "BLS12-381" if curve == "bls12_381" else "BN254"
    return f"{cname} G1 MSM, 2^{logn} points per GPU x {n_gpus} GPU(s), uniform random Montgomery-form Fr scalars"


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []
        self.marks = {}

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self, name):
        self.marks[name] = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = self.marks.get("t0", 0.0), self.marks.get("t1", float("inf"))
        inside = [ln for (ts, ln) in self.lines if t0 <= ts <= t1 + 0.05]
        for ln in (inside if inside else [ln for (_, ln) in self.lines]):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
def cpu_reference_arm(curve, logn, steps, warmup, quiet=False):
    """The reference's own CPU MSM (oracle/_ref = unmodified reference C; else the pinned port) on all host
    threads, each step a bounded sample of the workload."""
    import numpy as np
    from tests import refs
    refs.build_oracles()
    T = os.cpu_count() or 1
    n = 1 << logn
    per_thread = 1 << 15 if curve == "bls12_381" else 1 << 16   # ~2-3 s of work per thread and step
    n_s = min(n, per_thread * T)
    pts = refs.chain_points(curve, n_s)
    sc = refs.random_scalars(curve, n_s, seed=2)
    kind = "reference" if refs.have_ref() else "port"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        refs.ref_msm_threads(curve, sc, pts, mont=True, nthreads=T, use_ref=(kind == "reference"))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    avg = sum(times) / len(times)
    return dict(value=n_s / avg, unit=UNIT, cores=T, kind=kind, ms_per_step=avg * 1e3,
                sample=f"first {n_s} points/scalars of the workload, {T} contiguous shards on {T} threads "
                       f"({curve}_G1_proj_MSM_mont_coeff_proj_out per shard + proj_add + proj_to_affine)")


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner) write to fd 1, so fd 1
    is pointed at stderr for the whole run and the saved descriptor is used for the result."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    args = parse()
    out = _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    curve, logn = args.curve, args.logn
    n = 1 << logn
    config = {"workload": workload_name(curve, logn, world), "curve": curve, "points_per_gpu": n,
              "global_points": n * world, "entry_point": f"{curve}_G1_proj_MSM_mont_coeff_affine_out",
              "sharding": "contiguous, one process per GPU, partial points all-gathered (NCCL) and summed on rank 0",
              "l2": "inputs larger than L2 (points + scalars + sort pairs > 126 MB), no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference_arm(curve, logn, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        out.write(json.dumps(line) + "\n")
        out.flush()
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from zikkurat_algebra_b200 import build as zkbuild
    if rank == 0:
        zkbuild.build()
    if world > 1:
        dist.barrier()
    import zikkurat_algebra_b200 as zk
    from tests import pyec
    zk.set_device(local_rank)
    cv = pyec.CURVES[curve]
    L = cv.nlimbs_p

    # ---- synthetic inputs, generated on the device: global chain P_i = (s0 + i*s1)*G, this rank's slice ----
    p0 = np.frombuffer(cv.affine_to_bytes(cv.mul(0x1234567, cv.gen)), dtype=np.uint64).copy()
    d = np.frombuffer(cv.affine_to_bytes(cv.mul(0x7654321, cv.gen)), dtype=np.uint64).copy()
    d_pts = torch.empty((n, 2 * L), dtype=torch.int64, device="cuda")
    zk.gen_chain(curve, n, p0, d, start=rank * n, device_ptr=d_pts.data_ptr())
    g = torch.Generator(device="cuda")
    g.manual_seed(2 + rank)
    d_sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    d_sc[:, 3] &= (1 << 61) - 1                      # < 2^253 < r: a valid Montgomery representative
    h_pts = torch.empty((n, 2 * L), dtype=torch.int64, pin_memory=True)
    h_sc = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    h_pts.copy_(d_pts); h_sc.copy_(d_sc)
    torch.cuda.synchronize()
    np_pts = h_pts.numpy().view(np.uint64)
    np_sc = h_sc.numpy().view(np.uint64)
    from zikkurat_algebra_b200.distributed import msm_sharded
    part_words = (2 if world == 1 else 4) * L

    def step_device():
        return msm_sharded(curve, d_sc.data_ptr(), d_pts.data_ptr(), npoints=n, mont=True, resident=True, window=args.window)

    def step_e2e():
        return msm_sharded(curve, np_sc, np_pts, mont=True, resident=False, window=args.window)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, collect_stats=False):
        acc_ms, sort_ms, stats = 0.0, 0.0, None
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = None
        for _ in range(steps):
            res = fn()
            if collect_stats:
                stats = zk.last_stats()
                acc_ms += stats["phase_ms"]["accumulate"]
                sort_ms += stats["phase_ms"]["sort"]
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res, acc_ms, sort_ms, stats

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                      # nvidia-smi needs ~0.1 s to deliver its first sample: start before the warm-up
    for _ in range(max(args.warmup, 3)):
        step_device()
    launches0 = zk.launch_count()
    sampler.mark("t0")
    ms_dev, res_dev, acc_ms, sort_ms, stats = timed(step_device, args.steps, collect_stats=True)
    sampler.mark("t1")
    launches = zk.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e, res_e2e, _, _, _ = timed(step_e2e, args.steps)

    ms_pageable = None
    if args.pageable:
        pg_pts, pg_sc = np.array(np_pts, copy=True), np.array(np_sc, copy=True)   # ordinary malloc'ed memory

        def step_pageable():
            return msm_sharded(curve, pg_sc, pg_pts, mont=True, resident=False, window=args.window)
        for _ in range(2):
            step_pageable()
        ms_pageable, _, _, _, _ = timed(step_pageable, args.steps)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    assert res_dev.tobytes() == res_e2e.tobytes(), "device-resident and host-buffer paths disagree"

    total_points = n * world
    value = total_points * args.steps / (ms_dev * 1e-3)
    e2e_value = total_points * args.steps / (ms_e2e * 1e-3)
    ppi = PRODUCTS_PER_INSERTION[curve]
    t_acc = acc_ms / args.steps * 1e-3
    achieved = stats["insertions"] * ppi / t_acc                     # products/s on rank 0's GPU
    probe = zk.imad_peak(0, 4000)
    props = torch.cuda.get_device_properties(local_rank)
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    # IMAD.WIDE.U32(.X) occupies the FMA-heavy pipe for 4 cycles per warp instruction (ncu: 21.2 % fma issue
    # rate <-> 80.7 % sm__pipe_fmaheavy_cycles_active, profiles/r1_a_ncu_k_accumulate_*.txt):
    # peak = SMs x 4 SMSP x 32 lanes / 4 cycles x SM clock sampled during the run
    peak = props.multi_processor_count * 4 * 32 / 4.0 * sm_mhz * 1e6
    passes = (stats["window"] + 7) // 8
    sort_bytes = stats["insertions"] * 20 * passes
    R = int(stats.get("affine_levels", 0))
    traffic = NCU_TRAFFIC.get((curve, logn, R))
    # products really executed by the phase: with R levels of batched-affine pre-reduction, level r adds (at most)
    # insertions / 2^(r+1) pairs at 5M+1S each and the XYZZ insertion (6M+2S+fused) is left for insertions / 2^R records
    ins = stats["insertions"]
    executed = sum(ins / 2 ** (r + 1) for r in range(R)) * EXECUTED_PER_AFFINE_ADD[curve] + ins / 2 ** R * EXECUTED_PER_INSERTION[curve]
    kernel = "k_accumulate" if R == 0 else (f"bucket accumulation phase: {R} levels of batched-affine pre-reduction "
                                            "(k_aff_prod, inversion chain, k_aff_add) + k_accumulate_rec")
    note = ("achieved/frac use SURVEY.md 8d's algorithmic 10 Fp mul per insertion; the kernel executes fewer products "
            "(fused Y3 reduction, dedicated squarings)")
    if R:
        note += (f"; with {R} affine levels most additions cost 5M+1S instead of 8M+2S, so frac can exceed 1 -- "
                 "frac_of_pipe_peak is the utilisation of the IMAD pipe by the products really executed")
    roofline = {"bound": "imad", "kernel": kernel, "achieved": achieved / 1e9, "peak": peak / 1e9,
                "unit": "Gproducts/s (32x32->64-bit multiply-adds)", "frac": achieved / peak,
                "peak_source": f"IMAD pipe: {props.multi_processor_count} SM x 4 SMSP x 32 lanes / 4 cycles per IMAD.WIDE x "
                               f"{sm_mhz:.0f} MHz (SM clock sampled during the timed region); cross-check: ncu "
                               "sm__pipe_fmaheavy_cycles_active (profiles/), zkb200_imad_peak carry-chain probe on this GPU = "
                               f"{probe / 1e9:.0f} Gproducts/s",
                "probe_gproducts": probe / 1e9,
                "executed": {"products_per_insertion": executed / ins, "affine_levels": R,
                             "gproducts_per_s": executed / t_acc / 1e9,
                             "frac_of_pipe_peak": executed / t_acc / peak,
                             "note": note},
                "per_launch": {"insertions": stats["insertions"], "products_per_insertion": ppi, "window_c": stats["window"],
                               "nwindows": stats["nwindows"], "avg_ms": t_acc * 1e3,
                               "algorithmic_gather_bytes": stats["insertions"] * (2 * L * 8 + 8)},
                "traffic": traffic,
                "secondary_hbm": {"kernel": "radix sort (3 kernels x passes)", "bytes": sort_bytes,
                                  "achieved_gbs": sort_bytes / (sort_ms / args.steps * 1e-3) / 1e9 if sort_ms else None,
                                  "peak_gbs": _measured_peak("hbm_gbs", 6650.0)}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(world * (np_sc.nbytes + np_pts.nbytes)),
                    "d2h_bytes_per_step": int(world * part_words * 8),
                    "host_memory": "pinned (torch pin_memory), passed as plain pointers to the reference-named C symbol path"},
            "e2e_pageable_ms_per_step": (ms_pageable / args.steps) if ms_pageable else None,
            "gpu_launches": int(launches), "roofline": roofline,
            "phase_ms": stats["phase_ms"]}
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_reference_arm(curve, logn, steps=1, warmup=0)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    out.write(json.dumps(line) + "\n")
    out.flush()
    if world > 1:
        dist.destroy_process_group()


def _measured_peak(key, fallback):
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))[key]
    except (OSError, KeyError, ValueError):
        return fallback


if __name__ == "__main__":
    main()
