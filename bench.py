#!/usr/bin/env python3
"""bench.py -- G1 MSM throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config bls20|bn20|bn24|bls26|kzg] [--impl ours|reference]
                    [--in-library-devices] [--no-cpu-baseline] [--window C] [--row f1|f2|f3|f4]
    (aliases: --scaling strong --global-logn 24 --curve bn128  ==  --config bn24, and so on)

A "step" is one complete pass of the hot path over one batch of synthetic input (tests/workloads.py):

  config   BASELINE.json configs[i]                                       scaling at N GPUs
  bls20    [1] BLS12-381 G1 MSM 2^20 (DEFAULT; the headline)               weak: 2^20 points per GPU
  bn20     [0] BN254 G1 MSM (2^20 on the GPU)                              weak
  bn24     [2] BN254 G1 MSM 2^24 sharded across 2/4/8 GPUs                 strong: 2^24 points in total
  bls24    (north_star's target size, BLS12-381 twin of [2])               strong: 2^24 points in total
  bls26    [3] BLS12-381 G1 MSM 2^26 sharded across 8 GPUs (std scalars)   strong
  kzg      [4] 256 x BN254 G1 MSM of 2^14 over one shared SRS              strong: whole MSMs dealt to the GPUs

One process per GPU (torchrun), contiguous shards, each rank computes a partial point that stays on its GPU, one NCCL
all-gather of the N records, rank 0 adds them.  Inputs are a pure function of the global index, so rank 0 compares the
result bytes with the golden answer of the unmodified reference C (tests/golden/big_golden.json) OUTSIDE the timed
region: `"parity": "golden-match"`.

* `value`    points/s, inputs already resident in HBM (device pointers).
* `e2e`      the same through the host-buffer C-ABI path with ORDINARY (pageable) host arrays, which is what the reference's
             callers pass (mallocForeignPtrBytes, lib/src/ZK/Algebra/Class/Flat.hs:186-194): steady state, i.e. the point
             array is the same host array in every step and is served by the library's resident-copy cache, so a step
             moves the scalars in and the result out.  `e2e.first_call_ms` = the same call on a cold cache (scalars AND
             points cross PCIe), `e2e.pinned_ms_per_step` = steady state from pinned memory.
* `roofline` IMAD-pipe roofline of the bucket-accumulation phase.
* `cpu_baseline` the reference's own C MSM (oracle/_ref, unmodified sources) on this box's host cores.
`--impl reference` times only that CPU arm.  `--in-library-devices` runs ONE process that hands the whole host arrays
to the reference-named symbol with ZKB200_DEVICES = all N GPUs (the path an unchanged Haskell caller gets).
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
METRIC = "G1 MSM throughput"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=["bls20", "bn20", "bn24", "bls24", "bls26", "kzg"])
    ap.add_argument("--curve", default=None, choices=["bls12_381", "bn128"])
    ap.add_argument("--logn", type=int, default=None, help="log2(points per GPU), weak scaling")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--global-logn", type=int, default=None, help="log2(points in total), strong scaling")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--row", default=None, choices=["f1", "f2", "f3", "f4"],
                    help="one of the 'next' rows of the scope table instead of the MSM (tools/bench_rows.py), one GPU")
    ap.add_argument("--in-library-devices", action="store_true",
                    help="one process, the library shards the host arrays over --gpus devices itself (ZKB200_DEVICES path)")
    a = ap.parse_args()
    from tests import workloads
    if a.config is None:
        # aliases: (--curve, --logn) weak, or (--scaling strong, --global-logn, --curve)
        if a.scaling == "strong" or a.global_logn is not None:
            want = (a.curve or "bn128", a.global_logn or 24)
            a.config = {("bn128", 24): "bn24", ("bls12_381", 24): "bls24", ("bls12_381", 26): "bls26"}.get(want)
            if a.config is None:
                a.custom = dict(curve=want[0], logn=want[1], form="mont", seed=3 if want[0] == "bn128" else 2, weak=False, nmsm=1,
                                title=f"{want[0]} G1 MSM, 2^{want[1]} points in total")
        else:
            curve = a.curve or "bls12_381"
            logn = a.logn or 20
            a.config = {("bls12_381", 20): "bls20", ("bn128", 20): "bn20"}.get((curve, logn))
            if a.config is None:
                a.custom = dict(curve=curve, logn=logn, form="mont", seed=3 if curve == "bn128" else 2, weak=True, nmsm=1,
                                title=f"{curve} G1 MSM, 2^{logn} points per GPU")
    a.cfg = dict(workloads.CONFIGS[a.config]) if a.config else a.custom
    return a


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
def entry_point(cfg):
    return f"{cfg['curve']}_G1_proj_MSM_{cfg['form']}_coeff_affine_out"


def global_points(cfg, world):
    return (world << cfg["logn"]) if cfg["weak"] else (1 << cfg["logn"])


def describe(cfg, world):
    n_glob = global_points(cfg, world)
    d = {"workload": f"{cfg['title']}{' x ' + str(world) + ' GPU(s)' if cfg['weak'] else ''}, uniform random {'Montgomery' if cfg['form'] == 'mont' else 'standard'}-form Fr scalars",
         "curve": cfg["curve"], "global_points": n_glob, "scalar_form": cfg["form"], "seed": cfg["seed"],
         "entry_point": entry_point(cfg),
         "inputs": "points = chain (s0 + i*s1)*G, scalars = splitmix64(seed, 4i+j) < 2^253: tests/workloads.py",
         "l2": "inputs larger than L2 (points + scalars + sort pairs > 126 MB), no explicit flush"}
    if cfg["nmsm"] > 1:
        d["nmsm"] = cfg["nmsm"]
        d["points_per_msm"] = 1 << cfg["logn"]
        d["sharding"] = "whole MSMs dealt to the GPUs, SRS replicated and resident, result records all-gathered on the devices (NCCL)"
        d["l2"] = "scalars + sort pairs of a rank's share are larger than L2 up to 4 GPUs; no explicit flush"
    else:
        d["points_per_gpu"] = n_glob // world
        d["sharding"] = "contiguous, one process per GPU, XYZZ partial points all-gathered on the devices (NCCL) and summed on rank 0"
    return d


def cpu_reference_arm(cfg, steps, warmup):
    """The reference's own CPU MSM (oracle/_ref = unmodified reference C; else the pinned port) on all host threads.
    bls20 / bn20: the WHOLE per-GPU workload; the big configs: a bounded sample of it (stated)."""
    import numpy as np
    from tests import refs, workloads
    refs.build_oracles()
    T = os.cpu_count() or 1
    curve, form, seed = cfg["curve"], cfg["form"], cfg["seed"]
    kind = "reference" if refs.have_ref() else "port"
    n = 1 << cfg["logn"]
    times = []
    if cfg["nmsm"] > 1:
        nm = min(cfg["nmsm"], 4 * T)
        pts = refs.chain_points(curve, n)
        scs = workloads.batch_scalars(seed, nm, n)
        lib = refs.ref() if kind == "reference" else refs.oracle()
        sym = ("" if kind == "reference" else "zko_") + entry_point(cfg)
        L = refs.CURVE_LIMBS[curve]

        def one_step():
            def work(k):
                for m in range(k, nm, T):
                    refs.call_msm(lib, sym, scs[m].ravel(), pts.ravel(), 2 * L, n=n)
            ths = [threading.Thread(target=work, args=(k,)) for k in range(T)]
            [t.start() for t in ths]
            [t.join() for t in ths]
        n_s = nm * n
        sample = f"{nm} of the {cfg['nmsm']} MSMs (2^{cfg['logn']} points each, shared SRS), one {sym} call per MSM, {T} threads"
    else:
        n_s = min(n, 1 << 20 if cfg["weak"] else 1 << 22)
        pts = refs.chain_points(curve, n_s, nthreads=T)
        sc = refs.counter_scalars(seed, 0, n_s)

        def one_step():
            refs.ref_msm_threads(curve, sc, pts, mont=(form == "mont"), nthreads=T, use_ref=(kind == "reference"))
        whole = "the whole per-GPU workload" if n_s == n else f"the first 2^{n_s.bit_length() - 1} of the 2^{cfg['logn']} points/scalars"
        sample = (f"{whole}: {T} contiguous shards on {T} threads ({curve}_G1_proj_MSM_{form}_coeff_proj_out per shard + "
                  "proj_add + proj_to_affine)")
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one_step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    avg = sum(times) / len(times)
    return dict(value=n_s / avg, unit=UNIT, cores=T, kind=kind, ms_per_step=avg * 1e3, sample=sample)


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner) write to fd 1, so fd 1
    is pointed at stderr for the whole run and the saved descriptor is used for the result."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def _measured_peak(key, fallback):
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))[key]
    except (OSError, KeyError, ValueError):
        return fallback


def _ncu_traffic(curve, n_per_gpu, R):
    """DRAM bytes of the accumulation phase from the committed ncu capture of the CURRENT code
    (profiles/r2_traffic.json, written by tools/ncu_traffic.py), or None when no capture matches."""
    try:
        recs = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["captures"]
    except (OSError, ValueError, KeyError):
        return None, None
    for r in recs:
        if r["curve"] == curve and r["n"] == n_per_gpu and r["affine_levels"] == R:
            return r["dram_bytes"], r.get("source")
    return None, None


def main():
    args = parse()
    cfg = args.cfg
    out = _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    inlib = args.in_library_devices
    n_units = args.gpus if inlib else world     # GPUs the workload is spread over
    config = describe(cfg, n_units)
    scaling = "weak" if cfg["weak"] else "strong"

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference_arm(cfg, args.steps, args.warmup)
        config["reference_sample"] = cb["sample"]
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": n_units,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
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
    if args.row:
        if rank != 0:
            return
        from zikkurat_algebra_b200 import build as zkbuild
        zkbuild.build()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_rows
        sampler = ClockSampler(local_rank)
        sampler.start()
        line = bench_rows.run(args.row, args.steps, args.warmup, sampler, want_cpu=not args.no_cpu_baseline)
        out.write(json.dumps(line) + "\n")
        out.flush()
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from zikkurat_algebra_b200 import build as zkbuild
    if rank == 0:
        zkbuild.build()
    if world > 1:
        dist.barrier()
    import zikkurat_algebra_b200 as zk
    from tests import refs, workloads
    from zikkurat_algebra_b200.distributed import batch_range, msm_batch_dealt, msm_sharded, shard_range
    zk.set_device(local_rank)
    curve, form, seed, nmsm = cfg["curve"], cfg["form"], cfg["seed"], cfg["nmsm"]
    mont = form == "mont"
    L = zk.CURVES[curve]["nlimbs_p"]
    n_glob = global_points(cfg, n_units)
    batch = nmsm > 1

    # ---- this rank's share of the synthetic workload ----------------------------------------------------------
    p0, d = refs.chain_base(curve)
    if batch:
        n = n_glob                                     # every rank holds the whole SRS
        lo_p, m_lo, m_hi = 0, *batch_range(nmsm, world, rank)
        np_sc_pg = workloads.batch_scalars(seed, m_hi - m_lo, n, first=m_lo)          # (nmsm_mine, n, 4), ordinary memory
    elif inlib:
        n, lo_p = n_glob, 0
        np_sc_pg = refs.counter_scalars(seed, 0, n)
    else:
        lo_p, hi_p = shard_range(n_glob, world, rank)
        n = hi_p - lo_p
        np_sc_pg = refs.counter_scalars(seed, lo_p, n)
    d_pts = torch.empty((n, 2 * L), dtype=torch.int64, device="cuda")
    zk.gen_chain(curve, n, p0, d, start=lo_p, device_ptr=d_pts.data_ptr())
    np_pts_pg = np.empty((n, 2 * L), dtype=np.uint64)                                  # ordinary (pageable) memory
    torch.from_numpy(np_pts_pg.view(np.int64)).copy_(d_pts)
    h_pts = torch.from_numpy(np_pts_pg.view(np.int64)).pin_memory()
    h_sc = torch.from_numpy(np_sc_pg.view(np.int64)).pin_memory()
    d_sc = h_sc.cuda()
    torch.cuda.synchronize()
    np_pts_pin, np_sc_pin = h_pts.numpy().view(np.uint64), h_sc.numpy().view(np.uint64)

    if inlib:
        zk.set_devices(list(range(args.gpus)))

    def step_device():
        if batch:
            return msm_batch_dealt(curve, d_sc.data_ptr(), d_pts.data_ptr(), n, nmsm, mont=mont, resident=True, window=args.window)
        return msm_sharded(curve, d_sc.data_ptr(), d_pts.data_ptr(), npoints=n, mont=mont, resident=True, window=args.window)

    def make_host_step(sc, pts):
        def step():
            if (inlib or world == 1) and not batch:
                # exactly the call the reference's FFI makes
                return zk.call_reference_symbol(entry_point(cfg), sc, pts, npoints=n)
            if inlib:
                return zk.msm_batch(curve, sc, pts, mont=mont, out="affine", window=args.window)
            if batch:
                return msm_batch_dealt(curve, sc, pts, n, nmsm, mont=mont, resident=False, window=args.window)
            return msm_sharded(curve, sc, pts, npoints=n, mont=mont, resident=False, window=args.window)
        return step

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

    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    ms_dev = res_dev = stats = None
    acc_ms = sort_ms = 0.0
    launches = 0
    if rank == 0:
        sampler.start()                      # nvidia-smi needs ~0.1 s to deliver its first sample: start before the warm-up
    if not inlib:
        for _ in range(warm):
            step_device()
        launches0 = zk.launch_count()
        sampler.mark("t0")
        ms_dev, res_dev, acc_ms, sort_ms, stats = timed(step_device, args.steps, collect_stats=True)
        sampler.mark("t1")
        launches = zk.launch_count() - launches0

    # ---- end to end: ordinary host arrays; first call on a cold cache, then steady state --------------------
    step_pg = make_host_step(np_sc_pg, np_pts_pg)
    step_pg()                                # work arrays of the host-buffer path allocated (not what "first call" is about)
    zk.srs_cache_drop()                      # cold: no resident copy of the points
    ms_first, res_first, _, _, _ = timed(step_pg, 1)
    for _ in range(max(warm - 1, 2)):
        step_pg()
    if inlib:
        launches0 = zk.launch_count()
        sampler.mark("t0")
    ms_e2e, res_e2e, _, _, st_e2e = timed(step_pg, args.steps, collect_stats=True)
    if inlib:
        sampler.mark("t1")
        launches = zk.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    srs_hit = bool(st_e2e and st_e2e.get("srs_hit"))
    step_pin = make_host_step(np_sc_pin, np_pts_pin)
    for _ in range(2):
        step_pin()
    ms_pin, res_pin, _, _, _ = timed(step_pin, args.steps)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- parity (outside every timed region) -------------------------------------------------------------------
    results = {"first_call": res_first, "e2e": res_e2e, "pinned": res_pin}
    if res_dev is not None:
        results["resident"] = res_dev
    blobs = {k: np.ascontiguousarray(v).tobytes() for k, v in results.items()}
    ref_blob = blobs["e2e"]
    for k, v in blobs.items():
        assert v == ref_blob, f"the {k} path disagrees with the steady-state host-buffer path"
    golden = workloads.golden_bytes(curve, n_glob, form, seed, nmsm)
    if golden is None:
        parity = "no-golden (paths agree with each other)"
    else:
        assert ref_blob == golden, (f"PARITY FAILURE: result differs from the reference C's golden bytes for "
                                    f"{workloads.golden_key(curve, n_glob, form, seed, nmsm)}")
        parity = "golden-match"

    total_points = n_glob * nmsm
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": n_units, "steps": args.steps, "warmup": warm, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config, "clocks": clocks,
            "parity": parity}
    e2e_value = total_points * args.steps / (ms_e2e * 1e-3)
    sc_bytes = int(np_sc_pg.nbytes) * (1 if inlib else world)
    pt_bytes = int(np_pts_pg.nbytes) * (1 if inlib else world)
    res_bytes = int(len(ref_blob)) if (batch or n_units == 1) else 0
    line["e2e"] = {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                   "h2d_bytes_per_step": sc_bytes + (0 if srs_hit else pt_bytes),
                   "d2h_bytes_per_step": res_bytes if res_bytes else int(2 * L * 8),
                   "host_memory": "ordinary (pageable) numpy arrays passed as plain pointers"
                                  + (f" to {entry_point(cfg)}" if (inlib or (world == 1 and not batch)) else " to zkb200_msm_ex"),
                   "steady_state": "same host point array in every step: served by the library's resident-copy cache"
                                   if srs_hit else "points re-sent every step (cache off or too small)",
                   "srs_cache_hit": srs_hit,
                   "first_call_ms": ms_first, "first_call_h2d_bytes": sc_bytes + pt_bytes,
                   "pinned_ms_per_step": ms_pin / args.steps}
    if inlib:
        line.update({"value": e2e_value, "ms_per_step": ms_e2e / args.steps, "mode": "in-library-devices",
                     "note": "one process; value == e2e.value (host buffers only on this path)", "gpu_launches": int(launches)})
        config["sharding"] = f"inside the library: ZKB200_DEVICES = {args.gpus} GPUs, one host thread per device"
        out.write(json.dumps(line) + "\n")
        out.flush()
        return

    value = total_points * args.steps / (ms_dev * 1e-3)
    ppi = PRODUCTS_PER_INSERTION[curve]
    t_acc = acc_ms / args.steps * 1e-3
    achieved = stats["insertions"] * ppi / t_acc                     # products/s on rank 0's GPU
    probes = {name: zk.imad_peak(kind, 4000) / 1e9 for kind, name in
              ((0, "carry_chain_mad_lo_cc_madc_hi"), (1, "mad_wide_no_carry"), (3, "mul_wide_only"), (2, "mad_lo_32bit"))}
    props = torch.cuda.get_device_properties(local_rank)
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    # A 32x32->64-bit product is ONE IMAD.WIDE.U32 (or one IMAD.HI): it holds the FMA-heavy pipe for 4 cycles per warp
    # instruction whatever its form -- with carry (.X), without, product only (tools/imad_forms.cu, profiles/r2_imad_forms.json;
    # ncu: 21 % fma issue rate <-> 81 % sm__pipe_fmaheavy_cycles_active).  peak = SMs x 4 SMSP x 32 lanes / 4 cycles x SM clock.
    peak = props.multi_processor_count * 4 * 32 / 4.0 * sm_mhz * 1e6
    passes = (stats["window"] + 7) // 8
    sort_bytes = stats["insertions"] * 20 * passes
    R = int(stats.get("affine_levels", 0))
    traffic, traffic_src = _ncu_traffic(curve, n, R)
    ins = stats["insertions"]
    executed = sum(ins / 2 ** (r + 1) for r in range(R)) * EXECUTED_PER_AFFINE_ADD[curve] + ins / 2 ** R * EXECUTED_PER_INSERTION[curve]
    kernel = "k_accumulate" if R == 0 else (f"bucket accumulation phase: {R} levels of batched-affine pre-reduction "
                                            "(k_aff_prod, inversion chain, k_aff_add) + k_accumulate_rec")
    roofline = {"bound": "imad", "kernel": kernel, "achieved": achieved / 1e9, "peak": peak / 1e9,
                "unit": "Gproducts/s (32x32->64-bit multiply-adds)", "frac": achieved / peak,
                "frac_executed": executed / t_acc / peak,
                "frac_note": "frac = SURVEY.md 8d's ALGORITHMIC 10 Fp mul (x 2L^2+L products) per insertion / time / peak: above 1 "
                             "means multiplications were saved (batched-affine additions cost 5M+1S, fused Y3 reduction, dedicated "
                             "squarings), not that the pipe ran over; frac_executed = products really issued / time / peak is the "
                             "pipe utilisation",
                "peak_source": f"IMAD pipe: {props.multi_processor_count} SM x 32 IMAD.WIDE per clock x {sm_mhz:.0f} MHz (SM clock sampled "
                               "during the timed region); every 64-bit-product form measures the same on this GPU (imad_probe_gproducts: "
                               "carry chains, mad.wide, mul.wide all within 5 % of each other, 32-bit mad.lo twice that)",
                "imad_probe_gproducts": probes,
                "executed": {"products_per_insertion": executed / ins, "affine_levels": R,
                             "gproducts_per_s": executed / t_acc / 1e9},
                "per_launch": {"insertions": stats["insertions"], "products_per_insertion": ppi, "window_c": stats["window"],
                               "nwindows": stats["nwindows"], "avg_ms": t_acc * 1e3,
                               "algorithmic_gather_bytes": stats["insertions"] * (2 * L * 8 + 8)},
                "traffic": traffic, "traffic_source": traffic_src,
                "secondary_hbm": {"kernel": "radix sort (3 kernels x passes)", "bytes": sort_bytes,
                                  "achieved_gbs": sort_bytes / (sort_ms / args.steps * 1e-3) / 1e9 if sort_ms else None,
                                  "peak_gbs": _measured_peak("hbm_gbs", 6650.0)}}
    line.update({"value": value, "ms_per_step": ms_dev / args.steps, "gpu_launches": int(launches), "roofline": roofline,
                 "phase_ms": stats["phase_ms"]})
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_reference_arm(cfg, steps=1, warmup=0)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    out.write(json.dumps(line) + "\n")
    out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
