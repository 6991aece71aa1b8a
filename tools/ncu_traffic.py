#!/usr/bin/env python3
"""DRAM traffic of the bucket-accumulation phase from an ncu launch list -> profiles/r2_traffic.json (read by bench.py).

On the GPU box (one GPU, after the same command has exited 0 without ncu):
    python tools/one_msm.py bls12_381 20 > gpurun_out/plain.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/one_msm.py bls12_381 20
Here:
    python tools/ncu_traffic.py gpurun_out/launches.csv bls12_381 20 [--summary profiles/<name>.txt]
The SECOND MSM of the run is taken (the first one allocates); `traffic` = sum of dram__bytes_read + dram__bytes_write
over every kernel between the sort and the bucket reduction (affine tree, inversion chains, record / plain accumulation,
head fix-up), per MSM.  The capture's per-launch times are cold-cache and serialised: shares, not absolutes.
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PHASE = ("k_aff_", "k_binv_", "k_accumulate", "k_fixup_level")


def parse(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    idx = {h: i for i, h in enumerate(hdr)}
    launch = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        d = launch.setdefault(int(r[idx["ID"]]), {"name": r[idx["Kernel Name"]]})
        d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
    return launch


def main():
    path, curve, logn = sys.argv[1], sys.argv[2], int(sys.argv[3])
    launch = parse(path)
    ids = sorted(launch)
    first = max(i for i in ids if "k_recode" in launch[i]["name"])          # the last MSM of the run
    agg = collections.OrderedDict()
    for i in ids:
        if i < first:
            continue
        L = launch[i]
        name = L["name"].split("(")[0].replace("void ", "").replace("zk::", "")
        a = agg.setdefault(name, dict(n=0, us=0.0, rd=0.0, wr=0.0, fma=0.0))
        a["n"] += 1
        a["us"] += L.get("gpu__time_duration.sum", 0.0) / 1e3
        a["rd"] += L.get("dram__bytes_read.sum", 0.0)
        a["wr"] += L.get("dram__bytes_write.sum", 0.0)
        a["fma"] = max(a["fma"], L.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", 0.0))
    phase = {k: v for k, v in agg.items() if k.startswith(PHASE)}
    traffic = sum(v["rd"] + v["wr"] for v in phase.values())
    levels = sum(1 for k in phase if k.startswith("k_aff_add"))          # distinct instantiations: first / middle / last level
    lines = [f"{k:64s} n={v['n']:3d} t={v['us']:8.1f} us  dram_rd={v['rd'] / 1e6:8.1f} MB dram_wr={v['wr'] / 1e6:8.1f} MB  fmaheavy_max%={v['fma']:5.1f}"
             for k, v in agg.items()]
    total_us = sum(v["us"] for v in agg.values())
    lines.append(f"TOTAL {total_us:.1f} us serialised; accumulation phase: {sum(v['us'] for v in phase.values()):.1f} us, "
                 f"dram {traffic / 1e9:.3f} GB per MSM")
    print("\n".join(lines))
    if "--summary" in sys.argv:
        with open(sys.argv[sys.argv.index("--summary") + 1], "w") as f:
            f.write(f"ncu launch list of tools/one_msm.py {curve} {logn} (second MSM), from {os.path.basename(path)}\n" + "\n".join(lines) + "\n")
    out = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        db = json.load(open(out))
    except (OSError, ValueError):
        db = {"captures": []}
    R = int(os.environ.get("AFFINE_LEVELS", "-1"))
    rec = dict(curve=curve, n=1 << logn, affine_levels=R, dram_bytes=traffic, source=os.path.relpath(path, ROOT),
               kernels={k: dict(launches=v["n"], us=round(v["us"], 1), dram_bytes=v["rd"] + v["wr"]) for k, v in phase.items()})
    db["captures"] = [c for c in db["captures"] if not (c["curve"] == curve and c["n"] == rec["n"] and c["affine_levels"] == R)] + [rec]
    json.dump(db, open(out, "w"), indent=1)
    print("wrote", out, "(set AFFINE_LEVELS=<R of the run> so that bench.py can match it)" if R < 0 else "")


if __name__ == "__main__":
    main()
