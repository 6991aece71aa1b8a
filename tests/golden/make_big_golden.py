#!/usr/bin/env python3
"""Generate tests/golden/big_golden.json: the canonical affine results of BASELINE.json's full-size configs,
computed ONCE with the UNMODIFIED reference C (oracle/_ref/libzk_ref.so) in the build container.

    python -m tests.golden.make_big_golden [key ...]        # keys of tests.workloads.CONFIGS; default: all

Method ("large-n oracle strategy"): the canonical affine sum does not depend on how the
index set is cut, so the vectors are cut into T contiguous shards, every shard goes through the reference's own
`<curve>_G1_proj_MSM_{std,mont}_coeff_proj_out` on its own thread, the partial results are added with the
reference's `proj_add` and converted with `proj_to_affine`.  BLS12-381 2^26 uses the std entry point (the mont one
overflows `int` at 2^26, lib/cbits/curves/g1/proj/bn128_G1_proj.c:630).
Inputs: tests.workloads (points = chain, scalars = counter_scalars) -- nothing but (curve, n, form, seed) and the
resulting bytes are stored.  Existing entries are kept; only missing keys are computed.
"""
import json
import os
import sys
import time

import numpy as np

from tests import refs, workloads

T = os.cpu_count() or 1


def one_msm(curve, n, form, seed):
    t0 = time.time()
    pts = refs.chain_points(curve, n, nthreads=T)
    sc = refs.counter_scalars(seed, 0, n)
    t1 = time.time()
    out = refs.ref_msm_threads(curve, sc, pts, mont=(form == "mont"), nthreads=T, use_ref=True)
    print(f"  {curve} n={n} {form} seed={seed}: inputs {t1 - t0:.1f}s, reference MSM on {T} threads {time.time() - t1:.1f}s", flush=True)
    return out.tobytes().hex()


def batch(curve, n, form, seed, nmsm):
    import threading
    L = refs.CURVE_LIMBS[curve]
    pts = refs.chain_points(curve, n)
    lib = refs.ref()
    sym = f"{curve}_G1_proj_MSM_{form}_coeff_affine_out"
    res = [None] * nmsm

    def work(k):
        for m in range(k, nmsm, T):
            sc = refs.counter_scalars(seed + m, 0, n)
            res[m] = refs.call_msm(lib, sym, sc.ravel(), pts.ravel(), 2 * L, n=n).tobytes()

    t0 = time.time()
    ths = [threading.Thread(target=work, args=(k,)) for k in range(T)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    print(f"  {curve} batch {nmsm} x {n}: {time.time() - t0:.1f}s", flush=True)
    return b"".join(res).hex()


def main():
    assert refs.have_ref(), "oracle/_ref/libzk_ref.so missing: run `make -C oracle ref` where /root/reference exists"
    want = sys.argv[1:] or list(workloads.CONFIGS)
    vec = workloads.load_big_golden()
    jobs = []
    for name in want:
        c = workloads.CONFIGS[name]
        if c["nmsm"] > 1:
            jobs.append((c["curve"], 1 << c["logn"], c["form"], c["seed"], c["nmsm"]))
        elif c["weak"]:
            for g in (1, 2, 4, 8):
                jobs.append((c["curve"], g << c["logn"], c["form"], c["seed"], 1))
        else:
            jobs.append((c["curve"], 1 << c["logn"], c["form"], c["seed"], 1))
    for (curve, n, form, seed, nmsm) in jobs:
        key = workloads.golden_key(curve, n, form, seed, nmsm)
        if key in vec:
            print("have", key, flush=True)
            continue
        print("computing", key, flush=True)
        vec[key] = batch(curve, n, form, seed, nmsm) if nmsm > 1 else one_msm(curve, n, form, seed)
        with open(workloads.BIG_GOLDEN, "w") as f:
            json.dump(dict(generator="tests/golden/make_big_golden.py",
                           source="oracle/_ref/libzk_ref.so (unmodified reference C + platform.h shim), sharded over host threads",
                           key="curve:global_n:scalar_form:seed:nmsm -> affine bytes (hex), nmsm results concatenated",
                           vectors=vec), f, indent=1)
    print("wrote", workloads.BIG_GOLDEN)


if __name__ == "__main__":
    main()
