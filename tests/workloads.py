"""BASELINE.json's configs as reproducible synthetic workloads (shared by bench.py, the GPU tests and the golden
generator).  TEST / BENCH INFRASTRUCTURE ONLY.

A workload is a pure function of (curve, global n, scalar form, seed): points are the global chain
P_i = (s0 + i*s1)*G (tests.refs.chain_points on the CPU; zikkurat_algebra_b200.gen_chain on the GPU -- the two
are compared in tests/test_msm_gpu.py), scalars are tests.refs.counter_scalars(seed, i).  Every rank of a
multi-GPU run regenerates exactly its slice; tests/golden/make_big_golden.py computes the canonical affine result
of the WHOLE workload once with the unmodified reference C (oracle/_ref) and stores it in
tests/golden/big_golden.json, so that bench.py can assert the bytes of a 2/4/8-GPU run on a box where
/root/reference does not exist.
"""
from __future__ import annotations

import json
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
BIG_GOLDEN = os.path.join(HERE, "golden", "big_golden.json")

# name -> BASELINE.json config.  `weak`: points per GPU are fixed (global n = N * 2^logn), else the global n is fixed.
CONFIGS = {
    # configs[1]: BLS12-381 G1 MSM 2^20 on 1 x B200 (the headline; at N > 1 the driver's scaling run is weak: 2^20 per GPU)
    "bls20": dict(curve="bls12_381", logn=20, form="mont", seed=2, weak=True, nmsm=1,
                  title="BLS12-381 G1 MSM, 2^20 points per GPU"),
    # configs[0] scaled to the GPU: BN254 2^20 per GPU
    "bn20": dict(curve="bn128", logn=20, form="mont", seed=3, weak=True, nmsm=1,
                 title="BN254 G1 MSM, 2^20 points per GPU"),
    # configs[2]: BN254 G1 MSM 2^24 sharded across 2/4/8 B200 (strong scaling)
    "bn24": dict(curve="bn128", logn=24, form="mont", seed=3, weak=False, nmsm=1,
                 title="BN254 G1 MSM, 2^24 points in total"),
    # north_star's target size for both curves: 2^24 points on 8 GPUs (BN254 is config 3; this is the BLS12-381 twin)
    "bls24": dict(curve="bls12_381", logn=24, form="mont", seed=2, weak=False, nmsm=1,
                  title="BLS12-381 G1 MSM, 2^24 points in total"),
    # configs[3]: BLS12-381 G1 MSM 2^26 sharded across 8 B200; standard-form scalars (the reference's mont entry
    # point overflows `int` in malloc(8*expo_nlimbs*npoints) at 2^26, lib/cbits/curves/g1/proj/bn128_G1_proj.c:630)
    "bls26": dict(curve="bls12_381", logn=26, form="std", seed=4, weak=False, nmsm=1,
                  title="BLS12-381 G1 MSM, 2^26 points in total"),
    # configs[4]: batched KZG, 256 independent BN254 MSMs of 2^14 points over one shared SRS, whole MSMs dealt to the GPUs
    "kzg": dict(curve="bn128", logn=14, form="mont", seed=5, weak=False, nmsm=256,
                title="batched KZG commit: 256 x BN254 G1 MSM of 2^14 points over one shared SRS"),
}


def golden_key(curve: str, n_global: int, form: str, seed: int, nmsm: int = 1) -> str:
    return f"{curve}:{n_global}:{form}:{seed}:{nmsm}"


def load_big_golden() -> dict:
    try:
        with open(BIG_GOLDEN) as f:
            return json.load(f)["vectors"]
    except (OSError, ValueError, KeyError):
        return {}


def golden_bytes(curve: str, n_global: int, form: str, seed: int, nmsm: int = 1) -> Optional[bytes]:
    v = load_big_golden().get(golden_key(curve, n_global, form, seed, nmsm))
    return bytes.fromhex(v) if v else None


def batch_scalars(seed: int, nmsm: int, n: int, first: int = 0):
    """Scalars of MSMs [first, first + nmsm) of a batched workload: MSM m uses counter_scalars(seed + m, 0, n)."""
    import numpy as np
    from . import refs
    return np.stack([refs.counter_scalars(seed + first + m, 0, n) for m in range(nmsm)])
