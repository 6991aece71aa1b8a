#!/usr/bin/env python3
"""Generate tests/golden/msm_golden.json from the UNMODIFIED reference C (oracle/_ref/libzk_ref.so).

Run in the build container (needs /root/reference to have been compiled by oracle/Makefile):
    python -m tests.golden.make_golden
Inputs are fully determined by (curve, n, seed): points = tests.refs.chain_points(curve, n),
scalars = tests.refs.random_scalars(curve, n, seed); only the expected canonical affine bytes are stored.
The reference has no golden vectors of its own for the MSM path, so these pin
the oracle restatement and the CUDA path to the reference's actual output.
"""
import json
import os

import numpy as np

from tests import refs

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "msm_golden.json")
SIZES = [1, 2, 3, 31, 32, 33, 123, 1024, 1 << 14, 1 << 16]


def main():
    lib = refs.ref()
    vectors = []
    for curve in ("bn128", "bls12_381"):
        L = refs.CURVE_LIMBS[curve]
        for n in SIZES:
            pts = refs.chain_points(curve, n)
            for form in ("std", "mont"):
                seed = 1000 + n % 997 + (0 if form == "std" else 1)
                sc = refs.random_scalars(curve, n, seed=seed, reduce=(form == "mont"))
                got = refs.call_msm(lib, f"{curve}_G1_proj_MSM_{form}_coeff_affine_out", sc.ravel(), pts.ravel(), 2 * L, n=n)
                jac = refs.call_msm(lib, f"{curve}_G1_jac_MSM_{form}_coeff_affine_out", sc.ravel(), pts.ravel(), 2 * L, n=n)
                assert got.tobytes() == jac.tobytes()
                vectors.append(dict(curve=curve, n=n, form=form, seed=seed, affine_hex=got.tobytes().hex()))
                print(curve, n, form, got.tobytes().hex()[:32], flush=True)
    with open(OUT, "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py", source="oracle/_ref/libzk_ref.so (unmodified reference C + platform.h shim)",
                       vectors=vectors), f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
