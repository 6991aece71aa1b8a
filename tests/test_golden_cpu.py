"""Oracle restatement vs the committed golden vectors (generated from the unmodified reference C)."""
import json
import os

import pytest

from tests import refs

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "msm_golden.json")))["vectors"]


@pytest.mark.parametrize("v", [g for g in GOLDEN if g["n"] <= 1 << 14], ids=lambda g: f"{g['curve']}-{g['n']}-{g['form']}")
def test_restatement_matches_golden(v):
    curve, n, form = v["curve"], v["n"], v["form"]
    L = refs.CURVE_LIMBS[curve]
    pts = refs.chain_points(curve, n)
    sc = refs.random_scalars(curve, n, seed=v["seed"], reduce=(form == "mont"))
    got = refs.call_msm(refs.oracle(), f"zko_{curve}_G1_proj_MSM_{form}_coeff_affine_out", sc.ravel(), pts.ravel(), 2 * L, n=n)
    assert got.tobytes().hex() == v["affine_hex"]
