"""Small MSMs through every accumulation mode, for compute-sanitizer runs (memcheck / initcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import zikkurat_algebra_b200 as zk
from tests import refs
modes = [dict(ZKB200_AFFINE="0"), dict(ZKB200_AFFINE="3"), dict(ZKB200_AFFINE="5", ZKB200_STAGGER="0"),
         dict(ZKB200_AFFINE="2", ZKB200_STAGGER="4"), dict(ZKB200_AFFINE="3", ZKB200_STAGGER="0", ZKB200_AFF_GROUPS="1"),
         dict(ZKB200_AFFINE="1", ZKB200_SLICES="3"), dict(ZKB200_AFFINE="3", ZKB200_AFF_NEXT="1"),
         dict(ZKB200_AFFINE="2", ZKB200_RED2D="0"), dict(ZKB200_AFFINE="0", ZKB200_WINDOW="9")]
for curve in ("bn128", "bls12_381"):
    pts_all = refs.chain_points(curve, 3001)
    for n in (37, 1000, 3001):
        pts, sc = pts_all[:n], refs.random_scalars(curve, n, seed=n, reduce=True)
        ref = None
        for m in modes:
            for k in ("ZKB200_AFFINE", "ZKB200_STAGGER", "ZKB200_AFF_GROUPS", "ZKB200_SLICES", "ZKB200_AFF_NEXT", "ZKB200_RED2D", "ZKB200_WINDOW"):
                os.environ.pop(k, None)
            os.environ.update(m)
            got = zk.call_reference_symbol(f"{curve}_G1_proj_MSM_mont_coeff_affine_out", sc, pts).tobytes()
            ref = ref or got
            assert got == ref, (curve, n, m)
        print(curve, n, "ok", flush=True)
print("done")
