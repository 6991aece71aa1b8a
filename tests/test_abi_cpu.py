"""The C-ABI library loads without a GPU and exports every symbol include/zk_msm_b200.h declares."""
import ctypes
import os
import re

import zikkurat_algebra_b200 as zk
from zikkurat_algebra_b200 import build as zkbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_declared_symbols():
    path = zkbuild.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "zk_msm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(\w+)\s*\(", header)) - {"defined"}
    declared = {d for d in declared if d.startswith(("bn128_", "bls12_381_", "zkb200_"))}
    assert len(declared) >= 26
    for name in declared:
        assert hasattr(lib, name), f"missing export {name}"
    assert declared == set(zk.REFERENCE_SYMBOLS) | set(zk.EXTENSION_SYMBOLS) | set(zk.CONVERT_SYMBOLS) | set(zk.NTT_SYMBOLS) | set(zk.G2_SYMBOLS) | set(zk.GFFT_SYMBOLS) | set(zk.EXTRA_SYMBOLS)


def test_no_oracle_in_product():
    """the product library must not link or reference the CPU oracles"""
    data = open(zk.lib_path(), "rb").read()
    assert b"zko_" not in data and b"libzk_ref" not in data and b"libzk_oracle" not in data
    for root, _, files in os.walk(os.path.join(ROOT, "zikkurat_algebra_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.lower() or f == "__init__.py" and False, f"{f} mentions the oracle"


def test_version_string():
    assert b"sm_100a" in zk.lib().zkb200_version()
