"""CPU: the C-ABI library builds, loads and exports every symbol include/pslam_abi.h declares; without a GPU
context creation must fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pslam_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pslam_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from planarslam_b200 import _lib
    L = C.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(L, s), f"{s} declared in pslam_abi.h but not exported"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from planarslam_b200 import _lib
    with pytest.raises(_lib.PslamError) as e:
        _lib.Context(640, 480, 1)
    assert e.value.code == _lib.E_NO_DEVICE


def test_product_does_not_reference_oracle():
    """Nothing under planarslam_b200/ may import, include or link the oracle."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "planarslam_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc", "Makefile")):
                t = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"#include\s+[\"<][^\">]*oracle/|import\s+oracle|from\s+oracle|liboracle", t):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
