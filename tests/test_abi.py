"""The C-ABI libraries load and export every symbol include/lpmp_b200.h declares (no compute calls; CPU only)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lpmp_b200.h")).read()
    names = sorted(set(re.findall(r"LPMP_FN\((\w+)\)\s*\(", text)) - {"name"})
    assert len(names) >= 40
    return names


def test_binding_table_matches_header():
    from lp_mp_b200._capi import SIGNATURES
    assert sorted(SIGNATURES) == declared_symbols()


def test_cuda_library_exports_every_symbol():
    import lp_mp_b200
    path = lp_mp_b200.library_path()
    if not os.path.exists(path):
        pytest.skip("liblpmp_b200.so not built (run __graft_entry__.build())")
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, "lpmp_" + name), name


def test_oracle_library_exports_every_symbol(oracle_api):
    for name in declared_symbols():
        assert hasattr(oracle_api.lib, "lpmp_oracle_" + name), name


def test_product_has_no_cpu_fallback_and_never_names_the_oracle():
    """the package must fail loudly without its CUDA library and must not reference oracle/"""
    pkg = os.path.join(ROOT, "lp_mp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liblpmp_oracle" not in text and "lpmp_oracle_" not in text, f
    import lp_mp_b200
    from lp_mp_b200 import LpmpError
    import torch
    if not torch.cuda.is_available():
        if os.path.exists(lp_mp_b200.library_path()):
            with pytest.raises(LpmpError):   # no device: lpmp_create reports a CUDA error, nothing falls back
                lp_mp_b200.LP()
