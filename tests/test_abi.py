"""CPU: the C-ABI library loads and exports every symbol include/excel_b200.h declares; the ctypes table
matches the header; the product path refuses to run without CUDA (no fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "excel_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(excel_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header_symbols():
    import __graft_entry__
    __graft_entry__.build()
    from excel_b200 import _lib
    L = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/excel_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in excel_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == syms
    assert L.excel_version() >= 1
    assert L.excel_last_error() is not None


def test_no_cpu_fallback():
    from excel_b200.par import PAR
    from excel_b200 import clip
    with pytest.raises(RuntimeError):
        PAR([1, 2, 4, 8, 12, 24], 2)(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError):
        clip.clip_feature_surgery(torch.zeros(1, 5, 8), torch.zeros(3, 8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "excel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    assert "oracle" not in open(os.path.join(ROOT, "include", "excel_b200.h")).read().split("*/")[-1]
