"""The C-ABI library loads without a GPU, exports every symbol include/fl_ckks.h declares, and refuses
to create a context when no CUDA device is present (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(fl_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from fhe_linformer_b200 import LIB_PATH, load_library
    assert os.path.exists(LIB_PATH), "build with __graft_entry__.build()"
    lib = load_library()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fhe_linformer_b200 import Engine
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        Engine(logN=10, L=6, dnum=3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fhe_linformer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "ckks_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
