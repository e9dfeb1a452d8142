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


def test_ladder_key_plan_without_a_device():
    """fl_rotsum_rotations is pure host logic: the doubling keys come first, then the extra multiples of the hoisted groups."""
    import ctypes as C
    from fhe_linformer_b200 import load_library
    lib = load_library()
    lib.fl_rotsum_rotations.restype = C.c_int
    lib.fl_rotsum_rotations.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]

    def plan(steps, stride):
        out = (C.c_int * 64)()
        n = lib.fl_rotsum_rotations(steps, stride, out, 64)
        return [out[i] for i in range(n)]

    r = plan(7, 128)                                   # matmulRE's ladder: 3 + 4 doubling steps
    assert r[:7] == [128 << i for i in range(7)]
    assert set(r) == {128 * t for t in range(1, 8)} | {1024 * t for t in range(1, 16)}
    r = plan(6, 1)                                     # matmulCR: 3 + 3
    assert set(r) == set(range(1, 8)) | {8 * t for t in range(1, 8)}
    assert plan(1, -128) == [-128] and plan(0, 5) == []
    r = plan(5, -1)                                    # repeat(., 32): 2 + 3, negative stride
    assert set(r) == {-1, -2, -3} | {-4 * t for t in range(1, 8)}
