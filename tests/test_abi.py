"""CPU: the C-ABI library loads and exports every symbol include/pttools_b200.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pttools_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ptt_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_all_symbols():
    from pttools_b200 import build, _lib
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"


def test_no_gpu_calls_fail_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pttools_b200 import engine, _lib
    with pytest.raises(_lib.PttError):
        engine.sweep_shells_bag([0.5], [0.1])


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pttools_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "hostemu" not in txt, f
