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


def test_argument_checks_need_no_gpu():
    """Return codes and ptt_last_error work without a device: an empty batch is valid whatever the pointers, a negative
    count or a null pointer with n > 0 is PTT_E_ARG with a message."""
    import ctypes as C
    from pttools_b200 import _lib
    lib = _lib.load()
    null = C.c_void_p(0)
    assert lib.ptt_bag_alpha_n_max(0, null, null, null, _lib.MEM_DEVICE, null) == 0
    assert lib.ptt_pow_spec_batch(0, 10, null, null, 10, null, null) == 0
    assert lib.ptt_profile_integrals_batch(0, null, null, null, 16, null, null, null, null, null) == 0
    rc = lib.ptt_bag_alpha_n_max(-1, null, null, null, _lib.MEM_DEVICE, null)
    assert rc < 0 and b"ptt_bag_alpha_n_max" in lib.ptt_last_error()
    rc = lib.ptt_bag_alpha_n_max(3, null, null, null, _lib.MEM_DEVICE, null)
    assert rc < 0
    with pytest.raises(_lib.PttError):
        _lib.check(rc, "ptt_bag_alpha_n_max")
    assert lib.ptt_profile_row_capacity(2000) > 2 * 2000 and lib.ptt_generic_row_capacity(5000) == 2 * 5000 + 4
    cap = lib.ptt_gw_cells_capacity(10000, 10000, 1251.3)
    assert cap >= 1252 + 3 and cap % 4 == 0 and lib.ptt_gw_cells_coef_capacity(cap) >= 3 * cap
