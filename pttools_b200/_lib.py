"""ctypes binding of libpttools_b200.so (the C ABI declared in include/pttools_b200.h).

There is NO CPU fallback: if the CUDA library has not been built, or no CUDA device is present when a compute
call is made, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpttools_b200.so")

MEM_HOST, MEM_DEVICE = 0, 1
ST_OK, ST_NO_SOLUTION, ST_INTEGRATION_FAILED, ST_ROOT_NOT_CONVERGED, ST_INVALID_INPUT, ST_SOLVER_FAILED = range(6)

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int32)


class Eos(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("css2", C.c_double), ("csb2", C.c_double),
                ("a_s", C.c_double), ("a_b", C.c_double), ("V_s", C.c_double), ("V_b", C.c_double),
                ("T_ref", C.c_double)]


class A2Plan(C.Structure):
    _fields_ = [("n_z", C.c_int32), ("n_seg", C.c_int32), ("seg", C.c_void_p), ("z", C.c_void_p),
                ("frac", C.c_void_p), ("nxi", C.c_int32), ("phase_rule", C.c_int32), ("xi_re", C.c_void_p),
                ("z_exact_max", C.c_double), ("force_direct", C.c_int32), ("reserved", C.c_int32)]


class SdvPlan(C.Structure):
    _fields_ = [("n_q", C.c_int32), ("n_z", C.c_int32), ("nt", C.c_int32), ("z_tile", C.c_int32),
                ("qT", C.c_void_p), ("z", C.c_void_p), ("zterm", C.c_void_p), ("T", C.c_void_p),
                ("LT", C.c_void_p), ("W", C.c_void_p), ("inv_dlog", C.c_double), ("smem_bytes", C.c_int32),
                ("reserved", C.c_int32), ("omega_ab", C.c_void_p)]


class GwPlan(C.Structure):
    _fields_ = [("n_zl", C.c_int32), ("n_y", C.c_int32), ("n_int", C.c_int32), ("y_tile", C.c_int32),
                ("z_lookup", C.c_void_p), ("y", C.c_void_p), ("yterm", C.c_void_p), ("L1", C.c_void_p),
                ("L2", C.c_void_p), ("Z1", C.c_void_p), ("Z2", C.c_void_p), ("G", C.c_void_p),
                ("prefactor", C.c_double), ("smem_bytes", C.c_int32), ("reserved", C.c_int32)]


class GwCells(C.Structure):
    _fields_ = [("cap", C.c_int32), ("reserved", C.c_int32), ("ncell", C.c_void_p), ("meta", C.c_void_p),
                ("coef", C.c_void_p)]


# name -> (restype, argtypes); the test-suite checks that every symbol of include/pttools_b200.h is here and exported
SIGNATURES = {
    "ptt_last_error": (C.c_char_p, []),
    "ptt_version": (C.c_int, []),
    "ptt_profile_row_capacity": (C.c_int, [C.c_int]),
    "ptt_integrate_batch": (C.c_int, [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.POINTER(Eos)] + [C.c_void_p] * 4 +
                            [C.c_int, C.c_void_p]),
    "ptt_bag_alpha_n_max": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ptt_bag_find_alpha_plus": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ptt_bag_profiles": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_void_p]),
    "ptt_sweep_shells_bag": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ptt_sin_transform_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptt_sin_transform_unit_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptt_a2_batch": (C.c_int, [C.POINTER(A2Plan), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "ptt_a2_work_stride": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "ptt_spec_den_v_batch": (C.c_int, [C.POINTER(SdvPlan), C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_void_p]),
    "ptt_spec_den_v_group": (C.c_int, [C.POINTER(SdvPlan), C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ptt_spec_den_gw_batch": (C.c_int, [C.POINTER(GwPlan), C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptt_gw_cells_capacity": (C.c_int, [C.c_int, C.c_int, C.c_double]),
    "ptt_gw_cells_coef_capacity": (C.c_int, [C.c_int]),
    "ptt_gw_cells_build": (C.c_int, [C.POINTER(GwPlan), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptt_gw_cells_table_size": (C.c_int64, [C.c_int, C.c_int]),
    "ptt_spec_den_gw_cells_batch": (C.c_int, [C.POINTER(GwPlan), C.POINTER(GwCells), C.c_int, C.c_void_p, C.c_int64,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p]),
    "ptt_pow_spec_batch": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ptt_generic_row_capacity": (C.c_int, [C.c_int]),
    "ptt_sweep_shells_generic": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int, C.c_void_p]),
    "ptt_generic_rescue": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "ptt_shell_wall_values": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptt_profile_integrals_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptt_fp64_peak_probe": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    "ptt_fp64_pipe_probe": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p]),
}

_lib = None


class PttError(RuntimeError):
    pass


def load():
    """Load the CUDA library (does not touch the GPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m pttools_b200.build` (nvcc, sm_100a). "
                "pttools_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ptt_last_error()
        raise PttError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise PttError("pttools_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch
