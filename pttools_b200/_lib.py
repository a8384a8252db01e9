"""ctypes binding of libpttools_b200.so (the C ABI declared in include/pttools_b200.h).

There is NO CPU fallback: if the CUDA library has not been built, or no CUDA device is present when a compute
call is made, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTT_B200_LIB") or os.path.join(HERE, "libpttools_b200.so")   # override: instrumented builds (tools/)

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


class SweepPlan(C.Structure):
    _fields_ = [("a2", A2Plan), ("n_pass", C.c_int32), ("has_y_pass", C.c_int32), ("sdv", SdvPlan * 2), ("gw", GwPlan),
                ("cells", GwCells), ("h_zterm", C.c_void_p * 2), ("lt_first", C.c_double * 2), ("lt_last", C.c_double * 2),
                ("dl", C.c_double * 2), ("seg", C.c_int32 * 4)]


class PlanDesc(C.Structure):
    _fields_ = [("n_y", C.c_int32), ("nt", C.c_int32), ("n_z_lookup", C.c_int32), ("nxi", C.c_int32),
                ("nuc_type", C.c_int32), ("mode", C.c_int32), ("phase_rule", C.c_int32), ("nz_int", C.c_int32),
                ("flags", C.c_int32), ("reserved", C.c_int32), ("cs", C.c_double), ("z_st_thresh", C.c_double),
                ("nuc_a", C.c_double), ("gamma", C.c_double)]


class SweepOpts(C.Structure):
    _fields_ = [("n_xi", C.c_int32), ("thin_shell_limit", C.c_int32), ("chunk", C.c_int32), ("shell_chunk", C.c_int32),
                ("flags", C.c_int32), ("group_min", C.c_int32), ("t_end", C.c_double), ("wn_rtol", C.c_double),
                ("r_star", C.c_double), ("lifetime_multiplier", C.c_double), ("host_v_wall", C.c_void_p),
                ("host_sol_type", C.c_void_p)]


class SweepOut(C.Structure):
    _fields_ = [("pow_v", C.c_void_p), ("pow_gw", C.c_void_p), ("omgw0", C.c_void_p), ("scalars", C.c_void_p),
                ("status", C.c_void_p), ("rows_v", C.c_void_p), ("rows_w", C.c_void_p), ("rows_xi", C.c_void_p),
                ("row_stride", C.c_int64), ("off", C.c_void_p), ("npts", C.c_void_p), ("table_rows", C.c_void_p),
                ("snr", C.c_void_p),
                ("snr_weight", C.c_void_p), ("n_snr", C.c_int32), ("reserved", C.c_int32)]


N_STAGES = 12
STAGE_NAMES = ("shells_generic", "shells_retry", "thermo", "a2", "spec_den_v_group_y", "spec_den_v_group", "spec_den_gw_cells",
               "emit", "retry_fold")
SWEEP_NSCAL = 32
SWEEP_SCALARS = ("vp", "vm", "vp_tilde", "vm_tilde", "v_sh", "vm_sh", "vm_tilde_sh", "wp", "wm", "wm_sh", "v_cj", "wn",
                 "wn_estimate", "solver_failed", "va_kinetic_energy_density", "va_trace_anomaly_diff",
                 "va_thermal_energy_density_diff", "va_entropy_density_diff", "wbar", "ebar", "kinetic_energy_density",
                 "kinetic_energy_fraction", "va_thermal_energy_density", "thermal_energy_density", "va_enthalpy_density",
                 "kappa", "omega", "ubarf2", "mean_adiabatic_index", "nu_gdh2024", "source_lifetime_factor", "npts")
SWEEP_TIMING, SWEEP_NO_RETRY = 1, 2
PLAN_ONLY_FIRST_PASS, PLAN_ONLY_LOOKUP_PASS, PLAN_FORCE_DIRECT, PLAN_NO_GW_CELLS = 1, 2, 4, 8
SDV_ALIGNED_BAND = 1


class SweepStats(C.Structure):
    _fields_ = [("stage_ms", C.c_double * N_STAGES), ("stage_calls", C.c_int64 * N_STAGES), ("launches", C.c_int64),
                ("retried", C.c_int64), ("rescued", C.c_int64), ("group_profiles", C.c_int64), ("group_count", C.c_int64),
                ("group_kw_sum", C.c_int64), ("group_macs", C.c_int64), ("a2_profiles", C.c_int64), ("gw_profiles", C.c_int64),
                ("chunks_before_fold", C.c_int64), ("chunks_after_fold", C.c_int64), ("host_wait_retry_ms", C.c_double),
                ("host_wait_shells_ms", C.c_double), ("timing_serial", C.c_int64)]


CHUNK_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p)

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
    "ptt_spec_den_v_group_ex": (C.c_int, [C.POINTER(SdvPlan), C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "ptt_spec_den_gw_batch": (C.c_int, [C.POINTER(GwPlan), C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptt_spec_den_gw_general_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
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
    "ptt_nearest_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "ptt_shell_wall_values": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptt_profile_integrals_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptt_ssm_plan_create_host": (C.c_int, [C.POINTER(PlanDesc), C.c_void_p, C.POINTER(C.c_void_p)]),
    "ptt_ssm_plan_upload": (C.c_int, [C.c_void_p]),
    "ptt_ssm_plan_create": (C.c_int, [C.POINTER(PlanDesc), C.c_void_p, C.POINTER(C.c_void_p)]),
    "ptt_ssm_plan_destroy": (C.c_int, [C.c_void_p]),
    "ptt_ssm_plan_view": (C.POINTER(SweepPlan), [C.c_void_p]),
    "ptt_ssm_plan_array": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "ptt_ssm_plan_info": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ptt_sweep_spectra": (C.c_int, [C.POINTER(SweepPlan), C.POINTER(SweepOpts), C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(SweepOut), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ptt_ipc_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "ptt_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "ptt_ipc_close": (C.c_int, [C.c_void_p]),
    "ptt_ipc_free": (C.c_int, [C.c_void_p]),
    "ptt_sweep_last_stats": (C.c_int, [C.POINTER(SweepStats)]),
    "ptt_sweep_timing_sync": (C.c_int, []),
    "ptt_sweep_timing_collect": (C.c_int, [C.c_int64, C.POINTER(SweepStats)]),
    "ptt_sweep_release": (C.c_int, []),
    "ptt_sweep_shell_chunk": (C.c_int64, [C.c_int, C.c_int64]),
    "ptt_suppression_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                        C.c_void_p, C.c_void_p]),
    "ptt_snr_batch": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
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
