"""ctypes access to the TEST-ONLY host emulation of the device headers (see hostemu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhostemu.so")


def build(force=False, defines=(), so=None):
    """defines: extra -D switches of the device headers (e.g. PTT_NO_STEP_SEARCH), built into a library of their own."""
    so = so or _SO
    src = os.path.join(_HERE, "hostemu.cpp")
    hdrs = [os.path.join(_HERE, "..", "..", "pttools_b200", "csrc", h) for h in os.listdir(
        os.path.join(_HERE, "..", "..", "pttools_b200", "csrc")) if h.endswith(".cuh")]
    newest = max(os.path.getmtime(p) for p in [src] + hdrs)
    if force or not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC"] +
                              [f"-D{d}" for d in defines] + ["-o", so, src])
    return so


_lib = None
_variants = {}


def use_variant(*defines):
    """Switches the module to a build of the emulation with the given -D switches (no arguments: the plain build)."""
    global _lib
    key = tuple(sorted(defines))
    if key not in _variants:
        so = _SO if not key else os.path.join(_HERE, "libhostemu_" + "_".join(k.lower() for k in key) + ".so")
        _variants[key] = _declare(C.CDLL(build(defines=key, so=so)))
    _lib = _variants[key]


def lib():
    if _lib is None:
        use_variant()
    return _lib


def _declare(_lib):
    d, i, pd, pi = C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)
    _lib.emu_integrate.argtypes = [d, d, d, d, d, i, pd, pd, pd]
    _lib.emu_anmax.argtypes = [d, pi]
    _lib.emu_anmax.restype = d
    _lib.emu_bag_forward.argtypes = [d, d, i, i, pd]
    _lib.emu_bag_find_alpha_plus.argtypes = [d, d, i, d, pd, pi, pi]
    _lib.emu_bag_profile.argtypes = [d, d, i, i, d, pd, pd, pd, pi, pi, pi]
    _lib.emu_generic_solve.argtypes = [pd, d, d, i, d, d, pd, i, d, i, d, pd, pd, pd, pi, pi, pd]
    _lib.emu_generic_rescue.argtypes = [pd, d, d, i, d, d, pd, i, d, i, d, pd, pd, pd, pi, pi, pd, pi]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def integrate(v0, w0, xi0, cs2, t_end, n):
    v, w, xi = np.empty(n), np.empty(n), np.empty(n)
    rc = lib().emu_integrate(v0, w0, xi0, cs2, t_end, n, _p(v), _p(w), _p(xi))
    return rc, v, w, xi


def anmax(v_wall):
    st = C.c_int(0)
    return lib().emu_anmax(v_wall, C.byref(st)), st.value


def bag_forward(v_wall, ap, sol_type, n_xi):
    out = np.empty(8)
    st = lib().emu_bag_forward(v_wall, ap, sol_type, n_xi, _p(out))
    return st, dict(i1=int(out[0]), t_end2=out[1], n_fwd=int(out[2]), v=out[3], w=out[4], xi=out[5], w1=out[6], wn_raw=out[7])


def find_alpha_plus(v_wall, alpha_n, n_xi, an_max):
    ap, st, it = C.c_double(0), C.c_int(0), C.c_int(0)
    rc = lib().emu_bag_find_alpha_plus(v_wall, alpha_n, n_xi, an_max, C.byref(ap), C.byref(st), C.byref(it))
    return rc, ap.value, st.value, it.value


def bag_profile(v_wall, ap, sol_type, n_xi, w_n=1.0):
    cap = lib().emu_row_capacity(n_xi)
    v, w, xi = np.full(cap, np.nan), np.full(cap, np.nan), np.full(cap, np.nan)
    off, npts, nw = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = lib().emu_bag_profile(v_wall, ap, sol_type, n_xi, w_n, _p(v), _p(w), _p(xi), C.byref(off), C.byref(npts), C.byref(nw))
    s = slice(off.value, off.value + npts.value)
    return rc, v[s].copy(), w[s].copy(), xi[s].copy(), nw.value


def bag_machine(v_wall, alpha_n, n_xi, an_max, mode=0, ap_in=0.0, sol_type_in=0, w_n=1.0):
    """BagMachine through the host emulation: (root_status, profile_status, alpha_plus, sol_type, iters, v, w, xi,
    n_wall, xi_sh, wn_raw)."""
    cap = lib().emu_row_capacity(n_xi)
    v, w, xi = np.full(cap, np.nan), np.full(cap, np.nan), np.full(cap, np.nan)
    off, npts, nw, st, it, rs = (C.c_int(0) for _ in range(6))
    ap, xs, wr = C.c_double(0), C.c_double(0), C.c_double(0)
    f = lib().emu_bag_machine
    f.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double] + \
        [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_double), C.POINTER(C.c_int),
                                                                  C.POINTER(C.c_int), C.POINTER(C.c_double),
                                                                  C.POINTER(C.c_double), C.POINTER(C.c_int)]
    rc = f(mode, v_wall, alpha_n, n_xi, an_max, ap_in, sol_type_in, w_n, _p(v), _p(w), _p(xi), C.byref(off),
           C.byref(npts), C.byref(nw), C.byref(ap), C.byref(st), C.byref(it), C.byref(xs), C.byref(wr), C.byref(rs))
    s = slice(off.value, off.value + npts.value)
    return rs.value, rc, ap.value, st.value, it.value, v[s].copy(), w[s].copy(), xi[s].copy(), nw.value, xs.value, wr.value


GEN_SCALARS = ["vp", "vm", "vp_tilde", "vm_tilde", "v_sh", "vm_sh", "vm_tilde_sh", "wp", "wm", "wm_sh", "v_cj", "wn",
               "wn_estimate", "solver_failed"]


def generic_solve(eos, v_wall, alpha_n, sol_type, wn, v_cj, guess, n_xi=5000, t_end=50.0, thin=100, wn_rtol=1e-4):
    """eos: (css2, csb2, V_s, V_b, bag_name); guess: (vp, vp_tilde, wp, wm)."""
    css2, csb2, V_s, V_b, bag_name = eos
    e = np.array([css2, csb2, 1 + 1 / css2, 1 + 1 / csb2, V_s, V_b, float(bag_name)])
    g = np.array(guess, dtype=float)
    cap = lib().emu_gen_row_capacity(n_xi)
    v, w, xi = np.full(cap, np.nan), np.full(cap, np.nan), np.full(cap, np.nan)
    off, npts = C.c_int(0), C.c_int(0)
    scal = np.empty(16)
    rc = lib().emu_generic_solve(_p(e), v_wall, alpha_n, sol_type, wn, v_cj, _p(g), n_xi, t_end, thin, wn_rtol,
                                 _p(v), _p(w), _p(xi), C.byref(off), C.byref(npts), _p(scal))
    s = slice(off.value, off.value + npts.value)
    return rc, v[s].copy(), w[s].copy(), xi[s].copy(), dict(zip(GEN_SCALARS, scal))


def generic_rescue(eos, v_wall, alpha_n, sol_type, wn, v_cj, guess, s_coef, n_xi=5000, t_end=50.0, thin=100, wn_rtol=1e-4):
    """The retry search for a point whose first root search failed.  Returns (winner, status, v, w, xi, scalars);
    winner = -1 if no candidate start converged.  s_coef = (s_coef_s, s_coef_b): s(w) = s_coef w^(1 - 1/mu)."""
    css2, csb2, V_s, V_b, bag_name = eos
    e = np.array([css2, csb2, 1 + 1 / css2, 1 + 1 / csb2, V_s, V_b, float(bag_name), s_coef[0], s_coef[1]])
    g = np.array(guess, dtype=float)
    cap = lib().emu_gen_row_capacity(n_xi)
    v, w, xi = np.full(cap, np.nan), np.full(cap, np.nan), np.full(cap, np.nan)
    off, npts, st = C.c_int(0), C.c_int(0), C.c_int(0)
    scal = np.full(16, np.nan)
    win = lib().emu_generic_rescue(_p(e), v_wall, alpha_n, sol_type, wn, v_cj, _p(g), n_xi, t_end, thin, wn_rtol,
                                   _p(v), _p(w), _p(xi), C.byref(off), C.byref(npts), _p(scal), C.byref(st))
    s = slice(off.value, off.value + npts.value)
    return win, st.value, v[s].copy(), w[s].copy(), xi[s].copy(), dict(zip(GEN_SCALARS, scal))
