"""Oracle (test infrastructure only): bag-model fluid shell, restated from the reference.

Every function names the reference lines it follows (paths relative to the reference root).  The ODE stepping is
``scipy.integrate.odeint`` exactly as at pttools/bubble/integrate.py:180 (no tolerances passed => LSODA defaults),
and the alpha_+ root is ``scipy.optimize.fsolve(xtol=1e-6, factor=0.1)`` as at pttools/bubble/alpha.py:325-330.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.integrate import odeint
from scipy.optimize import fsolve

CS0 = 1.0 / np.sqrt(3.0)          # pttools/bubble/const.py:29
CS0_2 = 1.0 / 3.0                 # const.py:31
N_XI_DEFAULT = 5000               # const.py:12
T_END_DEFAULT = 50.0              # const.py:18

SUB_DEF, HYBRID, DETON, ERROR = 0, 1, 2, -1   # integer codes for boundary.py:37-62 SolutionType
SYMMETRIC, BROKEN = 0.0, 1.0                  # boundary.py:27-33 Phase


# ---------------------------------------------------------------------------------------------------------------
# a1/a2: ODE right-hand side and tau-grid integration
# ---------------------------------------------------------------------------------------------------------------

def rhs(y, _t, cs2):
    """(dv, dw, dxi)/dtau; pttools/bubble/integrate.py:62-73."""
    v, w, xi = y
    xiv = xi * v
    xi_v = xi - v
    v2 = v * v
    dv = 2.0 * v * cs2 * (1.0 - v2) * (1.0 - xiv)
    dw = (w / (1.0 - v2)) * (xi_v / (1.0 - xiv)) * (1.0 / cs2 + 1.0) * dv
    dxi = xi * (xi_v ** 2 - cs2 * (1.0 - xiv) ** 2)
    return (dv, dw, dxi)


_ODEINT_KW: dict = {}
_ANMAX_CACHE: dict = {}


class lsoda_tolerances:
    """Context manager: run the oracle with explicit LSODA tolerances instead of scipy's defaults (1.49e-8).
    Used by the parity tests to separate the reference's own integrator noise from implementation differences
    (the tau-grid flips under 1e-8 perturbations)."""

    def __init__(self, rtol, atol):
        self.kw = {"rtol": rtol, "atol": atol}

    def __enter__(self):
        self.saved = dict(_ODEINT_KW)
        _ODEINT_KW.update(self.kw)
        _ANMAX_CACHE.clear()
        return self

    def __exit__(self, *exc):
        _ODEINT_KW.clear()
        _ODEINT_KW.update(self.saved)
        _ANMAX_CACHE.clear()
        return False


def integrate(v0, w0, xi0, phase, t_end, n_xi, css2=CS0_2, csb2=CS0_2, **odeint_kw):
    """fluid_integrate_param with method="odeint"; integrate.py:81-135, 167-189.

    cs2(w, phase) = phase*csb2 + (1-phase)*css2 covers both in-scope models
    (models/bag.py:21-27, models/const_cs.py:553-557).
    """
    t = np.linspace(0.0, t_end, n_xi)
    cs2 = phase * csb2 + (1.0 - phase) * css2
    sol = odeint(rhs, np.array([v0, w0, xi0]), t, args=(cs2,), **{**_ODEINT_KW, **odeint_kw})
    return sol[:, 0], sol[:, 1], sol[:, 2], t


# ---------------------------------------------------------------------------------------------------------------
# a6: closed-form bag relations
# ---------------------------------------------------------------------------------------------------------------

def gamma2(v):
    """relativity.py:21-30."""
    return 1.0 / (1.0 - v ** 2)


def lorentz(xi, v):
    """relativity.py:32-43."""
    return (xi - v) / (1.0 - v * xi)


def enthalpy_ratio(v_m, v_p):
    """w_-/w_+ ; boundary.py:89-101."""
    return gamma2(v_m) * v_m / (gamma2(v_p) * v_p)


def v_plus(vm, ap, sol_type):
    """boundary.py:504-518 (scalar)."""
    x = vm + 1.0 / (3.0 * vm)
    b = 1.0 if sol_type == DETON else -1.0
    if b == -1.0 and ap > 1.0 / 3.0:
        return np.nan
    return (0.5 / (1.0 + ap)) * (x + b * np.sqrt(x ** 2 + 4.0 * ap ** 2 + (8.0 / 3.0) * ap - (4.0 / 3.0)))


def v_minus(vp, ap, sol_type=DETON, strong_branch=False):
    """boundary.py:376-408 (scalar)."""
    if vp < 0:
        return np.nan
    vp2 = vp ** 2
    y = vp2 + 1.0 / 3.0
    z = y - ap * (1.0 - vp2)
    x = (4.0 / 3.0) * vp2
    b = 1.0 if sol_type == DETON else -1.0
    c = -1.0 if strong_branch else 1.0
    with np.errstate(invalid="ignore"):
        return (0.5 / vp) * (z + b * c * np.sqrt(z ** 2 - x))


def v_shock_bag(xi):
    """shock.py:423-430; NaN below 1/sqrt(3). Scalar or array."""
    xi = np.asarray(xi, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = (3.0 * xi ** 2 - 1.0) / (2.0 * xi)
    out = np.where(xi < CS0, np.nan, out)
    return out if out.ndim else float(out)


def wm_shock_bag(xi, w_n=1.0):
    """shock.py:469-476 (nan_on_negative=True)."""
    xi = np.asarray(xi, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = w_n * (9.0 * xi ** 2 - 1.0) / (3.0 * (1.0 - xi ** 2))
    out = np.where(xi < CS0, np.nan, out)
    out = np.where(xi == 1.0, np.inf, out)
    return out if out.ndim else float(out)


def fluid_speeds_at_wall(v_wall, alpha_p, sol_type):
    """(vfp_w, vfm_w, vfp_p, vfm_p); boundary.py:109-158."""
    if sol_type == SUB_DEF:
        vfp_w, vfm_w = v_plus(v_wall, alpha_p, sol_type), v_wall
    elif sol_type == HYBRID:
        vfp_w, vfm_w = v_plus(CS0, alpha_p, sol_type), CS0
    elif sol_type == DETON:
        vfp_w, vfm_w = v_wall, v_minus(v_wall, alpha_p)
    else:
        raise ValueError("Unknown sol_type")
    return vfp_w, vfm_w, lorentz(v_wall, vfp_w), lorentz(v_wall, vfm_w)


def alpha_plus_max_detonation(v_wall):
    """alpha.py:192-204."""
    if v_wall < CS0:
        return 0.0
    return (1.0 - np.sqrt(3.0) * v_wall) ** 2 / (3.0 * (1.0 - v_wall ** 2))


def alpha_plus_min_hybrid(v_wall):
    """alpha.py:208-225."""
    if v_wall < CS0:
        return 0.0
    return (1.0 - np.sqrt(3.0) * v_wall) ** 2 / (9.0 * v_wall ** 2 - 1.0)


def find_v_index(xi, v_target):
    """props.py:13-19."""
    return int(np.argmax(xi >= v_target))


# ---------------------------------------------------------------------------------------------------------------
# a7: index-based event location
# ---------------------------------------------------------------------------------------------------------------

def trim_wall_to_shock(v, w, xi, t, sol_type):
    """trim.py:77-116. The index starts at -2; 0 is promoted to 1."""
    n = -2
    if sol_type != DETON:
        hit = np.nonzero(v <= v_shock_bag(xi))[0]     # NaN compares False, as in the reference loop
        if hit.size:
            n = int(hit[0])
    if n == 0:
        n = 1
    return v[:n + 1], w[:n + 1], xi[:n + 1], t[:n + 1]


def trim_wall_to_cs(v, w, xi, t, v_wall, sol_type, cs2_b=CS0_2):
    """trim.py:20-73."""
    n_start = 0
    n = -2
    if sol_type != SUB_DEF:
        hit = np.nonzero((v <= 0) | (xi ** 2 <= cs2_b))[0]
        if hit.size:
            n = int(hit[0])
    n_stop = 1 if n == 0 else n
    if xi[0] == v_wall and sol_type != DETON:
        n_start = 1
        n_stop += 1
    return v[n_start:n_stop], w[n_start:n_stop], xi[n_start:n_stop], t[n_start:n_stop]


def shock_zoom_last_element(v, w, xi):
    """shock.py:266-294; returns modified copies."""
    v, w, xi = v.copy(), w.copy(), xi.copy()
    v_sh = v_shock_bag(xi)
    if v[-1] < v_sh[-1] and v[-2] > v_sh[-2] and xi[-1] > xi[-2]:
        dxi = xi[-1] - xi[-2]
        dv = v[-1] - v[-2]
        dv_sh = v_sh[-1] - v_sh[-2]
        dw_sh = w[-1] - w[-2]
        dxi_sh = dxi * (v[-2] - v_sh[-2]) / (dv_sh - dv)
        xi[-1] = xi[-2] + dxi_sh
        v[-1] = v[-2] + (dv_sh / dxi) * dxi_sh
        w[-1] = w[-2] + (dw_sh / dxi) * dxi_sh
    return v, w, xi


# ---------------------------------------------------------------------------------------------------------------
# a3: profile for a given alpha_+
# ---------------------------------------------------------------------------------------------------------------

def shell_alpha_plus(v_wall, alpha_plus, sol_type, n_xi=N_XI_DEFAULT, w_n=1.0):
    """sound_shell_alpha_plus; fluid_bag.py:83-222."""
    if sol_type == ERROR:
        nan = np.array([np.nan])
        return nan, nan, nan
    vfp_w, vfm_w, vfp_p, vfm_p = fluid_speeds_at_wall(v_wall, alpha_plus, sol_type)
    wp = 1.0
    wm = wp / enthalpy_ratio(vfm_w, vfp_w)
    dxi = 1.0 / n_xi

    xif = np.linspace(v_wall + dxi, 1.0, 2)
    vf = np.zeros_like(xif)
    wf = np.ones_like(xif) * wp
    xib = np.linspace(min(CS0_2 ** 0.5, v_wall) - dxi, 0.0, 2)
    vb = np.zeros_like(xib)
    wb = np.ones_like(xib) * wm

    if sol_type != DETON:
        v, w, xi, t = integrate(vfp_p, wp, v_wall, SYMMETRIC, -T_END_DEFAULT, N_XI_DEFAULT)
        v, w, xi, t = trim_wall_to_shock(v, w, xi, t, sol_type)
        v, w, xi, t = integrate(vfp_p, wp, v_wall, SYMMETRIC, t[-1], n_xi)
        v, w, xi, t = trim_wall_to_shock(v, w, xi, t, sol_type)
        v, w, xi = shock_zoom_last_element(v, w, xi)
        vf = np.concatenate((v, vf))
        vfp_s = xi[-1]
        vfm_s = 1.0 / (3.0 * vfp_s)
        wf = np.concatenate((w, np.ones_like(xif) * w[-1] * enthalpy_ratio(vfm_s, vfp_s)))
        xif[0] = xi[-1]
        xif = np.concatenate((xi, xif))

    if sol_type != SUB_DEF:
        v, w, xi, t = integrate(vfm_p, wm, v_wall, BROKEN, -T_END_DEFAULT, N_XI_DEFAULT)
        v, w, xi, t = trim_wall_to_cs(v, w, xi, t, v_wall, sol_type)
        vb = np.concatenate((v, vb))
        wb = np.concatenate((w, np.ones_like(xib) * w[-1]))
        xib[0] = CS0_2 ** 0.5
        xib = np.concatenate((xi, xib))

    v = np.concatenate((np.flipud(vb), vf))
    w = np.concatenate((np.flipud(wb), wf))
    w = w * (w_n / w[-1])
    xi = np.concatenate((np.flipud(xib), xif)).copy()
    return v, w, xi


# ---------------------------------------------------------------------------------------------------------------
# a4/a5: alpha_n <-> alpha_+, classification, full shell
# ---------------------------------------------------------------------------------------------------------------

def alpha_n_max_deflagration(v_wall, n_xi=N_XI_DEFAULT):
    """alpha.py:44-50 (scalar).  Depends on v_wall only; memoised here because the reference recomputes it up to
    four times per point (transition.py:27, alpha.py:185, alpha.py:316, bubble.py:252)."""
    key = (float(v_wall), int(n_xi))
    if key not in _ANMAX_CACHE:
        sol_type = HYBRID if v_wall > CS0 else SUB_DEF
        ap = 1.0 / 3 - 1.0e-10
        _, w, xi = shell_alpha_plus(v_wall, ap, sol_type, n_xi)
        n_wall = find_v_index(xi, v_wall)
        _ANMAX_CACHE[key] = w[n_wall + 1] * (1.0 / 3)
    return _ANMAX_CACHE[key]


def identify_solution_type(v_wall, alpha_n):
    """identify_solution_type_bag; transition.py:19-44."""
    if alpha_n < alpha_plus_max_detonation(v_wall):
        return DETON
    if alpha_n < alpha_n_max_deflagration(v_wall):
        return SUB_DEF if v_wall <= CS0 else HYBRID
    return ERROR


def find_alpha_n(v_wall, alpha_p, sol_type, n_xi=N_XI_DEFAULT):
    """find_alpha_n_bag; alpha.py:229-255."""
    _, w, xi = shell_alpha_plus(v_wall, alpha_p, sol_type, n_xi)
    n_wall = find_v_index(xi, v_wall)
    return alpha_p * w[n_wall] / w[-1]


def alpha_plus_initial_guess(v_wall, alpha_n):
    """alpha.py:168-188."""
    if alpha_n < 0.05:
        return alpha_n
    ap_min = alpha_plus_min_hybrid(v_wall)
    ap_max = 1.0 / 3
    an_min = alpha_plus_max_detonation(v_wall)
    an_max = alpha_n_max_deflagration(v_wall)
    slope = (ap_max - ap_min) / (an_max - an_min)
    return ap_min + slope * (alpha_n - an_min)


def find_alpha_plus(v_wall, alpha_n, n_xi=N_XI_DEFAULT, xtol=1e-6):
    """_find_alpha_plus_scalar_bag; alpha.py:305-331."""
    if alpha_n < alpha_plus_max_detonation(v_wall):
        return alpha_n
    if alpha_n >= alpha_n_max_deflagration(v_wall):
        return np.nan
    sol_type = SUB_DEF if v_wall <= CS0 else HYBRID
    guess = alpha_plus_initial_guess(v_wall, alpha_n)

    def resid(ap):
        return find_alpha_n(v_wall, float(ap[0]), sol_type, n_xi) - alpha_n

    return float(fsolve(resid, guess, xtol=xtol, factor=0.1)[0])


def sound_shell_bag(v_wall, alpha_n, n_xi=N_XI_DEFAULT):
    """fluid_bag.py:33-79. Returns (v, w, xi); length-1 NaN arrays on failure."""
    sol_type = identify_solution_type(v_wall, alpha_n)
    nan = np.array([np.nan])
    if sol_type == ERROR:
        return nan, nan, nan
    al_p = find_alpha_plus(v_wall, alpha_n, n_xi)
    if np.isnan(al_p):
        return nan, nan, nan
    return shell_alpha_plus(v_wall, al_p, sol_type, n_xi)


# ---------------------------------------------------------------------------------------------------------------
# Scalars used by the reference's golden files
# ---------------------------------------------------------------------------------------------------------------

def ubarf_squared(v, w, xi, v_wall):
    """quantities.py:510-527 via mean_kinetic_energy :430-445."""
    return np.trapezoid(w * v ** 2 * gamma2(v), xi ** 3) / v_wall ** 3 / w[-1]


def mean_enthalpy_change(v, w, xi, v_wall):
    """quantities.py:409-426."""
    return np.trapezoid(w - w[-1], xi ** 3) / v_wall ** 3


def get_phase(xi, v_wall):
    """boundary.py:161-186: 1 where xi < v_wall."""
    return np.where(xi < v_wall, BROKEN, SYMMETRIC)


def de_from_w_bag(w, xi, v_wall, alpha_n):
    """quantities.py:38-54 with bag.e_bag/p_bag (bag.py:117-160): e = 3/4 w + theta(phase)."""
    theta_s = 0.75 * w[-1] * alpha_n
    theta = theta_s * (1.0 - get_phase(xi, v_wall))
    e = w - (0.25 * w - theta)
    return e - e[-1]


def sound_shell_dict(v_wall, alpha_n, n_p=N_XI_DEFAULT):
    """The numbers tests/paper/test_shells_bag.py:28-56 pins; fluid_bag.py:225-299."""
    sol_type = identify_solution_type(v_wall, alpha_n)
    v, w, xi = sound_shell_bag(v_wall, alpha_n, n_p)
    xi_even = np.linspace(1 / n_p, 1 - 1 / n_p, n_p)
    n_wall = find_v_index(xi, v_wall)
    r = w[n_wall + 1] / w[n_wall] if sol_type == DETON else w[n_wall] / w[n_wall - 1]
    ubarf2 = ubarf_squared(v, w, xi, v_wall)
    return {
        "v": v, "w": w, "xi": xi,
        "v_sh": v_shock_bag(xi_even), "w_sh": wm_shock_bag(xi_even),
        "n_wall": n_wall, "n_cs": int(math.floor(CS0 * n_p)), "n_sh": xi.size - 2, "r": r,
        "alpha_plus": alpha_n * w[-1] / w[n_wall],
        "ubarf2": ubarf2, "ke_frac": ubarf2 / (0.75 * (1 + alpha_n)), "kappa": ubarf2 / (0.75 * alpha_n),
        "dw": 0.75 * mean_enthalpy_change(v, w, xi, v_wall) / (0.75 * alpha_n * w[-1]),
    }
