"""Oracle (test infrastructure only): the model-independent fluid shell solver, restated from the reference.

Follows pttools/bubble/fluid.py (sound_shell_generic and its solvers), boundary.py (solve_junction), shock.py
(find_shock_index, v_shock, solve_shock), chapman_jouguet.py, fluid_reference.py, bubble.py (Bubble.solve) and
thermo.py.  The root finders are the same scipy entry points as at the reference's call sites: fsolve
(speedup/solvers.py:25,45) and root_scalar with the secant method (fluid.py:619-624).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import fsolve, root_scalar

from . import shell_bag as sb
from .models import BROKEN, SYMMETRIC, Eos
from .shell_bag import DETON, ERROR, HYBRID, SUB_DEF, N_XI_DEFAULT, T_END_DEFAULT, gamma2, lorentz

JUNCTION_RTOL = 1e-6                 # bubble/const.py:34
THIN_SHELL_T_POINTS_MIN = 100        # const.py:26


# --- speedup/solvers.py:12-55 ----------------------------------------------------------------------------------

def fsolve_vary(func, x0, args=(), abs_variations=1e-3, rel_variations=0.01):
    sol = fsolve(func, x0=x0, args=args, full_output=True)
    if sol[2] == 1:
        return sol
    for i in range(x0.shape[0]):
        for sign in (1, -1):
            x0_var = x0.copy()
            x0_var[i] *= 1 + sign * rel_variations
            x0_var[i] += sign * abs_variations
            sol2 = fsolve(func, x0=x0_var, args=args, full_output=True)
            if sol2[2] == 1:
                return sol2
    return sol


# --- boundary.py:232-373 ---------------------------------------------------------------------------------------

def _junction_devs(params, model: Eos, v1, w1, phase1, phase2):
    v2, w2 = params[0], params[1]
    d1 = w1 * gamma2(v1) * v1 - w2 * gamma2(v2) * v2
    d2 = w1 * gamma2(v1) * v1 ** 2 + model.p(w1, phase1) - w2 * gamma2(v2) * v2 ** 2 - model.p(w2, phase2)
    return np.array([float(d1), float(d2)])


def solve_junction(model: Eos, v1_tilde, w1, phase1, phase2, v2_tilde_guess, w2_guess, allow_failure=False,
                   allow_negative_entropy_flux_change=False, rtol=JUNCTION_RTOL):
    """boundary.py:255-357."""
    bad = (np.isnan(v1_tilde) or np.isnan(w1) or np.isnan(v2_tilde_guess) or np.isnan(w2_guess)
           or v1_tilde < 0 or v1_tilde > 1 or w1 < 0 or v2_tilde_guess < 0 or v2_tilde_guess > 1 or w2_guess < 0
           or np.isclose(v1_tilde, 0) or np.isclose(v1_tilde, 1) or np.isclose(v2_tilde_guess, 0)
           or np.isclose(v2_tilde_guess, 1) or np.isclose(w1, 0) or np.isclose(w2_guess, 0))
    if bad:
        return np.nan, np.nan
    with np.errstate(all="ignore"):
        sol = fsolve_vary(_junction_devs, x0=np.array([v2_tilde_guess, w2_guess]),
                          args=(model, v1_tilde, w1, phase1, phase2))
    v2, w2 = sol[0][0], sol[0][1]
    if sol[2] != 1 and not allow_failure:
        return np.nan, np.nan
    with np.errstate(all="ignore"):
        devs_rel = _junction_devs(np.array([v2, w2]), model, v1_tilde, w1, phase1, phase2) / w1
    if np.max(np.abs(devs_rel)) > rtol and not allow_failure:
        return np.nan, np.nan
    if (v2 < 0 or v2 > 1 or w2 < 0) and not allow_failure:
        return np.nan, np.nan
    # check_entropy_fluxes, boundary.py:65-85
    with np.errstate(all="ignore"):
        f1 = np.sqrt(gamma2(v1_tilde)) * v1_tilde * model.s(w1, phase1)
        f2 = np.sqrt(gamma2(v2)) * v2 * model.s(w2, phase2)
    fail = f1 < 0 or f2 < 0
    if not allow_negative_entropy_flux_change:
        fail = fail or (phase1 == SYMMETRIC and phase2 == BROKEN and f1 - f2 < 0) or \
            (phase1 == BROKEN and phase2 == SYMMETRIC and f2 - f1 < 0)
    if fail and not allow_failure:
        return np.nan, np.nan
    return v2, w2


def w2_junction(v1, w1, v2):
    """boundary.py:638-648."""
    return w1 * gamma2(v1) * v1 / (gamma2(v2) * v2)


# --- shock.py:297-408 ------------------------------------------------------------------------------------------

def solve_shock(model: Eos, v1_tilde, w1, csp, phase=SYMMETRIC):
    if v1_tilde < 0 or v1_tilde > 1 or np.isclose(v1_tilde, 0) or np.isnan(v1_tilde):
        return np.nan, np.nan
    if np.isclose(w1, 0) or np.isnan(w1):
        return np.nan, np.nan
    if np.isclose(v1_tilde, 1):
        return 1, np.nan
    csp2 = csp ** 2
    if np.isclose(v1_tilde, csp):
        return v1_tilde, w1
    v2_guess = csp2 * v1_tilde
    if np.isclose(v2_guess, 0) or np.isclose(v2_guess, 1):
        return np.nan, np.nan
    if v1_tilde < csp:
        return np.nan, np.nan
    w2_guess = w2_junction(v1_tilde, w1, v2_guess)
    if w2_guess < 0 or np.isclose(w2_guess, 0):
        return np.nan, np.nan
    return solve_junction(model, v1_tilde, w1, phase, phase, v2_guess, w2_guess)


def v_shock(model: Eos, wn, xi, cs_n):
    if xi <= cs_n or np.isclose(xi, cs_n):
        return 0.0
    if np.isclose(xi, 1):
        return 1.0
    vt, _ = solve_shock(model, xi, wn, cs_n)
    ret = lorentz(xi, vt)
    return 0.0 if ret < 0 else ret


def find_shock_index(model: Eos, v, w, xi, v_wall, wn, cs_n, sol_type):
    """shock.py:49-248 with error_on_failure=False, zero_on_failure=True."""
    if sol_type == DETON:
        return sb.find_v_index(xi, v_wall)
    if model.name == "bag":
        with np.errstate(invalid="ignore"):
            after = np.logical_and(xi > v_wall, v <= sb.v_shock_bag(xi))
        if np.sum(after):
            return int(np.argmax(after))
        near = np.logical_and(np.isclose(xi, sb.CS0), np.isclose(v, 0))
        if np.sum(near):
            return int(np.argmax(near))
        raise RuntimeError("Did not find shock for the bag model.")
    if np.any(np.isnan(xi)) or np.any(np.isnan(v)):
        return 0
    if not np.isclose(xi[0], v_wall):
        return 0
    i_right = int(np.argmax(xi))
    if i_right == 0:
        return 0
    i_close = int(np.argmax(np.logical_and(np.isclose(xi, cs_n), np.isclose(v, 0))))
    i_cs_n = int(np.argmax(xi > cs_n))
    if i_close != 0 and ((i_cs_n == 0 and xi[0] < cs_n) or i_close <= i_cs_n):
        return i_close
    i_left = i_cs_n
    if i_left > i_right:
        return 0
    if not np.all(np.diff(xi[i_left:i_right])):
        return 0
    if v[0] < v_shock(model, wn, xi[0], cs_n):
        return 0
    v_sh_xi_max = v_shock(model, wn, xi[i_right], cs_n)
    if v[i_right] > v_sh_xi_max:
        if v_sh_xi_max - v[i_right] < 3.5e-8:
            return i_right
        return 0
    i_sh = 0
    while i_right - i_left > 1:
        i_sh = i_left + (i_right - i_left) // 2
        v_sh = v_shock(model, wn, xi[i_sh], cs_n)
        if v[i_sh] > v_sh:
            i_left = i_sh
        elif v[i_sh] < v_sh:
            i_right = i_sh
        else:
            return i_sh + 1
    if i_sh == 0:
        return 0
    if v[i_sh] > v_shock(model, wn, xi[i_sh], cs_n):
        if v[i_sh + 1] > v_shock(model, wn, xi[i_sh + 1], cs_n):
            return 0
        return i_sh + 1
    return i_sh


# --- chapman_jouguet.py:151-264 --------------------------------------------------------------------------------

def v_chapman_jouguet(model: Eos, alpha_n, wn):
    if model.kind == "bag":
        return 1 / np.sqrt(3) * (1 + np.sqrt(2 * alpha_n + 3 * alpha_n ** 2)) / (1 + alpha_n)
    atbp = model.alpha_theta_bar_n_from_alpha_n(alpha_n, wn)
    csb = np.sqrt(model.csb2)
    disc = 3 * atbp * (1 - model.csb2 + 3 * model.csb2 * atbp)
    ret = (1 + np.sqrt(disc)) / (1 / csb + 3 * csb * atbp)
    if ret < csb or ret > 1:
        raise ValueError(f"Invalid v_CJ for alpha_theta_bar_plus={atbp}: {ret}")
    return ret


def solution_type(model: Eos, v_wall, alpha_n, wn):
    if model.kind == "bag":
        return sb.identify_solution_type(v_wall, alpha_n)          # bag.py:286-294
    if v_wall ** 2 < model.csb2:                                     # const_cs.py:633-655
        return SUB_DEF
    atbp = model.alpha_theta_bar_n_from_alpha_n(alpha_n, wn)
    b = 3 * atbp - 1 - v_wall ** 2 * (1 / model.csb2 + 3 * atbp)
    if b > 0:
        return HYBRID
    a = v_wall / model.csb2
    c = v_wall
    if b ** 2 - 4 * a * c < 0:
        return HYBRID
    return DETON


# --- fluid_reference.py ---------------------------------------------------------------------------------------

class FluidReference:
    """Nearest-neighbour starting guesses from the 100 x 100 bag table (fluid_reference.py:24-163).
    The table itself comes from a fixture written by the reference (tests/golden/ref_fluid_reference.npz) or from
    `build()` with the oracle's own bag solver."""

    def __init__(self, table: dict):
        from scipy.interpolate import NearestNDInterpolator
        self.v_wall, self.alpha_n = table["v_wall"], table["alpha_n"]
        self.data = np.stack([table[k] for k in ("vp", "vm", "vp_tilde", "vm_tilde", "wp", "wm")], axis=2)
        self.interp = {
            SUB_DEF: NearestNDInterpolator(x=table["coords_sub_def"], y=table["inds_sub_def"]),
            HYBRID: NearestNDInterpolator(x=table["coords_hybrid"], y=table["inds_hybrid"]),
            DETON: NearestNDInterpolator(x=table["coords_detonation"], y=table["inds_detonation"]),
        }

    def get(self, v_wall, alpha_n, sol_type):
        ind = int(self.interp[sol_type](v_wall, alpha_n))
        return self.data[ind // self.v_wall.size, ind % self.v_wall.size]


def v_and_w_from_solution(v, w, xi, v_wall, sol_type):
    """props.py:52-87."""
    i_wall = int(np.argmax(v))
    if sol_type == DETON:
        i_wall += 1
    vp, vm = v[i_wall], v[i_wall - 1]
    return vp, vm, lorentz(v_wall, vp), lorentz(v_wall, vm), w[i_wall], w[i_wall - 1]


# --- fluid.py --------------------------------------------------------------------------------------------------

_NAN13 = (np.array([np.nan]),) * 3 + (np.nan,) * 10


def deflagration_common(model: Eos, v_wall, vm_tilde, wn, wm, cs_n, v_cj, vp_tilde_guess, wp_guess, sol_type, n_xi,
                        t_end, thin_shell_limit, allow_failure, allow_neg):
    """sound_shell_deflagration_common, fluid.py:121-288."""
    if v_wall < 0 or v_wall > 1 or vm_tilde < 0 or vm_tilde > 1 or wn < 0 or wm < 0 or cs_n < 0 or cs_n > 1 \
            or vp_tilde_guess < 0 or vp_tilde_guess > 1 or wp_guess < 0 or v_wall > v_cj:
        return _NAN13
    vp_tilde, wp = solve_junction(model, vm_tilde, wm, BROKEN, SYMMETRIC, vp_tilde_guess, wp_guess,
                                  allow_failure=allow_failure, allow_negative_entropy_flux_change=allow_neg)
    vp = -lorentz(vp_tilde, v_wall)
    if np.isnan(vp) or vp < 0 or wp < 0:
        return _NAN13
    css2, csb2 = model.css2, model.csb2
    v, w, xi, t = sb.integrate(vp, wp, v_wall, SYMMETRIC, -t_end, n_xi, css2, csb2)
    if np.argmax(xi) == 0:
        return _NAN13
    i_shock = find_shock_index(model, v, w, xi, v_wall, wn, cs_n, sol_type)
    if i_shock == 0 or i_shock + 1 >= xi.size:
        return _NAN13
    attempts = 5
    if i_shock < thin_shell_limit:
        step = 20
        t_end2 = t[i_shock + step]
        for _ in range(attempts):
            v2, w2, xi2, t2 = sb.integrate(vp, wp, v_wall, SYMMETRIC, t_end2, n_xi, css2, csb2)
            if np.argmax(xi2) == 0:
                break
            i2 = find_shock_index(model, v2, w2, xi2, v_wall, wn, cs_n, sol_type)
            if i2 == 0 or i2 + step >= xi2.size:
                break
            i_shock, v, w, xi, t = i2, v2, w2, xi2, t2
            if i_shock >= thin_shell_limit:
                break
            t_end2 = t[i_shock + step]
    if i_shock <= 1:
        return _NAN13
    v, w, xi = v[:i_shock], w[:i_shock], xi[:i_shock]
    xi_sh, vm_sh, wm_sh = xi[-1], v[-1], w[-1]
    vm_tilde_sh = lorentz(xi_sh, vm_sh)
    wn_estimate = w2_junction(vm_tilde_sh, wm_sh, xi_sh)
    vm = lorentz(vm_tilde, v_wall)
    return v, w, xi, vp, vm, vp_tilde, vm_tilde, xi_sh, vm_sh, vm_tilde_sh, wp, wn_estimate, wm_sh


def deflagration(model: Eos, v_wall, wn, w_center, cs_n, v_cj, vp_guess, wp_guess, t_end, n_xi, thin_shell_limit,
                 allow_failure=False, allow_neg=False):
    """sound_shell_deflagration, fluid.py:48-118."""
    if vp_guess is None or np.isnan(vp_guess) or wp_guess is None or np.isnan(wp_guess):
        alpha_minus = 4 * (0 - 1) / (3 * w_center)
        vp_tilde_guess = sb.v_minus(v_wall, alpha_minus, SUB_DEF)
        vp_guess = -lorentz(vp_tilde_guess, v_wall)
        wp_guess = w2_junction(v_wall, w_center, vp_tilde_guess)
    else:
        vp_tilde_guess = -lorentz(vp_guess, v_wall)
    if np.isnan(wn) or wn < 0 or np.isnan(w_center) or w_center < 0 or np.isnan(vp_tilde_guess) \
            or vp_tilde_guess < 0 or vp_tilde_guess > 1 or np.isnan(vp_guess) or vp_guess < 0 \
            or np.isnan(wp_guess) or wp_guess < 0:
        return _NAN13
    if wp_guess < wn:
        wp_guess = 1.1 * wn
    return deflagration_common(model, v_wall, v_wall, wn, w_center, cs_n, v_cj, vp_tilde_guess, wp_guess, SUB_DEF,
                               n_xi, t_end, thin_shell_limit, allow_failure, allow_neg)


def hybrid(model: Eos, v_wall, wn, wm, cs_n, v_cj, vp_tilde_guess, wp_guess, t_end, n_xi, thin_shell_limit,
           allow_failure=False, allow_neg=False):
    """sound_shell_hybrid, fluid.py:481-546."""
    vm_tilde = np.sqrt(model.cs2(wm, BROKEN))
    if np.isnan(vp_tilde_guess):
        vp_tilde_guess = 0.75 * vm_tilde
    if np.isnan(wp_guess):
        wp_guess = 2 * wm
    ret = deflagration_common(model, v_wall, vm_tilde, wn, wm, cs_n, v_cj, vp_tilde_guess, wp_guess, HYBRID, n_xi,
                              t_end, thin_shell_limit, allow_failure, allow_neg)
    if not np.isnan(ret[4]):
        return ret
    vm = lorentz(v_wall, vm_tilde)
    v_sh_est = v_shock(model, wn, v_wall, cs_n)
    vp_guess = lorentz(v_wall, vp_tilde_guess)
    if np.isnan(vp_tilde_guess) or vp_guess < v_sh_est or vp_guess < vm or np.isnan(wp_guess) or wp_guess < wn \
            or wp_guess < wm:
        vp_guess = 1.05 * v_sh_est
        vp_tilde_guess = lorentz(v_wall, vp_guess)
        wp_guess = wn + 1.3 * np.abs(wm - wn)
    ret2 = deflagration_common(model, v_wall, vm_tilde, wn, wm, cs_n, v_cj, vp_tilde_guess, wp_guess, HYBRID, n_xi,
                               t_end, thin_shell_limit, allow_failure, allow_neg)
    if not np.isnan(ret2[4]):
        return ret2
    return ret


def detonation(model: Eos, v_wall, alpha_n, wn, v_cj, vm_tilde_guess, wm_guess, t_end, n_xi):
    """sound_shell_detonation, fluid.py:351-478."""
    vp_tilde_bag, vm_tilde_bag, _, _ = sb.fluid_speeds_at_wall(v_wall, alpha_n, DETON)
    wm_bag = w2_junction(vp_tilde_bag, wn, vm_tilde_bag)
    if not np.isnan(vm_tilde_bag):
        vm_tilde_guess = vm_tilde_bag
    if not np.isnan(wm_bag):
        wm_guess = wm_bag
    v_mu_guess = lorentz(v_wall, np.sqrt(model.cs2(wm_guess, BROKEN)))
    vm_guess = lorentz(v_wall, vm_tilde_guess)
    if vm_guess > v_mu_guess:
        vm_guess = v_mu_guess
        vm_tilde_guess = lorentz(v_wall, vm_guess)
    vm_tilde, wm = solve_junction(model, v_wall, wn, SYMMETRIC, BROKEN, vm_tilde_guess, wm_guess,
                                  allow_negative_entropy_flux_change=True)
    vm = lorentz(v_wall, vm_tilde)
    csb = np.sqrt(model.cs2(wm, BROKEN))
    v_mu = lorentz(v_wall, csb)
    found = bool(vm <= v_mu)
    for guess_v in (0.7 * csb, lorentz(v_wall, 0.5 * v_mu)):
        if found:
            break
        vt2, wm2 = solve_junction(model, v_wall, wn, SYMMETRIC, BROKEN, guess_v, (wm_guess + wn) / 2,
                                  allow_negative_entropy_flux_change=True)
        vm2 = lorentz(v_wall, vt2)
        found = bool(vm2 < v_mu)
        if found:
            vm_tilde, vm, wm = vt2, vm2, wm2
    v, w, xi, t = sb.integrate(vm, wm, v_wall, BROKEN, -t_end, n_xi, model.css2, model.csb2)
    v, w, xi, t = sb.trim_wall_to_cs(v, w, xi, t, v_wall, DETON, cs2_b=model.csb2)
    return np.flip(v), np.flip(w), np.flip(xi), 0, vm, v_wall, vm_tilde, v_wall, vm, vm_tilde, wn, wm, wm, found


def sound_shell_generic(model: Eos, v_wall, alpha_n, ref: FluidReference, sol_type=None, wn=None, wn_rtol=1e-4,
                        t_end=T_END_DEFAULT, n_xi=N_XI_DEFAULT, thin_shell_limit=THIN_SHELL_T_POINTS_MIN):
    """fluid.py:848-1052 (default arguments: no user guesses, not the bag or Giese solver).
    Returns dict with v, w, xi and the scalars of the reference's 17-tuple."""
    if wn is None:
        wn = model.wn(alpha_n)
    cs_n = float(np.sqrt(model.cs2(wn, SYMMETRIC)))
    if sol_type is None:
        sol_type = solution_type(model, v_wall, alpha_n, wn)
    if sol_type == ERROR:
        raise ValueError("Could not determine solution type")
    vp_ref, vm_ref, vp_tilde_ref, vm_tilde_ref, wp_ref, wm_ref = ref.get(v_wall, alpha_n, sol_type)
    vp_guess, vp_tilde_guess = vp_ref, vp_tilde_ref
    wp_guess = wp_ref * wn
    wm_guess = 0.3 * wn if np.isnan(wm_ref) else wm_ref * wn
    if vp_guess < 0 or vp_guess > 1 or vp_tilde_guess < 0 or vp_tilde_guess > 1 or wm_guess < 0 or wp_guess < wn:
        raise ValueError("Got invalid guesses")
    v_cj = v_chapman_jouguet(model, alpha_n, wn)
    dxi = 1.0 / n_xi

    if sol_type == DETON:
        v, w, xi, vp, vm, vp_tilde, vm_tilde, v_sh, vm_sh, vm_tilde_sh, wp, wm, wm_sh, found = detonation(
            model, v_wall, alpha_n, wn, v_cj, vm_tilde_ref, wm_ref, t_end, n_xi)
    elif sol_type == SUB_DEF:
        if v_wall ** 2 > model.csb2:
            raise ValueError("Invalid parameters for a subsonic deflagration")
        if vp_guess > v_wall:
            vp_guess = 0.95 * v_wall

        def solvable(w_center):
            if isinstance(w_center, np.ndarray):
                w_center = w_center[0]
            if np.isnan(w_center) or w_center < 0:
                return np.nan
            out = deflagration(model, v_wall, wn, w_center, cs_n, v_cj, vp_guess, wp_guess, t_end, n_xi,
                               thin_shell_limit, allow_failure=True, allow_neg=True)
            return out[11] - wn

        sol = root_scalar(solvable, x0=0.99 * wm_guess, x1=1.01 * wm_guess)
        wm, found = sol.root, sol.converged
        if not found:
            s2 = fsolve_vary(solvable, np.array([wm_guess]))
            wm, found = s2[0][0], s2[2] == 1
        v, w, xi, vp, vm, vp_tilde, vm_tilde, v_sh, vm_sh, vm_tilde_sh, wp, wn_est, wm_sh = deflagration(
            model, v_wall, wn, wm, cs_n, v_cj, vp_guess, wp_guess, t_end, n_xi, thin_shell_limit, allow_neg=True)
        if found and not np.isclose(wn_est, wn, rtol=wn_rtol):
            found = False
    else:
        if v_wall >= v_cj:
            raise RuntimeError("Invalid v_wall for a hybrid")

        def solvable(wm_):
            if isinstance(wm_, np.ndarray):
                wm_ = wm_[0]
            if np.isnan(wm_) or wm_ < 0:
                return np.nan
            out = hybrid(model, v_wall, wn, wm_, cs_n, v_cj, vp_tilde_guess, wp_guess, t_end, n_xi, thin_shell_limit,
                         allow_failure=True, allow_neg=True)
            return out[11] - wn

        s2 = fsolve_vary(solvable, np.array([wm_guess]))
        wm, found = s2[0][0], s2[2] == 1
        if not found:
            # backup scan, fluid.py:731-795: fsolve from every wm_i whose v_plus lies above the shock curve at the wall
            v_sh_wall = v_shock(model, wn, v_wall, cs_n)
            for wm_i in np.linspace(0.3 * wm_guess, 3 * wm_guess, 20):
                vpt_i, _ = solve_junction(model, np.sqrt(model.cs2(wm_i, BROKEN)), wm_i, BROKEN, SYMMETRIC,
                                          vp_tilde_guess, wp_guess)
                vp_i = -lorentz(vpt_i, v_wall)                               # v_plus_hybrid, boundary.py:600-618
                if not vp_i > v_sh_wall:
                    continue
                s3 = fsolve(solvable, x0=wm_i, full_output=True)
                if s3[2] == 1:
                    wm, found = s3[0][0], True
                    break
        v, w, xi, vp, vm, vp_tilde, vm_tilde, v_sh, vm_sh, vm_tilde_sh, wp, wn_est, wm_sh = hybrid(
            model, v_wall, wn, wm, cs_n, v_cj, vp_tilde_guess, wp_guess, t_end, n_xi, thin_shell_limit, allow_neg=True)
        if found and not np.isclose(wn_est, wn, rtol=wn_rtol):
            found = False
        vm = lorentz(v_wall, np.sqrt(model.cs2(wm, BROKEN)))
        vt, wt, xt, _ = sb.integrate(vm, wm, v_wall, BROKEN, -t_end, n_xi, model.css2, model.csb2)
        v = np.concatenate((np.flip(vt), v))
        w = np.concatenate((np.flip(wt), w))
        xi = np.concatenate((np.flip(xt), xi))

    xif = np.linspace(xi[-1] + dxi, 1, 2)
    xib = np.linspace(0, xi[0] - dxi, 2)
    w_center = min(wm, w[0])
    v = np.concatenate((np.zeros(2), v, np.zeros(2)))
    w = np.concatenate((np.ones(2) * w_center, w, np.ones(2) * wn))
    xi = np.concatenate((xib, xi, xif))
    return dict(v=v, w=w, xi=xi, sol_type=sol_type, vp=vp, vm=vm, vp_tilde=vp_tilde, vm_tilde=vm_tilde, v_sh=v_sh,
                vm_sh=vm_sh, vm_tilde_sh=vm_tilde_sh, wp=wp, wm=wm, wm_sh=wm_sh, v_cj=v_cj, solver_failed=not found,
                wn=wn)


# --- bubble.py / props.py / thermo.py ---------------------------------------------------------------------------

def find_phase(xi, v_wall):
    """props.py:22-32."""
    i_wall = sb.find_v_index(xi, v_wall)
    phase = np.zeros_like(xi)
    if i_wall == 0:
        return phase
    phase[:i_wall - 1] = BROKEN
    if np.isclose(xi[i_wall], v_wall):
        phase[i_wall - 1] = BROKEN
    return phase


def bubble_scalars(model: Eos, sol: dict, v_wall):
    """The cached thermodynamic properties of Bubble that the spectrum needs (bubble.py:377-632, thermo.py)."""
    v, w, xi, wn = sol["v"], sol["w"], sol["xi"], sol["wn"]
    phase = find_phase(xi, v_wall)
    xi3 = xi ** 3
    ek_va = 4 * np.pi / 3 * np.trapezoid(w * v ** 2 * gamma2(v), xi3)                       # thermo.py:169-181
    theta = model.theta(w, phase)
    dtheta = 4 * np.pi / 3 * np.trapezoid(theta - model.theta(w[-1], SYMMETRIC), xi3)       # :213-221
    dq = 4 * np.pi / 3 * np.trapezoid(0.75 * (w - w[-1]), xi3)                              # :206-210
    ds = 4 * np.pi / 3 * np.trapezoid(model.s(w, phase) - model.s(w[-1], SYMMETRIC), xi3)   # :156-166
    ebar = float(model.e(wn, SYMMETRIC))                                                     # :89-91
    ek_bva = 3 / (4 * np.pi * v_wall ** 3) * ek_va                                           # :42-46
    eq_va = np.pi * wn * sol["v_sh"] ** 3 - ek_va - dtheta                                   # :191-195
    eq_bva = 3 / (4 * np.pi * v_wall ** 3) * eq_va                                           # :56-60
    w_rev = w[::-1]
    i_max = w.size - int(np.argmax(w_rev != w[-1])) - 1
    if i_max == 0:
        i_max = -1
    wbar = 1 / (xi[i_max] ** 3) * np.trapezoid(w[:i_max + 1], xi[:i_max + 1] ** 3)          # :139-149
    return dict(phase=phase, va_kinetic_energy_density=ek_va, va_trace_anomaly_diff=dtheta,
                va_thermal_energy_density_diff=dq, va_entropy_density_diff=ds, ebar=ebar,
                kinetic_energy_density=ek_bva, kinetic_energy_fraction=ek_bva / ebar,
                va_thermal_energy_density=eq_va, thermal_energy_density=eq_bva, va_enthalpy_density=4 / 3 * eq_bva,
                kappa=ek_va / np.abs(dtheta), omega=dq / np.abs(dtheta), ubarf2=ek_bva / w[-1], wbar=wbar)
