"""Oracle (test infrastructure only): Sound Shell Model spectrum chain, restated from the reference with numpy.

Follows pttools/ssmtools/calculators.py, ssm.py and spectrum.py (line numbers in each docstring).  Loops that the
reference runs under numba ``prange`` are evaluated here in row blocks of plain numpy so that memory stays bounded.
"""
from __future__ import annotations

import numpy as np

from . import shell_bag as sb

Z_ST_THRESH = 50.0            # ssmtools/const.py:32
DZ_ST_BLEND = np.pi           # const.py:35
T_TILDE_MIN, T_TILDE_MAX = 0.01, 20.0   # const.py:38-40
NXI, NT, NZ_LOOKUP = 2000, 10000, 10000  # const.py:10-18
GAMMA = 4.0 / 3.0             # const.py:46
Y_DEFAULT = np.logspace(-1, 3, 1000)   # const.py:20
_BLOCK = 2 ** 21              # elements per numpy temporary


def gradient(f):
    """speedup/functions.py:10-20: central differences, un-halved one-sided ends."""
    out = np.empty_like(f)
    out[1:-1] = (f[2:] - f[:-2]) / 2.0
    out[0] = f[1] - f[0]
    out[-1] = f[-1] - f[-2]
    return out


def logspace(start, stop, num):
    """speedup/functions.py:24-27."""
    return 10.0 ** np.linspace(start, stop, num)


# ---------------------------------------------------------------------------------------------------------------
# a15: sine transform
# ---------------------------------------------------------------------------------------------------------------

def sin_transform_core(xi, f, z):
    """calculators.py:208-228: trapezoid of f*sin(z*xi) for every z."""
    out = np.empty(z.size)
    rows = max(1, _BLOCK // max(1, xi.size))
    for a in range(0, z.size, rows):
        zz = z[a:a + rows]
        out[a:a + rows] = np.trapezoid(f[None, :] * np.sin(zz[:, None] * xi[None, :]), xi, axis=1)
    return out


def envelope(xi, f):
    """calculators.py:17-84."""
    nz = np.nonzero(f)[0]
    xi_nonzero = xi[nz]
    xi1 = np.min(xi_nonzero)
    xi2 = np.max(xi_nonzero)
    ind1 = np.where(xi == xi1)[0][0]
    ind2 = np.where(xi == xi2)[0][0]
    f1 = f[ind1 - 1]
    xi1 = xi[ind1 - 1]
    i_max = int(np.argmax(f))
    f_max = f[i_max]
    xi_w = xi[i_max]
    if i_max + 1 == f.shape[0]:
        df = f[i_max] - f[i_max - 2]
    else:
        df = f[i_max + 1] - f[i_max - 1]
    if df > 0:
        f_m, f_p, f2 = f[i_max - 1], f_max, f[ind2]
    else:
        f_m, f_p, f2 = f_max, 0.0, 0.0
    return (xi1, xi_w, xi_w, xi2), (f1, f_m, f_p, f2)


def sin_transform_approx(z, xi, f):
    """calculators.py:231-258."""
    (xi1, xi_w, _, xi2), (f1, f_m, f_p, f2) = envelope(xi, f)
    out = -(f2 * np.cos(z * xi2) - f_p * np.cos(z * xi_w)) / z
    out += -(f_m * np.cos(z * xi_w) - f1 * np.cos(z * xi1)) / z
    return out


def sin_transform(z, xi, f, z_st_thresh=Z_ST_THRESH):
    """calculators.py:119-158 (array branch; z ascending). ``len(lo) < len(z)`` there compares a 1-tuple's
    length with len(z), so the high-z branch runs whenever len(z) > 1."""
    z = np.asarray(z, dtype=float)
    lo = z <= z_st_thresh
    z_lo = z[lo]
    integral = sin_transform_core(xi, f, z_lo)
    if 1 < z.size:
        hi_sel = z > z_st_thresh - DZ_ST_BLEND
        z_hi = z[hi_sel]
        i_hi = sin_transform_approx(z_hi, xi, f) if z_hi.size else np.empty(0)
        if z_hi.size + z_lo.size > z.size:
            hi_blend = z_hi <= z_st_thresh
            z_hi_blend = z_hi[hi_blend]
            lo_blend = z_lo > z_st_thresh - DZ_ST_BLEND
            zb_max, zb_min = np.max(z_hi_blend), np.min(z_hi_blend)
            if zb_max > zb_min:
                s = (z_hi_blend - zb_min) / (zb_max - zb_min)
            else:
                s = 0.5 * np.ones_like(z_hi_blend)
            frac = 3 * s ** 2 - 2 * s ** 3
            integral[lo_blend] = i_hi[hi_blend] * frac + integral[lo_blend] * (1 - frac)
        integral = np.concatenate((integral, i_hi[z_hi > z_st_thresh]))
    return integral


def resample_uniform_xi(xi, f, n_xi=NXI):
    """calculators.py:87-101."""
    xi_re = np.linspace(0, 1 - 1 / n_xi, n_xi)
    return xi_re, np.interp(xi_re, xi, f)


# ---------------------------------------------------------------------------------------------------------------
# a16: single-bubble A^2
# ---------------------------------------------------------------------------------------------------------------

def a2_from_profile(z, v, w, xi, lam_orig, cs, z_st_thresh=Z_ST_THRESH, nxi=NXI):
    """Common body of a2_e_conserving (ssm.py:35-81) and a2_e_conserving_bag (ssm.py:84-140).
    ``lam_orig`` is (e - e_n)/w_n; the w v^2/w_n term is added here as in ssm.py:62 / :127."""
    f = (4.0 * np.pi / z) * sin_transform(z, xi, v, z_st_thresh)
    v_ft = gradient(f) / gradient(z)
    lam = lam_orig + w * v * v / w[-1]
    xi_re, lam_re = resample_uniform_xi(xi, lam, nxi)
    lam_ft = (4.0 * np.pi / z) * sin_transform(z, xi_re, xi_re * lam_re, z_st_thresh)
    a2 = 0.25 * (v_ft ** 2 + (cs * lam_ft) ** 2)
    return a2, v_ft ** 2 / 2, (cs * lam_ft) ** 2 / 2


def a2_e_conserving_bag(z, v_wall, alpha_n, npt=(NXI, NT, NZ_LOOKUP), z_st_thresh=Z_ST_THRESH, profile=None):
    """ssm.py:84-140 with de_method=standard."""
    v, w, xi = profile if profile is not None else sb.sound_shell_bag(v_wall, alpha_n, npt[0])
    lam_orig = sb.de_from_w_bag(w, xi, v_wall, alpha_n) / w[-1]
    return a2_from_profile(z, v, w, xi, lam_orig, sb.CS0, z_st_thresh, npt[0])


# ---------------------------------------------------------------------------------------------------------------
# a17: velocity spectral density
# ---------------------------------------------------------------------------------------------------------------

def nu(t, nuc_type="exponential", a=1.0):
    """spectrum.py:189-205."""
    if nuc_type == "simultaneous":
        return 0.5 * a * (a * t) ** 2 * np.exp(-(a * t) ** 3 / 6)
    if nuc_type == "exponential":
        return a * np.exp(-a * t)
    raise ValueError("Nucleation type not recognized")


def qt_lookup_grid(z):
    """spectrum.py:503-515."""
    lmin, lmax = np.log10(np.min(z)), np.log10(np.max(z))
    dl = (lmax - lmin) / z.size
    return 10 ** np.arange(lmin + np.log10(T_TILDE_MIN), lmax + np.log10(T_TILDE_MAX), dl)


def spec_den_v_core(z, a2_lookup, qt_lookup, v_wall, nuc_type="exponential", a=1.0, nt=NT):
    """spectrum.py:451-486."""
    t = logspace(np.log10(T_TILDE_MIN), np.log10(T_TILDE_MAX), nt)
    b_r = (8.0 * np.pi) ** (1.0 / 3.0)
    factor = 2.0 / (b_r * v_wall) ** 6
    wgt = t ** 6 * nu(t, nuc_type, a)
    out = np.empty(z.size)
    rows = max(1, _BLOCK // nt)
    for s in range(0, z.size, rows):
        q = z[s:s + rows, None] * t[None, :] / (b_r * v_wall)
        arr = wgt[None, :] * np.interp(q, qt_lookup, a2_lookup)
        out[s:s + rows] = np.trapezoid(arr, t, axis=1) * factor
    return out


def spec_den_v_bag(z, params, npt=(NXI, NT, NZ_LOOKUP), z_st_thresh=Z_ST_THRESH, profile=None):
    """spectrum.py:536-603 (method e_conserving, no file)."""
    v_wall, alpha_n = params[0], params[1]
    nuc_type = params[2] if len(params) > 2 else "exponential"
    nuc_args = params[3] if len(params) > 3 else (1,)
    qt = qt_lookup_grid(z)
    a2 = a2_e_conserving_bag(qt, v_wall, alpha_n, npt, z_st_thresh, profile=profile)[0]
    return spec_den_v_core(z, a2, qt, v_wall, nuc_type, nuc_args[0], npt[1])


# ---------------------------------------------------------------------------------------------------------------
# a18/a19: GW spectral density and power
# ---------------------------------------------------------------------------------------------------------------

def lookup_limits(y, cs, eps=0.0):
    """spectrum.py:183-186."""
    return y.min() * 0.5 * (1.0 - cs) / cs - eps, y.max() * 0.5 * (1.0 + cs) / cs + eps


def gen_lookup(y, cs, n_z_lookup=NZ_LOOKUP, eps=0.0):
    """spectrum.py:169-180."""
    lo, hi = lookup_limits(y, cs, eps)
    return logspace(np.log10(lo), np.log10(hi), n_z_lookup)


def spec_den_gw_scaled(z_lookup, p_v_lookup, y=None, cs=sb.CS0, gamma=GAMMA, source_lifetime_factor=1.0, nz_int=None):
    """spectrum.py:327-430."""
    if y is None:
        zmax = z_lookup.max() * 2.0 * cs / (1.0 + cs)
        zmin = z_lookup.min() * 2.0 * cs / (1.0 - cs)
        y = logspace(np.log10(zmin), np.log10(zmax), z_lookup.size)
    else:
        lo, hi = lookup_limits(y, cs)
        if z_lookup.max() < hi or z_lookup.min() > lo:
            raise ValueError("Range of z_lookup is not large enough.")
    n_int = z_lookup.size if nz_int is None else nz_int
    cs2 = cs ** 2
    zpf = (1 + cs) / (2 * cs)
    zmf = (1 - cs) / (2 * cs)
    pref = ((1 - cs2) / cs2) ** 2 / (4 * np.pi * cs)
    p_gw = np.zeros_like(y)
    for i in range(y.size):
        z_plus = y[i] * zpf
        z_minus = y[i] * zmf
        z = logspace(np.log10(z_minus), np.log10(z_plus), n_int)
        integrand = ((z - z_plus) ** 2 * (z - z_minus) ** 2) / (z * (z_plus + z_minus - z)) \
            * np.interp(z, z_lookup, p_v_lookup) \
            * np.interp(z_plus + z_minus - z, z_lookup, p_v_lookup)
        p_gw[i] = pref / y[i] * np.trapezoid(integrand, z)
    return 0.75 * gamma ** 2 * p_gw * source_lifetime_factor, y


def pow_spec(z, spec_den):
    """spectrum.py:230-240."""
    return z ** 3 / (2.0 * np.pi ** 2) * spec_den


def power_gw_scaled_bag(z, params, npt=(NXI, NT, NZ_LOOKUP), z_st_thresh=Z_ST_THRESH, profile=None):
    """spectrum.py:243-294."""
    z = np.asarray(z, dtype=float)
    if np.any(z <= 0.0):
        raise ValueError("z values must all be positive.")
    eps = 1e-8
    xmax = max(z) * (0.5 * (1.0 + sb.CS0) / sb.CS0) + eps
    xmin = min(z) * (0.5 * (1.0 - sb.CS0) / sb.CS0) - eps
    x = np.logspace(np.log10(xmin), np.log10(xmax), npt[2])
    sd_v = spec_den_v_bag(x, params, npt, z_st_thresh, profile=profile)
    sd_gw, _ = spec_den_gw_scaled(x, sd_v, z)
    return pow_spec(z, sd_gw)


def power_v_bag(z, params, npt=(NXI, NT, NZ_LOOKUP), z_st_thresh=Z_ST_THRESH, profile=None):
    """spectrum.py:297-324."""
    return pow_spec(z, spec_den_v_bag(z, params, npt, z_st_thresh, profile=profile))


# ---------------------------------------------------------------------------------------------------------------
# a16-a20 for Bubble objects: SSMSpectrum.compute and Spectrum.omgw0
# ---------------------------------------------------------------------------------------------------------------

def a2_e_conserving(model, sol, scal, z, cs, z_st_thresh=Z_ST_THRESH, nxi=NXI):
    """ssm.py:35-81 for a solved bubble (sol: profile dict of shell_generic.sound_shell_generic, scal: bubble_scalars)."""
    v, w, xi = sol["v"], sol["w"], sol["xi"]
    e = model.e(w, scal["phase"])
    lam_orig = (e - e[-1]) / w[-1]
    return a2_from_profile(z, v, w, xi, lam_orig, cs, z_st_thresh, nxi)


def ssm_spectrum(model, sol, scal, v_wall, y=None, z_st_thresh=Z_ST_THRESH, nuc_type="exponential", nt=NT,
                 n_z_lookup=NZ_LOOKUP, r_star=None, lifetime_multiplier=1.0):
    """SSMSpectrum.compute (spectrum.py:96-136). Returns dict(a2, cs, spec_den_v, pow_v, z_lookup, spec_den_gw, pow_gw)."""
    y = Y_DEFAULT if y is None else y
    cs = float(np.sqrt(model.cs2(scal["va_enthalpy_density"], 1.0)))
    qt = qt_lookup_grid(y)
    a2 = a2_e_conserving(model, sol, scal, qt, cs, z_st_thresh)[0]
    sdv = spec_den_v_core(y, a2, qt, v_wall, nuc_type, 1.0, nt)
    pow_v = pow_spec(y, sdv)
    z_lookup = gen_lookup(y, cs, n_z_lookup, eps=1e-8)
    qt2 = qt_lookup_grid(z_lookup)
    a2_2 = a2_e_conserving(model, sol, scal, qt2, cs, z_st_thresh)[0]
    sdv2 = spec_den_v_core(z_lookup, a2_2, qt2, v_wall, nuc_type, 1.0, nt)
    nu_ = float(model.nu_gdh2024(scal["va_enthalpy_density"]))
    slf = 1 / (1 + 2 * nu_)
    if r_star is not None and not np.isinf(lifetime_multiplier):
        slf *= 1 - (1 + lifetime_multiplier * 2 * r_star / np.sqrt(scal["kinetic_energy_fraction"])) ** (-1 - 2 * nu_)
    sd_gw, _ = spec_den_gw_scaled(z_lookup, sdv2, y, cs, source_lifetime_factor=slf)
    return dict(a2=a2, cs=cs, spec_den_v=sdv, pow_v=pow_v, z_lookup=z_lookup, spec_den_gw=sd_gw,
                pow_gw=pow_spec(y, sd_gw), source_lifetime_factor=slf)


def omgw0_scale(r_star, suppression_factor, g_star=100.0, gs_star=100.0, g0=2.0, gs0=3.91):
    """r_* F_gw0 Sigma of Spectrum.omgw0 (omgw0_ssm.py:123-131, 191-204; const.py:10-40)."""
    om_gamma0 = 3.57e-5 * gs0 ** (4 / 3) / g0
    return r_star * om_gamma0 * (gs0 / gs_star) ** (4 / 3) * g_star / g0 * suppression_factor
