"""pttools.ssmtools call surface on top of the CUDA kernels (calculators.py, ssm.py, spectrum.py of the reference).

The batched operators live in `engine` / `sweep`; the functions here are one-profile batches with the reference's
names, arguments and error behaviour."""
from __future__ import annotations

import enum
import typing as tp

import numpy as np

from . import bubble as bub_mod
from . import engine
from .bubble import Bubble
from .models import Phase

NXIDEFAULT, NTDEFAULT, N_Z_LOOKUP_DEFAULT = 2000, 10000, 10000       # ssmtools/const.py:10-18
NPTDEFAULT = (NXIDEFAULT, NTDEFAULT, N_Z_LOOKUP_DEFAULT)
Y_DEFAULT = np.logspace(-1, 3, 1000)
Z_ST_THRESH = 50
DZ_ST_BLEND = np.pi
T_TILDE_MAX, T_TILDE_MIN = 20.0, 0.01
CS0 = bub_mod.CS0
GAMMA = 4 / 3


@enum.unique
class NucType(str, enum.Enum):
    """spectrum.py:25-29."""
    EXPONENTIAL = "exponential"
    SIMULTANEOUS = "simultaneous"


DEFAULT_NUC_TYPE = NucType.EXPONENTIAL


def _nuc(nuc_type) -> str:
    return nuc_type.value if isinstance(nuc_type, NucType) else str(nuc_type)


def _one_profile(v, w, xi):
    torch = engine._torch()
    rows = [engine.dev_f64(np.asarray(a, dtype=float))[None, :].contiguous() for a in (v, w, xi)]
    off = torch.zeros(1, dtype=torch.int32, device="cuda")
    npts = torch.full((1,), rows[0].shape[1], dtype=torch.int32, device="cuda")
    return engine.ShellBatch(None, None, None, None, None, rows[0], rows[1], rows[2], off, npts, None, 0)


def sin_transform(z, xi: np.ndarray, f: np.ndarray, z_st_thresh: float = Z_ST_THRESH, v_wall=None, v_sh=None):
    """calculators.py:161-190."""
    scalar = np.isscalar(z)
    zz = np.atleast_1d(np.asarray(z, dtype=float))
    out = engine.sin_transform_batch(zz, np.asarray(xi, dtype=float), np.asarray(f, dtype=float), [0], [len(xi)],
                                     z_st_thresh)[0].cpu().numpy()
    return float(out[0]) if scalar else out


def gen_lookup(y: np.ndarray, cs: float, n_z_lookup: int = N_Z_LOOKUP_DEFAULT, eps: float = 0.) -> np.ndarray:
    """spectrum.py:169-186."""
    lo = y.min() * 0.5 * (1. - cs) / cs - eps
    hi = y.max() * 0.5 * (1. + cs) / cs + eps
    return 10.0 ** np.linspace(np.log10(lo), np.log10(hi), n_z_lookup)


def nu(T, nuc_type=NucType.SIMULTANEOUS, a: float = 1.):
    """spectrum.py:189-205."""
    return engine.nu(np.asarray(T, dtype=float), _nuc(nuc_type), a)


def pow_spec(z, spec_den):
    """spectrum.py:230-240 (elementwise; host arrays in, host arrays out)."""
    return z ** 3 / (2. * np.pi ** 2) * spec_den


def _bubble_pp(b: Bubble):
    m = b.model
    return engine.point_params([b.v_wall], 0.0, 1 - 1 / m.mu_s, 1 - 1 / m.mu_b, m.V_s, m.V_b)


def a2_e_conserving(bub: Bubble, z: np.ndarray, cs: float, z_st_thresh: float = Z_ST_THRESH, nxi: int = NXIDEFAULT):
    """ssm.py:35-81: (A2, fp2_2, lam2)."""
    if not bub.solved:
        bub.solve()
    plan = engine.A2OnlyPlan(np.asarray(z, dtype=float), nxi=nxi, z_st_thresh=z_st_thresh, phase_rule=1)
    pp = _bubble_pp(bub)
    pp[:, 1] = cs
    out = engine.a2_batch(plan, _one_profile(bub.v, bub.w, bub.xi), pp, want_parts=True)
    return tuple(o[0].cpu().numpy() for o in out)


def a2_e_conserving_bag(z: np.ndarray, v_wall: float, alpha_n: float, npt=NPTDEFAULT, z_st_thresh: float = Z_ST_THRESH,
                        v_ip=None, w_ip=None, xi=None, **_):
    """ssm.py:84-140 (de_method standard)."""
    if v_ip is None or np.size(v_ip) <= 1:
        v_ip, w_ip, xi = bub_mod.sound_shell_bag(v_wall, alpha_n, npt[0])
    plan = engine.A2OnlyPlan(np.asarray(z, dtype=float), nxi=npt[0], z_st_thresh=z_st_thresh, phase_rule=0)
    pp = engine.point_params([v_wall], CS0, 0.75, 0.75, 0.75 * w_ip[-1] * alpha_n, 0.0)
    out = engine.a2_batch(plan, _one_profile(v_ip, w_ip, xi), pp, want_parts=True)
    return tuple(o[0].cpu().numpy() for o in out)


def spec_den_v(bub: Bubble, z: np.ndarray, a: float, nuc_type, nt: int = NTDEFAULT, z_st_thresh: float = Z_ST_THRESH,
               cs: float = None, return_a2: bool = False):
    """spectrum.py:489-533."""
    if not bub.solved:
        bub.solve()
    z = np.asarray(z, dtype=float)
    plan = engine.SsmPlan(z, cs=cs, nt=nt, n_z_lookup=4, z_st_thresh=z_st_thresh, nuc_type=_nuc(nuc_type), a=a,
                          mode="ssm", only_first_pass=True)
    pp = _bubble_pp(bub)
    pp[:, 1] = cs
    a2 = engine.a2_batch(plan, _one_profile(bub.v, bub.w, bub.xi), pp)
    sdv = engine.spec_den_v_batch(plan, 0, a2, pp[:, 0].contiguous())[0].cpu().numpy()
    if return_a2:
        return sdv, a2[0].cpu().numpy()
    return sdv


def _parse_params(params):
    """spectrum.py:209-227."""
    v_wall, alpha = params[0], params[1]
    nuc_type = params[2] if len(params) > 2 else DEFAULT_NUC_TYPE
    nuc_args = params[3] if len(params) > 3 else (1,)
    return v_wall, alpha, nuc_type, nuc_args


def _check_physical(params):
    v_wall, alpha_n = params[0], params[1]
    if not 0.0 <= v_wall <= 1.0:
        raise ValueError("Unphysical parameter: v_wall. See the log for details.")
    if alpha_n > bub_mod.alpha_n_max_bag(float(v_wall)):
        raise ValueError("Unphysical parameter(s). See the log for details.")


def spec_den_v_bag(z: np.ndarray, params, npt=NPTDEFAULT, z_st_thresh=Z_ST_THRESH, **_):
    """spectrum.py:536-603 (method e_conserving, no file)."""
    _check_physical(params)
    v_wall, alpha, nuc_type, nuc_args = _parse_params(tuple(params))
    z = np.asarray(z, dtype=float)
    plan = engine.SsmPlan(z, cs=CS0, nt=npt[1], n_z_lookup=4, nxi=npt[0], z_st_thresh=z_st_thresh,
                          nuc_type=_nuc(nuc_type), a=nuc_args[0], mode="ssm", phase_rule=0, only_first_pass=True)
    shells = engine.sweep_shells_bag([v_wall], [alpha], n_xi=npt[0])
    pp = engine.point_params([v_wall], CS0, 0.75, 0.75, 0.75 * alpha, 0.0)
    a2 = engine.a2_batch(plan, shells, pp)
    return engine.spec_den_v_batch(plan, 0, a2, pp[:, 0].contiguous())[0].cpu().numpy()


def spec_den_gw_scaled(z_lookup: np.ndarray, P_v_lookup: np.ndarray, y: np.ndarray = None, cs: float = CS0,
                       Gamma: float = GAMMA, source_lifetime_factor: float = 1., nz_int: int = None):
    """spectrum.py:402-430.  Any ascending z_lookup: the log-uniform grids the reference itself passes
    (gen_lookup / np.logspace) take the arithmetic-index kernel, anything else np.interp by binary search."""
    z_lookup = np.asarray(z_lookup, dtype=float)
    P_v_lookup = np.asarray(P_v_lookup, dtype=float)
    if z_lookup.shape != P_v_lookup.shape:
        raise TypeError("z_lookup and P_v_lookup must be of the same shape.")
    if y is None:
        zmax = z_lookup.max() * 2. * cs / (1. + cs)
        zmin = z_lookup.min() * 2. * cs / (1. - cs)
        y = 10.0 ** np.linspace(np.log10(zmin), np.log10(zmax), z_lookup.size)
    elif isinstance(y, np.ndarray):
        if z_lookup.max() < y.max() * 0.5 * (1. + cs) / cs or z_lookup.min() > y.min() * 0.5 * (1. - cs) / cs:
            raise ValueError("Range of z_lookup is not large enough.")
    else:
        raise TypeError(f"Unknown type for z: {type(y)}")
    plan = engine.GwOnlyPlan(z_lookup, y, cs=cs, gamma=Gamma, nz_int=nz_int)
    torch = engine._torch()
    sd, _, _ = engine.spec_den_gw_batch(plan, engine.dev_f64(P_v_lookup)[None, :].contiguous(),
                                        slf=[source_lifetime_factor])
    return sd[0].cpu().numpy(), y


def power_gw_scaled_bag(z: np.ndarray, params, npt=NPTDEFAULT, z_st_thresh: float = Z_ST_THRESH, **_):
    """spectrum.py:243-294."""
    z = np.asarray(z, dtype=float)
    if np.any(z <= 0.0):
        raise ValueError("z values must all be positive.")
    _check_physical(params)
    v_wall, alpha, nuc_type, nuc_args = _parse_params(tuple(params))
    from . import sweep
    res = sweep.power_gw_scaled_bag_sweep([v_wall], [alpha], z, npt, _nuc(nuc_type), nuc_args, z_st_thresh)
    return res.pow_gw[0].cpu().numpy()


def power_v_bag(z: np.ndarray, params, npt=NPTDEFAULT, z_st_thresh: float = Z_ST_THRESH, **_):
    """spectrum.py:297-324."""
    return pow_spec(np.asarray(z, dtype=float), spec_den_v_bag(z, params, npt, z_st_thresh))


class SSMSpectrum:
    """spectrum.py:36-158."""

    def __init__(self, bubble: Bubble, y: np.ndarray = None, z_st_thresh: float = Z_ST_THRESH,
                 nuc_type: NucType = DEFAULT_NUC_TYPE, nt: int = NTDEFAULT, n_z_lookup: int = N_Z_LOOKUP_DEFAULT,
                 r_star: float = None, lifetime_multiplier: float = 1, compute: bool = True, label_latex: str = None,
                 label_unicode: str = None):
        self.bubble = bubble
        self.nuc_type = nuc_type
        self.y = Y_DEFAULT if y is None else np.asarray(y, dtype=float)
        self.z_st_thresh, self.nt, self.n_z_lookup = z_st_thresh, nt, n_z_lookup
        self.r_star, self.lifetime_multiplier = r_star, lifetime_multiplier
        self.label_latex = label_latex or bubble.label_latex
        self.label_unicode = label_unicode or bubble.label_unicode
        self.a2 = self.cs = self.spec_den_v = self.spec_den_gw = self.pow_v = self.pow_gw = self.z_lookup = None
        if compute:
            self.compute()

    @property
    def source_lifetime_factor(self) -> float:
        """spectrum.py:123-136."""
        b = self.bubble
        nu_ = float(b.model.nu_gdh2024(b.va_enthalpy_density))
        ret = 1 / (1 + 2 * nu_)
        if self.r_star is not None and not np.isinf(self.lifetime_multiplier):
            ret *= 1 - (1 + self.lifetime_multiplier * 2 * self.r_star / np.sqrt(b.kinetic_energy_fraction)) ** (-1 - 2 * nu_)
        return ret

    def compute(self):
        b = self.bubble
        if not b.solved:
            b.solve()
        self.cs = float(np.sqrt(b.model.cs2(b.va_enthalpy_density, Phase.BROKEN.value)))
        plan = engine.SsmPlan(self.y, cs=self.cs, nt=self.nt, n_z_lookup=self.n_z_lookup, z_st_thresh=self.z_st_thresh,
                              nuc_type=_nuc(self.nuc_type), a=1.0, mode="ssm")
        pp = _bubble_pp(b)
        pp[:, 1] = self.cs
        res = engine.spectra_from_shells(plan, _one_profile(b.v, b.w, b.xi), pp, slf=[self.source_lifetime_factor],
                                         keep=True)
        self.a2 = res.a2[0].cpu().numpy()
        self.spec_den_v = res.sd_v[0].cpu().numpy()
        self.pow_v = res.pow_v[0].cpu().numpy()
        self.z_lookup = plan.z_lookup
        self.spec_den_gw = res.sd_gw[0].cpu().numpy()
        self.pow_gw = res.pow_gw[0].cpu().numpy()
