"""pttools.omgw0 call surface: gravitational-wave power spectrum today (omgw0_ssm.py, suppression.py, const.py).

Everything here is elementwise post-processing of `pow_gw`; in sweeps the per-point scale r_* F_gw0 Sigma is folded
into the last kernel (ptt_spec_den_gw_batch's `scale`)."""
from __future__ import annotations

import enum
import os

import numpy as np

from . import noise
from . import ssmtools as ssm
from .bubble import Bubble
from .models import Phase

G_STAR_DEFAULT = 100            # omgw0/const.py:6
T_default = 100                 # const.py:7  (GeV)
Fgw0 = 3.57e-5                  # const.py:10
fs0_ref = 2.6e-6                # const.py:13
G0 = 2                          # const.py:19
GS0 = 3.91                      # const.py:21
OMEGA_RADIATION = Fgw0 * GS0 ** (4 / 3) / G0      # const.py:40

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class SuppressionMethod(enum.StrEnum):
    """suppression.py:19-22."""
    NONE = "none"
    NO_EXT = DEFAULT = "no_ext"
    EXT_CONSTANT = "ext_constant"


def alpha_n_max_approx(vw):
    """suppression.py:99-103."""
    return 1 / 3 * (1 + 3 * vw ** 2) / (1 - vw ** 2)


class Suppression:
    """suppression.py:25-96: linear (Delaunay) interpolation of the simulation-measured suppression factor."""

    def __init__(self, v_walls, alpha_ns, suppressions, name=None):
        if not (v_walls.size == alpha_ns.size == suppressions.size):
            raise ValueError("Input arrays must have the same size.")
        self.v_walls, self.alpha_ns, self.suppressions, self.name = v_walls, alpha_ns, suppressions, name
        self.alpha_n_min = float(np.min(alpha_ns))
        self._interp = None

    @classmethod
    def from_file(cls, path, name=None):
        d = np.load(path)
        return cls(d["vw_sim"], d["alpha_sim"], d["sup_ssm"], name)

    def _linear(self):
        if self._interp is None:
            from scipy.interpolate import LinearNDInterpolator       # what griddata(method="linear") builds
            self._interp = LinearNDInterpolator(np.stack([self.v_walls, self.alpha_ns], axis=1), self.suppressions)
        return self._interp

    def suppression(self, v_wall, alpha_n, method=SuppressionMethod.DEFAULT, interpolation="linear"):
        if interpolation != "linear":
            raise NotImplementedError("only the reference's default interpolation='linear' is implemented")
        scalar = np.isscalar(v_wall) and np.isscalar(alpha_n)
        if method == SuppressionMethod.NONE:
            return 1. if scalar else np.ones(np.broadcast(v_wall, alpha_n).shape)
        if method not in (SuppressionMethod.NO_EXT, SuppressionMethod.EXT_CONSTANT):
            raise ValueError(f"Got invalid suppression method: {method}")
        if scalar:
            sup = self._linear()(np.array([[v_wall, alpha_n]]))[0]
            if method == SuppressionMethod.EXT_CONSTANT and alpha_n < self.alpha_n_min:
                sup = 1.0
            return sup
        vw_m, an_m = np.meshgrid(v_wall, alpha_n)
        sup = self._linear()(np.stack([vw_m.ravel(), an_m.ravel()], axis=1)).reshape(vw_m.shape)
        if method == SuppressionMethod.EXT_CONSTANT:
            sup[an_m < self.alpha_n_min] = 1
        sup[an_m > alpha_n_max_approx(vw_m)] = np.nan
        return sup

    def _triangles(self):
        """The Delaunay triangulation griddata builds (scipy.spatial.Delaunay = Qhull), as flat vertex / value tables."""
        if getattr(self, "_tri", None) is None:
            from scipy.spatial import Delaunay
            pts = np.stack([self.v_walls, self.alpha_ns], axis=1)
            simplices = Delaunay(pts).simplices
            self._tri = (np.ascontiguousarray(pts[simplices].reshape(-1, 6)), np.ascontiguousarray(self.suppressions[simplices]))
        return self._tri

    def points_device(self, v_wall, alpha_n, method=SuppressionMethod.DEFAULT):
        """points() on the device (ptt_suppression_batch): device tensors in, device tensor out."""
        import ctypes as C
        from . import _lib, engine
        torch = engine._torch()
        v_wall, alpha_n = engine.dev_f64(v_wall).reshape(-1), engine.dev_f64(alpha_n).reshape(-1)
        if method == SuppressionMethod.NONE:
            return torch.ones_like(v_wall)
        if method not in (SuppressionMethod.NO_EXT, SuppressionMethod.EXT_CONSTANT):
            raise ValueError(f"Got invalid suppression method: {method}")
        if getattr(self, "_tri_dev", None) is None:
            tri, val = self._triangles()
            self._tri_dev = (engine.dev_f64(tri), engine.dev_f64(val))
        tri, val = self._tri_dev
        out = torch.empty_like(v_wall)
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(_lib.load().ptt_suppression_batch(v_wall.numel(), p(v_wall), p(alpha_n), tri.shape[0], p(tri), p(val),
                                                     int(method == SuppressionMethod.EXT_CONSTANT), float(self.alpha_n_min),
                                                     p(out), engine._stream_ptr(torch)), "ptt_suppression_batch")
        return out

    def points(self, v_wall, alpha_n, method=SuppressionMethod.DEFAULT):
        """Pointwise version for sweeps: suppression factor for matching arrays of (v_wall, alpha_n)."""
        v_wall, alpha_n = np.asarray(v_wall, dtype=float), np.asarray(alpha_n, dtype=float)
        if method == SuppressionMethod.NONE:
            return np.ones_like(v_wall)
        sup = self._linear()(np.stack([v_wall.ravel(), alpha_n.ravel()], axis=1)).reshape(v_wall.shape)
        if method == SuppressionMethod.EXT_CONSTANT:
            sup = np.where(alpha_n < self.alpha_n_min, 1.0, sup)
        return sup


def _extend(sup: Suppression) -> Suppression:
    """suppression.py:121-144: one extra low-alpha point at v_w = 0.24, linear extrapolation of the 0.24 column."""
    a = np.array([0.05000, 0.07300, 0.11000, 0.16000, 0.23000, 0.34000])
    s = np.array([0.01675, 0.01218, 0.00696, 0.00251, 0.00054, 0.00007])
    ext = s[0] + (s[1] - s[0]) / (a[1] - a[0]) * (0.005 - a[0])          # k=1 spline, ext=0 => linear extrapolation
    return Suppression(np.concatenate(([0.24], sup.v_walls)), np.concatenate(([0.005], sup.alpha_ns)),
                       np.concatenate(([ext], sup.suppressions)), name="No hybrids, extended")


NO_HYBRIDS = Suppression.from_file(os.path.join(_DATA, "suppression_no_hybrids_ssm.npz"), name="No hybrids")
NO_HYBRIDS_EXT = _extend(NO_HYBRIDS)
WITH_HYBRIDS = Suppression.from_file(os.path.join(_DATA, "suppression_2_ssm.npz"), name="With hybrids")
DEFAULT_SUPPRESSION = NO_HYBRIDS_EXT


def f(z, r_star, f_star0):
    """omgw0_ssm.py:162-171."""
    return z / r_star * f_star0


def f_star0(Tn, g_star=100):
    """omgw0_ssm.py:179-188."""
    return fs0_ref * (Tn / 100) * (g_star / 100) ** (1 / 6)


def f0(rs, T_n=T_default, g_star=100):
    return f_star0(T_n, g_star) / rs


def F_gw0(g_star, g0=G0, gs0=GS0, gs_star=None, om_gamma0=OMEGA_RADIATION):
    """omgw0_ssm.py:191-204."""
    if g0 is None or gs0 is None or gs_star is None or om_gamma0 is None:
        return 3.57e-5 * (100 / g_star) ** (1 / 3)
    return om_gamma0 * (gs0 / gs_star) ** (4 / 3) * g_star / g0


def J(r_star, K_frac, nu=0):
    """omgw0_ssm.py:207-217."""
    return r_star * (1 - (np.sqrt(1 + 2 * r_star / np.sqrt(K_frac)) ** (-1 - 2 * nu)))


class Spectrum(ssm.SSMSpectrum):
    """omgw0_ssm.py:24-159."""

    def __init__(self, bubble: Bubble, y=None, z_st_thresh=ssm.Z_ST_THRESH, nuc_type=ssm.DEFAULT_NUC_TYPE,
                 nt=ssm.NTDEFAULT, n_z_lookup=ssm.N_Z_LOOKUP_DEFAULT, r_star=None, lifetime_multiplier=1, compute=True,
                 label_latex=None, label_unicode=None, Tn=None, g_star=None, gs_star=None):
        super().__init__(bubble, y, z_st_thresh, nuc_type, nt, n_z_lookup, r_star, lifetime_multiplier, compute,
                         label_latex, label_unicode)
        self.override_necessary = not self.bubble.model.temperature_is_physical
        self.Tn_manual_override = Tn is not None
        self.g_star_manual_override = g_star is not None
        self.gs_star_manual_override = gs_star is not None
        self.Tn_override = T_default if Tn is None else Tn
        self.g_star_override = G_STAR_DEFAULT if g_star is None else gs_star      # sic: omgw0_ssm.py:74
        self.gs_star_override = self.g_star_override if gs_star is None else gs_star

    @property
    def g_star(self):
        if self.override_necessary or self.gs_star_manual_override:                # sic: omgw0_ssm.py:99
            return self.g_star_override
        return self.bubble.model.gp(w=self.bubble.va_enthalpy_density, phase=Phase.BROKEN.value)

    @property
    def gs_star(self):
        if self.override_necessary or self.gs_star_manual_override:
            return self.gs_star_override
        return self.bubble.model.gs(w=self.bubble.va_enthalpy_density, phase=Phase.BROKEN.value)

    @property
    def Tn(self):
        if self.override_necessary or self.Tn_manual_override:
            return self.Tn_override
        return self.bubble.Tn

    @property
    def f_star0(self):
        return f_star0(Tn=self.Tn, g_star=self.g_star)

    def f(self, z=None):
        return f(z=self.y if z is None else z, r_star=self.r_star, f_star0=self.f_star0)

    def F_gw0(self, g0=G0, gs0=GS0):
        return F_gw0(g_star=self.g_star, g0=g0, gs0=gs0, gs_star=self.gs_star)

    def suppression_factor(self, suppression=DEFAULT_SUPPRESSION, method=SuppressionMethod.DEFAULT):
        return suppression.suppression(v_wall=self.bubble.v_wall, alpha_n=self.bubble.alpha_n, method=method)

    def omgw0(self, g0=G0, gs0=GS0, sup=DEFAULT_SUPPRESSION, sup_method=SuppressionMethod.DEFAULT):
        """omgw0_ssm.py:123-131."""
        return self.r_star * self.F_gw0(g0=g0, gs0=gs0) * self.pow_gw * self.suppression_factor(sup, sup_method)

    def omgw0_peak(self, **kw):
        om = self.omgw0(**kw)
        i = int(np.argmax(om))
        return self.f()[i], om[i]

    def noise(self):
        """omgw0_ssm.py:117-118."""
        return noise.omega_noise(self.f())

    def noise_ins(self):
        """omgw0_ssm.py:120-121."""
        return noise.omega_ins(self.f())

    def signal_to_noise_ratio(self) -> float:
        """omgw0_ssm.py:143-144."""
        return float(noise.signal_to_noise_ratio(f=self.f(), signal=self.omgw0(), noise=self.noise()))

    def signal_to_noise_ratio_instrument(self) -> float:
        """omgw0_ssm.py:146-147."""
        return float(noise.signal_to_noise_ratio(f=self.f(), signal=self.omgw0(), noise=self.noise_ins()))
