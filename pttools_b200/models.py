"""Host-side equations of state with the reference's Model interface (pttools/models/{model,analytic,bag,const_cs}.py).

Only the two models named by the hot path are provided.  Both are "mu-nu" models,
    p = a (T/T0)^(mu-4) T^4 - V,   w = mu a (T/T0)^(mu-4) T^4,   e = w - p,   mu = 1 + 1/cs2  (per phase),
the bag model being mu_s = mu_b = 4.  Everything here is closed-form scalar/array algebra that the reference
also evaluates on the host once per model or per parameter point; the device receives the few numbers it needs
(cs2, mu, V per phase) as a POD (ptt_eos_t).
"""
from __future__ import annotations

import enum
import logging

import warnings

import numpy as np

logger = logging.getLogger(__name__)


@enum.unique
class Phase(float, enum.Enum):
    """pttools/bubble/boundary.py:27-33."""
    SYMMETRIC = 0.
    BROKEN = 1.


@enum.unique
class SolutionType(str, enum.Enum):
    """pttools/bubble/boundary.py:37-62."""
    DETON = "Detonation"
    ERROR = "Error"
    HYBRID = "Hybrid"
    SUB_DEF = "Subsonic deflagration"
    UNKNOWN = "Unknown"


SOL_CODE = {SolutionType.SUB_DEF: 0, SolutionType.HYBRID: 1, SolutionType.DETON: 2, SolutionType.ERROR: -1,
            SolutionType.UNKNOWN: -1}
CODE_SOL = {0: SolutionType.SUB_DEF, 1: SolutionType.HYBRID, 2: SolutionType.DETON, -1: SolutionType.ERROR}


class Model:
    """Common implementation; see BagModel / ConstCSModel for the constructors of the reference."""
    DEFAULT_NAME = None
    TEMPERATURE_IS_PHYSICAL = False
    DEFAULT_T_MIN = 1e-3            # models/base.py:22

    def __init__(self, css2, csb2, V_s=1.0, V_b=0.0, a_s=None, a_b=None, T_ref=1.0, T_min=None, name=None,
                 label_latex=None, label_unicode=None, allow_invalid=False):
        if a_b is None:
            a_b = 1.0                                   # analytic.py:184-189
        if a_s is None:
            a_s = 1.1 * a_b
        if V_s < V_b and not allow_invalid:
            raise ValueError(f"The bubble will not expand, when V_s < V_b. Got: V_s={V_s}, V_b={V_b}.")
        self.css2, self.csb2 = float(css2), float(csb2)
        self.css, self.csb = np.sqrt(self.css2), np.sqrt(self.csb2)
        self.mu_s, self.mu_b = 1 + 1 / self.css2, 1 + 1 / self.csb2
        self.V_s, self.V_b, self.a_s, self.a_b = float(V_s), float(V_b), float(a_s), float(a_b)
        self.T_ref = float(T_ref)
        self.T_min = self.DEFAULT_T_MIN if T_min is None else T_min
        self.T_max = np.inf
        self.name = self.DEFAULT_NAME if name is None else name
        self.temperature_is_physical = self.TEMPERATURE_IS_PHYSICAL
        self.bag_wn_const = 4 / 3 * (self.V_s - self.V_b)            # analytic.py:80
        self.label_latex = label_latex or self.name
        self.label_unicode = label_unicode or self.name
        self.w_min = max(self.w(self.T_min, Phase.SYMMETRIC), self.w(self.T_min, Phase.BROKEN))   # model.py:95-100
        self.w_max = np.inf
        self.T_crit = self.critical_temp(allow_fail=allow_invalid)
        with np.errstate(all="ignore"):
            self.w_crit = float(self.w(self.T_crit, Phase.SYMMETRIC))    # model.py:460-474
        self.w_at_alpha_n_min, self.alpha_n_min = self.w_crit, self._alpha_n_min_checked(allow_fail=allow_invalid)

    # --- thermodynamic functions of temperature
    def _powfac(self, temp, mu):
        return (temp / self.T_ref) ** (mu - 4) * temp ** 4

    def p_temp(self, temp, phase):
        ps = self.a_s * self._powfac(temp, self.mu_s) - self.V_s
        pb = self.a_b * self._powfac(temp, self.mu_b) - self.V_b
        return pb * phase + ps * (1 - phase)

    def e_temp(self, temp, phase):
        es = (self.mu_s - 1) * self.a_s * self._powfac(temp, self.mu_s) + self.V_s
        eb = (self.mu_b - 1) * self.a_b * self._powfac(temp, self.mu_b) + self.V_b
        return eb * phase + es * (1 - phase)

    def s_temp(self, temp, phase):
        ss = self.mu_s * self.a_s * (temp / self.T_ref) ** (self.mu_s - 4) * temp ** 3
        sb = self.mu_b * self.a_b * (temp / self.T_ref) ** (self.mu_b - 4) * temp ** 3
        return sb * phase + ss * (1 - phase)

    def w(self, temp, phase):
        ws = self.mu_s * self.a_s * self._powfac(temp, self.mu_s)
        wb = self.mu_b * self.a_b * self._powfac(temp, self.mu_b)
        return wb * phase + ws * (1 - phase)

    def temp(self, w, phase):
        w = np.where(np.asarray(w, dtype=float) < 0, np.nan, w)
        ts = self.T_ref * (w / (self.mu_s * self.a_s * self.T_ref ** 4)) ** (1 / self.mu_s)
        tb = self.T_ref * (w / (self.mu_b * self.a_b * self.T_ref ** 4)) ** (1 / self.mu_b)
        out = tb * phase + ts * (1 - phase)
        return out if np.ndim(out) else float(out)

    # --- functions of enthalpy (model.py:641-766)
    def cs2(self, w, phase):
        return (phase * self.csb2 + (1 - phase) * self.css2) * np.ones_like(w)

    def cs2_max(self, w_max=None, phase=Phase.BROKEN, **kw):
        return (self.csb2 if phase == Phase.BROKEN else self.css2), np.nan

    cs2_min = cs2_max

    def p(self, w, phase):
        return self.p_temp(self.temp(w, phase), phase)

    def e(self, w, phase):
        return self.e_temp(self.temp(w, phase), phase)

    def s(self, w, phase):
        return self.s_temp(self.temp(w, phase), phase)

    def theta(self, w, phase):
        return 0.25 * (self.e(w, phase) - 3 * self.p(w, phase))

    def omega(self, w, phase):
        t = self.temp(w, phase)
        return self.p_temp(t, phase) / self.e_temp(t, phase)

    def nu_gdh2024(self, w, phase=Phase.BROKEN):
        om = self.omega(w, phase)
        return (1 - 3 * om) / (1 + 3 * om)

    def gp(self, w, phase):
        t = self.temp(w, phase)
        return 30 / np.pi ** 2 * self.e_temp(t, phase) / t ** 4       # model.py:683-686 (calls ge_temp)

    gs = gp

    def Psi_n(self, wn):
        tn = self.temp(wn, Phase.SYMMETRIC)
        return self.w(tn, Phase.BROKEN) / self.w(tn, Phase.SYMMETRIC)

    # --- transition strength
    def delta_theta(self, wp, wm):
        return self.theta(wp, Phase.SYMMETRIC) - self.theta(wm, Phase.BROKEN)

    def alpha_n(self, wn, **kw):
        tn = self.temp(wn, Phase.SYMMETRIC)                             # const_cs.py:161-199
        return (1 - 4 / self.mu_s) / 3 - (1 - 4 / self.mu_b) / 3 * self.w(tn, Phase.BROKEN) / wn + self.bag_wn_const / wn

    def alpha_plus(self, wp, wm, vp_tilde=None, sol_type=None, error_on_invalid=False, nan_on_invalid=True, **kw):
        ap = (1 - 4 / self.mu_s) / 3 - (1 - 4 / self.mu_b) * wm / (3 * wp) + self.bag_wn_const / wp   # const_cs.py:454
        if nan_on_invalid:                                               # model.py:368-405
            if sol_type in (SolutionType.SUB_DEF, SolutionType.HYBRID):
                bad = (ap < 0) | (ap >= 1 / 3)
            elif vp_tilde is not None:
                bad = (ap < 0) | (((1 + ap) * vp_tilde + (1 - 3 * ap) / (3 * vp_tilde)) ** 2 - 4 / 3 < 0)
            else:
                bad = ap < 0
            ap = np.where(bad, np.nan, ap)
            ap = ap if np.ndim(ap) else float(ap)
        return ap

    def alpha_theta_bar_n_from_alpha_n(self, alpha_n, wn=None, **kw):
        if wn is None:
            wn = self.wn(alpha_n)
        tn = self.temp(wn, Phase.SYMMETRIC)                              # model.py:347-354
        diff = (1 - 1 / (3 * self.csb2)) * (self.p_temp(tn, Phase.SYMMETRIC) - self.p_temp(tn, Phase.BROKEN)) / wn
        return alpha_n + diff

    def alpha_theta_bar_plus(self, wp, **kw):
        return (1 - self.mu_b / self.mu_s) / 3 + self.mu_b / 4 * self.bag_wn_const / wp   # const_cs.py:507-519

    def validate_alpha_n(self, alpha_n, allow_invalid=False, log_invalid=True):
        if alpha_n < 0 or alpha_n < self.alpha_n_min:                    # model.py:813-823
            msg = f"Invalid alpha_n={alpha_n}. Minimum for the model: {self.alpha_n_min}"
            if log_invalid:
                logger.error(msg)
            if not allow_invalid:
                raise ValueError(msg)

    def wn(self, alpha_n, wn_guess=None, **kw):
        """Enthalpy at the nucleation temperature for given alpha_n (vectorised)."""
        alpha_n = np.asarray(alpha_n, dtype=float)
        if np.isclose(self.mu_b, 4):                                     # bag.py:316-339, const_cs.py:715-728
            out = self.bag_wn_const / (alpha_n + (4 / self.mu_s - 1) / 3)
            out = np.where(out < 0, np.nan, out)
        else:
            if alpha_n.size > 64:
                # grids repeat the same alpha_n for every wall speed: solve each distinct value once (the iteration below
                # stops when ALL its values have converged, and the set of values is the same: identical results)
                uniq, inv = np.unique(alpha_n.ravel(), return_inverse=True)
                if uniq.size < alpha_n.size:
                    return np.asarray(self.wn(uniq))[inv.ravel()].reshape(alpha_n.shape)
            # alpha_n(wn) is monotonic below w_crit: safeguarded Newton on log(wn) (the reference: fsolve, model.py:856-922)
            lo = np.full(alpha_n.shape, np.log(self.w_min))
            hi = np.full(alpha_n.shape, np.log(self.w_crit))
            x = np.full(alpha_n.shape, np.log(0.9 * self.w_crit))
            f_lo = self.alpha_n(np.exp(lo)) - alpha_n
            sign = np.sign(f_lo)
            for _ in range(200):
                f = self.alpha_n(np.exp(x)) - alpha_n
                same = np.sign(f) == sign
                lo = np.where(same, x, lo)
                hi = np.where(same, hi, x)
                h = 1e-6
                df = (self.alpha_n(np.exp(x + h)) - alpha_n - f) / h
                with np.errstate(all="ignore"):
                    xn = x - f / df
                bad = ~np.isfinite(xn) | (xn <= lo) | (xn >= hi)
                xn = np.where(bad, 0.5 * (lo + hi), xn)
                if np.all(np.abs(xn - x) <= 1e-15 * np.abs(x) + 1e-300):
                    x = xn
                    break
                x = xn
            out = np.exp(x)
            out = np.where(np.abs(self.alpha_n(out) - alpha_n) > 1e-9 * np.maximum(1.0, np.abs(alpha_n)), np.nan, out)
        return out if out.ndim else float(out)

    def _critical_temp_opt(self, temp):
        """Zero at T_crit, where p_s = p_b (const_cs.py:521-523)."""
        const = (self.V_b - self.V_s) * self.T_ref ** 4
        return self.a_s * (temp / self.T_ref) ** self.mu_s - self.a_b * (temp / self.T_ref) ** self.mu_b + const

    def critical_temp(self, guess=None, allow_fail=False):
        """T_crit with p_s(T_crit) = p_b(T_crit).  Bag: closed form (bag.py:186-192).  Otherwise the reference's procedure
        (model.py:476-557): MINPACK hybrd (scipy.optimize.fsolve, the same library call) from the guess T_ref
        (const_cs.py:116-117), with the same validity checks and the same exceptions - the parameter search of
        ConstCSModel(alpha_n_min=...) depends on which trial models are rejected."""
        if np.isclose(self.mu_s, 4) and np.isclose(self.mu_b, 4):
            return ((self.V_s - self.V_b) / (self.a_s - self.a_b)) ** 0.25
        from scipy.optimize import fsolve
        if guess is None:
            guess = self.T_ref

        def fail(kind, msg):
            if not allow_fail:
                raise kind(msg)

        p_s_min, p_b_min = self.p_temp(self.T_min, Phase.SYMMETRIC), self.p_temp(self.T_min, Phase.BROKEN)
        if p_s_min >= p_b_min:
            fail(ValueError, "All models should have p_s(T=T_min) < p_b(T=T_min) for T_crit to exist. "
                             f"Got: T_min={self.T_min}, p_s={p_s_min}, p_b={p_b_min}.")
        t_arr = np.logspace(np.log10(self.T_min), np.log10(1e4), 10)
        if np.all(self.p_temp(t_arr, Phase.SYMMETRIC) <= self.p_temp(t_arr, Phase.BROKEN)):
            fail(ValueError, "All models should have p_s(T>T_crit) > p_b(T>T_crit) for T_crit to exist.")
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sol = fsolve(self._critical_temp_opt, x0=np.array([guess], dtype=float), full_output=True)
        t_crit = float(sol[0][0])
        if sol[2] != 1:
            fail(RuntimeError, f"Could not find Tc with guess={guess}. Using Tc={t_crit}. Reason: {sol[3]}")
        if t_crit <= self.T_min:
            fail(ValueError, f"T_crit should be higher than T_min. Got: T_crit={t_crit}, T_min={self.T_min}")
        with np.errstate(all="ignore"):
            p_crit_s, p_crit_b = self.p_temp(t_crit, Phase.SYMMETRIC), self.p_temp(t_crit, Phase.BROKEN)
        if not np.isclose(p_crit_s, p_crit_b):
            fail(ValueError, f"Pressures do not match at T_crit. Got: p_s={p_crit_s}, p_b={p_crit_b}")
        if p_crit_s < 0 or p_crit_b < 0:
            fail(ValueError, f"Pressure cannot be negative at T_crit. Got: p_s={p_crit_s}, p_b={p_crit_b}")
        return t_crit

    def _alpha_n_min_checked(self, allow_fail=False):
        """alpha_n at w_crit with the range checks of the reference (model.py:141-199, const_cs.py:161-199)."""
        w = self.w_crit
        if not allow_fail:
            if np.isnan(w):
                pass                                                     # logged only by the reference
            elif w < self.w_min:
                raise ValueError(f"Got wn={w} < w_min={self.w_min} for alpha_n.")
        with np.errstate(all="ignore"):
            an = float(self.alpha_n(w))
        if an < 0 and not allow_fail:
            raise ValueError(f"Got negative alpha_n={an} with wn={w}, mu_s={self.mu_s}, mu_b={self.mu_b}, "
                             f"t_crit={self.T_crit}.")
        return an

    # --- wall-speed classification
    def v_chapman_jouguet(self, alpha_n, wn=None):
        """chapman_jouguet.py:151-264 (analytical branches)."""
        if self.DEFAULT_NAME == "bag":
            return 1 / np.sqrt(3) * (1 + np.sqrt(2 * alpha_n + 3 * alpha_n ** 2)) / (1 + alpha_n)
        atbp = self.alpha_theta_bar_n_from_alpha_n(alpha_n, wn=wn)
        disc = 3 * atbp * (1 - self.csb2 + 3 * self.csb2 * atbp)
        return (1 + np.sqrt(disc)) / (1 / self.csb + 3 * self.csb * atbp)

    def solution_type_codes(self, v_wall, alpha_n, wn=None, an_max_bag=None):
        """Vectorised solution_type -> integer codes (0 sub-def, 1 hybrid, 2 detonation, -1 error)."""
        raise NotImplementedError

    def solution_type(self, v_wall, alpha_n, wn=None, **kw):
        return CODE_SOL[int(self.solution_type_codes(np.asarray([v_wall]), np.asarray([alpha_n]),
                                                     None if wn is None else np.asarray([wn]))[0])]

    def eos_params(self):
        """(css2, csb2, V_s, V_b, name_is_bag) for the device."""
        return self.css2, self.csb2, self.V_s, self.V_b, 1.0 if self.name == "bag" else 0.0

    def export(self):
        return {"name": self.name, "label_latex": self.label_latex, "label_unicode": self.label_unicode,
                "T_min": self.T_min, "T_max": self.T_max, "t_ref": self.T_ref, "V_s": self.V_s, "V_b": self.V_b,
                "a_s": self.a_s, "a_b": self.a_b, "css2": self.css2, "csb2": self.csb2, "mu_s": self.mu_s,
                "mu_b": self.mu_b}


class BagModel(Model):
    """pttools/models/bag.py:32-346."""
    DEFAULT_NAME = "bag"

    def __init__(self, V_s=1.0, V_b=0.0, a_s=None, a_b=None, alpha_n_min=None, name=None, **kw):
        if alpha_n_min is not None:
            # bag.py:120-137: a_s = a_b / (1 - 3 alpha_n_min * 0.999), capped by the default
            a_b_ = 1.0 if a_b is None else a_b
            a_s_def = 1.1 * a_b_ if a_s is None else a_s
            a_s = min(a_s_def, a_b_ / (1 - 3 * alpha_n_min * 0.999))
            a_b = a_b_
        if (1.1 * (1.0 if a_b is None else a_b) if a_s is None else a_s) <= (1.0 if a_b is None else a_b):
            raise ValueError("The bag model must have a_s > a_b for the critical temperature to be non-negative. "
                             f"Got: a_s={a_s}, a_b={a_b}")                       # bag.py:76-80
        if V_s <= V_b and not kw.get("allow_invalid", False):
            raise ValueError(f"The bubble will not expand in the Bag model, when V_s <= V_b. Got: V_s = V_b = {V_s}")
        super().__init__(1 / 3, 1 / 3, V_s=V_s, V_b=V_b, a_s=a_s, a_b=a_b, name=name, **kw)

    def theta(self, w, phase):
        return (self.V_b * phase + self.V_s * (1 - phase)) * np.ones_like(w)     # bag.py:311-316

    def alpha_n(self, wn, **kw):
        return self.bag_wn_const / wn                                            # analytic.py:97-119

    def alpha_plus(self, wp, wm=None, vp_tilde=None, sol_type=None, nan_on_invalid=True, **kw):
        ap = self.bag_wn_const / wp
        if nan_on_invalid:
            if sol_type in (SolutionType.SUB_DEF, SolutionType.HYBRID):
                bad = (ap < 0) | (ap >= 1 / 3)
            elif vp_tilde is not None:
                bad = (ap < 0) | (((1 + ap) * vp_tilde + (1 - 3 * ap) / (3 * vp_tilde)) ** 2 - 4 / 3 < 0)
            else:
                bad = ap < 0
            ap = np.where(bad, np.nan, ap)
            ap = ap if np.ndim(ap) else float(ap)
        return ap

    def solution_type_codes(self, v_wall, alpha_n, wn=None, an_max_bag=None):
        """identify_solution_type_bag (transition.py:19-44); an_max_bag = alpha_n_max_deflagration_bag(v_wall)."""
        from . import engine
        v_wall = np.asarray(v_wall, dtype=float)
        if an_max_bag is None:
            an_max_bag = engine.bag_alpha_n_max_unique(v_wall)[0].cpu().numpy()
        cs0 = 1 / np.sqrt(3)
        ap_max_det = np.where(v_wall < cs0, 0.0, (1 - np.sqrt(3) * v_wall) ** 2 / (3 * (1 - v_wall ** 2)))
        out = np.full(v_wall.shape, -1, dtype=np.int32)
        defl = alpha_n < an_max_bag
        out[defl & (v_wall <= cs0)] = 0
        out[defl & (v_wall > cs0)] = 1
        out[alpha_n < ap_max_det] = 2
        return out


class ConstCSModel(Model):
    """pttools/models/const_cs.py:43-737."""
    DEFAULT_NAME = "const_cs"

    def __init__(self, css2, csb2, V_s=1.0, V_b=0.0, a_s=None, a_b=None, T_ref=1.0, name=None, alpha_n_min=None, **kw):
        css2, csb2 = float(css2), float(csb2)
        for nm, c in (("css2", css2), ("csb2", csb2)):
            if not 0 < c <= 1:
                raise ValueError(f"{nm} has to be 0 < c_s^2 <= 1")
        if css2 > 1 / 3 and np.isclose(css2, 1 / 3):
            css2 = 1 / 3
        if csb2 > 1 / 3 and np.isclose(csb2, 1 / 3):
            csb2 = 1 / 3
        if alpha_n_min is not None:
            a_b_ = 1.0 if a_b is None else a_b                            # analytic.py:163-188 (get_a_g)
            a_s_ = 1.1 * a_b_ if a_s is None else a_s
            a_s, a_b, V_s, V_b = self.alpha_n_min_find_params(css2, csb2, alpha_n_min, a_s_default=a_s_, a_b=a_b_,
                                                              V_s_default=V_s, V_b=V_b)
        super().__init__(css2, csb2, V_s=V_s, V_b=V_b, a_s=a_s, a_b=a_b, T_ref=T_ref, name=name, **kw)

    @classmethod
    def alpha_n_min_find_params(cls, css2, csb2, alpha_n_min_target, a_s_default, a_b, V_s_default=1.0, V_b=0.0,
                                safety_factor_alpha=0.999, safety_factor_a=1.001, safety_factor_V=0.001, a_max=1e4):
        """(a_s, a_b, V_s, V_b) of a model whose alpha_n_min is just below the target: the search of const_cs.py:225-397
        with the same optimiser calls (scipy's bounded Brent minimiser on a_s for V_s and V_s/100, then L-BFGS-B on
        (a_s, V_s)), the same acceptance tests and the same fall-back to the defaults."""
        from scipy.optimize import minimize, minimize_scalar
        if safety_factor_a < 1 or safety_factor_V < 0:
            raise ValueError(f"Got invalid safety factors: a={safety_factor_a}, V={safety_factor_V}")
        defaults = (a_s_default, a_b, V_s_default, V_b)
        if np.isclose(1 + 1 / css2, 4) and np.isclose(1 + 1 / csb2, 4):        # bag.py:119-135
            if a_s_default < 0 or a_b < 0 or V_s_default < 0 or V_b < 0:
                raise ValueError(f"Invalid parameters: a_s_default={a_s_default}, a_b={a_b}, V_s_default={V_s_default}, "
                                 f"V_b={V_b}")
            a_s = a_b / (1 - 3 * alpha_n_min_target * safety_factor_alpha)
            return (a_s_default if a_s_default < a_s else a_s), a_b, V_s_default, V_b
        target = safety_factor_alpha * alpha_n_min_target

        def objective(a_s, V_s):
            try:
                trial = cls(css2, csb2, a_s=a_s, a_b=a_b, V_s=V_s, V_b=V_b)
            except (ValueError, RuntimeError):
                return np.nan
            return (trial.alpha_n_min - target) ** 2

        base = None
        try:
            base = cls(css2, csb2, a_s=a_s_default, a_b=a_b, V_s=V_s_default)      # V_b = 0 here (const_cs.py:331)
            if base.alpha_n_min < alpha_n_min_target:                            # already below the target
                return defaults
        except ValueError:
            pass
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for V_s in (V_s_default, V_s_default / 100):
                sol = minimize_scalar(objective, bracket=((a_s_default + a_b) / 2, a_s_default, 2 * a_s_default),
                                      bounds=(a_b * safety_factor_a, a_max), args=(V_s,))
                if sol.success:
                    found = cls(css2, csb2, a_s=sol.x, a_b=a_b, V_s=V_s, V_b=V_b)
                    if found.alpha_n_min <= alpha_n_min_target:
                        return sol.x, a_b, V_s, V_b
            sol = minimize(lambda x: objective(x[0], x[1]), x0=np.array([a_s_default, V_s_default], dtype=float),
                           bounds=((a_b * safety_factor_a, None), (V_b + safety_factor_V, None)))
        if not sol.success:
            logger.error("Failed to find alpha_n_min. Reason: %s. Using given defaults a_s=%s, a_b=%s, V_s=%s, V_b=%s "
                         "for css2=%s, csb2=%s.", sol.message, a_s_default, a_b, V_s_default, V_b, css2, csb2)
            return defaults
        a_s, V_s = sol.x[0], sol.x[1]
        found = cls(css2, csb2, a_s=a_s, a_b=a_b, V_s=V_s, V_b=V_b)
        if base is None:
            raise RuntimeError("alpha_n_min search: the default parameters do not give a valid model to compare with")
        if found.alpha_n_min > base.alpha_n_min:
            logger.error("alpha_n_min solver returned greater alpha_n_min than with default values. Using defaults "
                         "a_s=%s, a_b=%s, V_s=%s, V_b=%s for css2=%s, csb2=%s.", a_s_default, a_b, V_s_default, V_b, css2,
                         csb2)
            return defaults
        return a_s, a_b, V_s, V_b

    def solution_type_codes(self, v_wall, alpha_n, wn=None, an_max_bag=None):
        """const_cs.py:633-655."""
        v_wall = np.asarray(v_wall, dtype=float)
        alpha_n = np.asarray(alpha_n, dtype=float)
        if wn is None:
            wn = self.wn(alpha_n)
        atbp = self.alpha_theta_bar_n_from_alpha_n(alpha_n, wn=wn)
        b = 3 * atbp - 1 - v_wall ** 2 * (1 / self.csb2 + 3 * atbp)
        a = v_wall / self.csb2
        disc = b ** 2 - 4 * a * v_wall
        out = np.full(v_wall.shape, 2, dtype=np.int32)
        out[(b > 0) | (disc < 0)] = 1
        out[v_wall ** 2 < self.csb2] = 0
        return out
