"""Oracle (test infrastructure only): the two in-scope equations of state, restated from the reference.

BagModel      pttools/models/bag.py:32-346      p = a T^4 - V,  e = 3 a T^4 + V,  w = 4 a T^4,  cs2 = 1/3
ConstCSModel  pttools/models/const_cs.py:43-737 p = a (T/T0)^(mu-4) T^4 - V,  e = (mu-1) a (...) + V,  w = mu a (...)
with mu = 1 + 1/cs2 per phase.  A bag model is the constant-cs model with mu_s = mu_b = 4.
"""
from __future__ import annotations

import dataclasses

import numpy as np
from scipy.optimize import fsolve, root_scalar

SYMMETRIC, BROKEN = 0.0, 1.0


@dataclasses.dataclass
class Eos:
    kind: str            # "bag" | "const_cs"
    css2: float = 1 / 3
    csb2: float = 1 / 3
    a_s: float = 1.1
    a_b: float = 1.0
    V_s: float = 1.0
    V_b: float = 0.0
    T_ref: float = 1.0
    T_min: float = 1e-3   # models/base.py:22

    def __post_init__(self):
        self.mu_s = 1 + 1 / self.css2        # const_cs.py:19-24
        self.mu_b = 1 + 1 / self.csb2
        self.bag_wn_const = 4 / 3 * (self.V_s - self.V_b)     # analytic.py:80
        self.name = self.kind
        self.T_crit = self.critical_temp()
        self.w_crit = self.w(self.T_crit, SYMMETRIC)           # model.py:460-474
        self.alpha_n_min = self.alpha_n(self.w_crit)           # bag.py:117-118, const_cs.py:218-223

    # --- phase-wise helpers
    def _mu(self, phase):
        return phase * self.mu_b + (1 - phase) * self.mu_s

    def _a(self, phase):
        return phase * self.a_b + (1 - phase) * self.a_s

    def _V(self, phase):
        return phase * self.V_b + (1 - phase) * self.V_s

    def cs2(self, w, phase):
        """bag.py:21-27, const_cs.py:553-557."""
        return (phase * self.csb2 + (1 - phase) * self.css2) * np.ones_like(w)

    def temp(self, w, phase):
        """bag.py:296-309, const_cs.py:657-672."""
        w = np.where(np.asarray(w) < 0, np.nan, w)
        ts = self.T_ref * (w / (self.mu_s * self.a_s * self.T_ref ** 4)) ** (1 / self.mu_s)
        tb = self.T_ref * (w / (self.mu_b * self.a_b * self.T_ref ** 4)) ** (1 / self.mu_b)
        return tb * phase + ts * (1 - phase)

    def w(self, temp, phase):
        ws = self.mu_s * self.a_s * (temp / self.T_ref) ** (self.mu_s - 4) * temp ** 4
        wb = self.mu_b * self.a_b * (temp / self.T_ref) ** (self.mu_b - 4) * temp ** 4
        return wb * phase + ws * (1 - phase)

    def p_temp(self, temp, phase):
        ps = self.a_s * (temp / self.T_ref) ** (self.mu_s - 4) * temp ** 4 - self.V_s
        pb = self.a_b * (temp / self.T_ref) ** (self.mu_b - 4) * temp ** 4 - self.V_b
        return pb * phase + ps * (1 - phase)

    def e_temp(self, temp, phase):
        es = (self.mu_s - 1) * self.a_s * (temp / self.T_ref) ** (self.mu_s - 4) * temp ** 4 + self.V_s
        eb = (self.mu_b - 1) * self.a_b * (temp / self.T_ref) ** (self.mu_b - 4) * temp ** 4 + self.V_b
        return eb * phase + es * (1 - phase)

    def s_temp(self, temp, phase):
        ss = self.mu_s * self.a_s * (temp / self.T_ref) ** (self.mu_s - 4) * temp ** 3
        sb = self.mu_b * self.a_b * (temp / self.T_ref) ** (self.mu_b - 4) * temp ** 3
        return sb * phase + ss * (1 - phase)

    def p(self, w, phase):
        return self.p_temp(self.temp(w, phase), phase)      # model.py:700-706

    def e(self, w, phase):
        return self.e_temp(self.temp(w, phase), phase)      # model.py:641-647

    def s(self, w, phase):
        return self.s_temp(self.temp(w, phase), phase)      # model.py:727-733

    def theta(self, w, phase):
        if self.kind == "bag":
            return (self.V_b * phase + self.V_s * (1 - phase)) * np.ones_like(w)    # bag.py:311-316
        return 0.25 * (self.e(w, phase) - 3 * self.p(w, phase))                     # model.py:758-766

    def omega(self, w, phase):
        t = self.temp(w, phase)
        return self.p_temp(t, phase) / self.e_temp(t, phase)      # model.py:695-698

    def nu_gdh2024(self, w, phase=BROKEN):
        om = self.omega(w, phase)
        return (1 - 3 * om) / (1 + 3 * om)                        # model.py:688-693

    def critical_temp(self):
        if self.kind == "bag":
            return ((self.V_s - self.V_b) / (self.a_s - self.a_b)) ** 0.25      # bag.py:186-192

        def opt(t):    # const_cs.py:521-523
            return self.a_s * (t / self.T_ref) ** self.mu_s - self.a_b * (t / self.T_ref) ** self.mu_b \
                + (self.V_b - self.V_s) * self.T_ref ** 4
        return float(fsolve(opt, x0=np.array([self.T_ref]))[0])                 # model.py:517-522

    def alpha_n(self, wn):
        if self.kind == "bag":
            return self.bag_wn_const / wn                              # analytic.py:97-119
        tn = self.temp(wn, SYMMETRIC)                                  # const_cs.py:161-199
        return (1 - 4 / self.mu_s) / 3 - (1 - 4 / self.mu_b) / 3 * self.w(tn, BROKEN) / wn + self.bag_wn_const / wn

    def wn(self, alpha_n):
        if self.kind == "bag":
            return self.bag_wn_const / alpha_n                         # bag.py:316-339
        if np.isclose(self.mu_b, 4):                                   # const_cs.py:715-728
            return self.bag_wn_const / (alpha_n + (4 / self.mu_s - 1) / 3)
        guess = 0.9 * self.w_crit                                      # model.py:931-974
        sol = fsolve(lambda x: self.alpha_n(x[0]) - alpha_n, x0=np.array([guess]), full_output=True)
        if sol[2] != 1:
            sol = fsolve(lambda x: self.alpha_n(x[0]) - alpha_n, x0=np.array([1.0]), full_output=True)
        if sol[2] != 1:
            r = root_scalar(lambda x: self.alpha_n(x) - alpha_n, x0=guess, x1=0.99 * self.w_crit,
                            bracket=(self.w(self.T_min, SYMMETRIC), self.w_crit))
            return float(r.root)
        return float(sol[0][0])

    def alpha_theta_bar_n_from_alpha_n(self, alpha_n, wn):
        """model.py:347-354."""
        tn = self.temp(wn, SYMMETRIC)
        diff = (1 - 1 / (3 * self.csb2)) * (self.p_temp(tn, SYMMETRIC) - self.p_temp(tn, BROKEN)) / wn
        return alpha_n + diff

    def alpha_plus(self, wp, wm):
        if self.kind == "bag":
            return self.bag_wn_const / wp                                            # analytic.py:121-151
        return (1 - 4 / self.mu_s) / 3 - (1 - 4 / self.mu_b) * wm / (3 * wp) + self.bag_wn_const / wp   # const_cs.py:454

    def gp_dummy(self):
        return None


def bag_model(a_s=1.1, a_b=1.0, V_s=1.0, V_b=0.0):
    return Eos("bag", 1 / 3, 1 / 3, a_s, a_b, V_s, V_b)


def const_cs_model(css2, csb2, a_s=1.1, a_b=1.0, V_s=1.0, V_b=0.0, T_ref=1.0):
    return Eos("const_cs", css2, csb2, a_s, a_b, V_s, V_b, T_ref)
