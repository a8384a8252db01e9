"""pttools.bubble call surface on top of the CUDA kernels.

Batched operators (`sweep_bubbles`, `BubbleBatch`) are the product; the scalar functions and the `Bubble` class are
one-point batches with the reference's names, arguments and error behaviour
(pttools/bubble/{bubble,fluid,fluid_bag,alpha,transition,integrate}.py).
"""
from __future__ import annotations

import dataclasses
import logging
import typing as tp

import numpy as np

from . import _lib, engine
from . import fluid_reference
from .models import CODE_SOL, SOL_CODE, Model, Phase, SolutionType

logger = logging.getLogger(__name__)

N_XI_DEFAULT = 5000                # pttools/bubble/const.py:12
T_END_DEFAULT = 50.0               # const.py:18
THIN_SHELL_T_POINTS_MIN = 100      # const.py:26
CS0 = 1 / np.sqrt(3)
CS0_2 = 1 / 3


class NotYetSolvedError(RuntimeError):
    """pttools/bubble/bubble.py:30-31."""


# ---------------------------------------------------------------------------------------------------------------
# Old bag-model API (fluid_bag.py, alpha.py, transition.py)
# ---------------------------------------------------------------------------------------------------------------

def alpha_n_max_deflagration_bag(v_wall, n_xi: int = N_XI_DEFAULT):
    """alpha.py:71-88.  (As in the reference, the value does not depend on n_xi in practice: alpha.py:316.)"""
    scalar = np.isscalar(v_wall)
    vw = np.atleast_1d(np.asarray(v_wall, dtype=float))
    if np.any(vw >= 1.0) or np.any(vw <= 0.0):
        raise ValueError("Unphysical parameter(s): at least one value outside 0 < v_wall < 1.")
    out = engine.bag_alpha_n_max(vw)[0].cpu().numpy()
    return float(out[0]) if scalar else out


alpha_n_max_bag = alpha_n_max_deflagration_bag


def identify_solution_type_bag(v_wall: float, alpha_n: float, exit_on_error: bool = False) -> SolutionType:
    """transition.py:19-44."""
    from .models import BagModel
    code = int(BagModel.solution_type_codes(None, np.asarray([v_wall]), np.asarray([alpha_n]))[0])
    if code < 0 and exit_on_error:
        raise RuntimeError("No solution for given v_wall, alpha_n")
    return CODE_SOL[code]


def find_alpha_plus_bag(v_wall, alpha_n_given: float, n_xi: int = N_XI_DEFAULT, **_):
    """alpha.py:361-388: alpha_+ from alpha_n (NaN when alpha_n >= alpha_n_max_deflagration_bag)."""
    scalar = np.isscalar(v_wall)
    vw = np.atleast_1d(np.asarray(v_wall, dtype=float))
    ap, _, _ = engine.bag_find_alpha_plus(vw, np.full_like(vw, alpha_n_given), n_xi)
    ap = ap.cpu().numpy()
    return float(ap[0]) if scalar else ap


def sound_shell_bag(v_wall: float, alpha_n: float, n_xi: int = N_XI_DEFAULT, **_):
    """fluid_bag.py:33-79: (v, w, xi); length-1 NaN arrays when there is no solution."""
    if not 0.0 <= v_wall <= 1.0:
        raise ValueError("Unphysical parameter: v_wall. See the log for details.")
    sh = engine.sweep_shells_bag([v_wall], [alpha_n], n_xi=n_xi)
    return sh.profile(0)


def sound_shell_alpha_plus(v_wall: float, alpha_plus: float, sol_type=SolutionType.UNKNOWN, n_xi: int = N_XI_DEFAULT,
                           w_n: float = 1.0, **_):
    """fluid_bag.py:83-222."""
    if not 0.0 <= v_wall <= 1.0:
        raise ValueError("Unphysical parameter: v_wall. See the log for details.")
    if isinstance(sol_type, SolutionType):
        if sol_type == SolutionType.UNKNOWN:
            sol_type = identify_solution_type_alpha_plus(v_wall, alpha_plus)
        code = SOL_CODE[sol_type]
    else:
        code = int(sol_type)
    sh = engine.bag_profiles([v_wall], [alpha_plus], [code], n_xi=n_xi, w_n=[w_n])
    return sh.profile(0)


def find_alpha_n_bag(v_wall: float, alpha_p: float, sol_type=SolutionType.UNKNOWN, n_xi: int = N_XI_DEFAULT, **_) -> float:
    """alpha_n from alpha_+ for the bag model (alpha.py:229-255): alpha_+ w(wall) / w_n of the shell."""
    _, w, xi = sound_shell_alpha_plus(v_wall, alpha_p, sol_type, n_xi)
    n_wall = int(np.argmax(xi >= v_wall)) if np.any(xi >= v_wall) else 0        # props.find_v_index
    return float(alpha_p * w[n_wall] / w[-1])


def identify_solution_type_alpha_plus(v_wall: float, alpha_p: float) -> SolutionType:
    """transition.py:97-124."""
    if v_wall <= CS0:
        sol = SolutionType.SUB_DEF
    else:
        ap_max_det = (1 - np.sqrt(3) * v_wall) ** 2 / (3 * (1 - v_wall ** 2))
        sol = SolutionType.DETON if alpha_p < ap_max_det else SolutionType.HYBRID
    if alpha_p > 1 / 3 and sol != SolutionType.DETON:
        sol = SolutionType.ERROR
    return sol


fluid_shell_bag = sound_shell_bag
fluid_shell_alpha_plus = sound_shell_alpha_plus


def bag_quantities(v_wall, alpha_n, n_xi: int = N_XI_DEFAULT) -> tp.Dict[str, np.ndarray]:
    """The bubble-averaged quantities of the old bag API for a batch of points, one shell kernel + one integral kernel:
    ubarf2 = mean_kinetic_energy / w[-1] (quantities.py:442-456, 509-527), kappa = ubarf2 / (0.75 alpha_n) (:81-117),
    ke_frac = ubarf2 / (0.75 (1 + alpha_n)) (:250-263), mean_enthalpy_change (:423-439).  NaN where the reference
    has no solution (SolutionType.ERROR)."""
    torch = engine._torch()
    vw = np.atleast_1d(np.asarray(v_wall, dtype=np.float64))
    an = np.broadcast_to(np.asarray(alpha_n, dtype=np.float64), vw.shape).copy()
    shells = engine.sweep_shells_bag(vw, an, n_xi=n_xi)
    pp2 = torch.zeros((vw.size, 12), dtype=torch.float64, device="cuda")
    pp2[:, 0] = engine.dev_f64(vw)
    pp2[:, 1], pp2[:, 2] = 4.0, 4.0                 # bag: mu = 4 in both phases
    pp2[:, 5], pp2[:, 6], pp2[:, 7] = 1.0, 1.0, 1.0
    I = engine.profile_integrals(shells, pp2).cpu().numpy()
    ok = shells.status.cpu().numpy() == 0
    with np.errstate(invalid="ignore", divide="ignore"):
        ubarf2 = np.where(ok, I[:, 0] / vw ** 3 / I[:, 7], np.nan)
        out = {"ubarf2": ubarf2, "kappa": ubarf2 / (0.75 * an), "ke_frac": ubarf2 / (0.75 * (1 + an)),
               "mean_enthalpy_change": np.where(ok, I[:, 2] / vw ** 3, np.nan)}
    return out


def _scalar_or_array(v_wall, arr):
    return arr if isinstance(v_wall, np.ndarray) else type(v_wall)(arr[0])


def get_ubarf2_bag(v_wall, alpha_n: float, n_xi: int = N_XI_DEFAULT, verbosity: int = 0):
    """Mean square fluid velocity (quantities.py:328-346); v_wall may be a float or an array."""
    if not isinstance(v_wall, (float, np.floating, np.ndarray)):
        raise TypeError(f"Unknown type for v_wall: {type(v_wall)}")
    return _scalar_or_array(v_wall, bag_quantities(v_wall, alpha_n, n_xi)["ubarf2"])


def get_kappa_bag(v_wall, alpha_n: float, n_xi: int = N_XI_DEFAULT, verbosity: int = 0):
    """Efficiency factor kappa (quantities.py:81-117)."""
    return _scalar_or_array(v_wall, bag_quantities(v_wall, alpha_n, n_xi)["kappa"])


def get_ke_frac_bag(v_wall, alpha_n: float, n_xi: int = N_XI_DEFAULT):
    """Kinetic energy fraction of the total energy, bag equation of state (quantities.py:250-263)."""
    return _scalar_or_array(v_wall, bag_quantities(v_wall, alpha_n, n_xi)["ke_frac"])


def fluid_integrate_param(v0, w0, xi0, phase=-1., t_end=T_END_DEFAULT, n_xi=N_XI_DEFAULT, df_dtau_ptr=None,
                          method: str = "cuda", cs2_s: float = CS0_2, cs2_b: float = CS0_2):
    """integrate.py:81-135 as a one-lane batch.  `df_dtau_ptr` may be a Model (its sound speeds are used) or None
    (bag); `method` is accepted for signature compatibility and ignored: there is one integrator, the CUDA one."""
    if isinstance(df_dtau_ptr, Model):
        cs2_s, cs2_b = df_dtau_ptr.css2, df_dtau_ptr.csb2
    v, w, xi, st = engine.integrate_batch([v0], [w0], [xi0], [float(phase)], [t_end], n_xi, cs2_s, cs2_b)
    if int(st[0]) != 0:
        raise RuntimeError("Integration failed")
    return v[0].cpu().numpy(), w[0].cpu().numpy(), xi[0].cpu().numpy(), np.linspace(0., t_end, n_xi)


# ---------------------------------------------------------------------------------------------------------------
# Model-independent solver, batched
# ---------------------------------------------------------------------------------------------------------------

@dataclasses.dataclass
class BubbleBatch:
    """Solved bubbles of a sweep, all on the device.  `scal` holds the reference's per-bubble scalars
    (fluid.py:1052 and the cached properties of bubble.py:377-632) as [n] tensors (columns of the result table of
    ptt_sweep_spectra)."""
    model: tp.Union[Model, tp.Sequence[Model]]
    shells: engine.ShellBatch
    v_wall: tp.Any
    alpha_n: tp.Any
    wn: tp.Any
    sol_type: tp.Any          # int32 codes
    status: tp.Any
    scal: tp.Dict[str, tp.Any]
    pp: tp.Any                # [n,8] parameters of ptt_a2_batch
    n_xi: int
    invalid: tp.Any = None    # host bool [n]: points the reference's Bubble constructor rejects

    def __len__(self):
        return self.v_wall.numel()

    def validity(self, sum_rtol_error: float = 5e-2) -> tp.Dict[str, np.ndarray]:
        """The validity flags of Bubble.solve (bubble.py:123-131, 290-352) for every point of the sweep, as host bool
        arrays: solved, solver_failed, no_solution_found, unphysical_alpha_plus, negative_entropy_flux,
        negative_net_entropy_change, numerical_error.  One model for all points."""
        if not isinstance(self.model, Model):
            raise NotImplementedError("validity(): one model per sweep")
        m = self.model
        st = self.status.cpu().numpy()
        sc = {k: self.scal[k].cpu().numpy() for k in ("vp_tilde", "vm_tilde", "wp", "wm", "kappa", "omega",
                                                      "va_entropy_density_diff")}
        code = self.sol_type.cpu().numpy()
        wn = self.wn.cpu().numpy()
        solved = (st == 0) | (st == 5)
        with np.errstate(invalid="ignore", divide="ignore"):
            sp = np.where(code == 2, m.s(wn, Phase.SYMMETRIC), m.s(sc["wp"], Phase.SYMMETRIC))
            sm = m.s(sc["wm"], Phase.BROKEN)
            g = lambda v: 1 / np.sqrt(1 - v ** 2)
            fp, fm = g(sc["vp_tilde"]) * sc["vp_tilde"] * sp, g(sc["vm_tilde"]) * sc["vm_tilde"] * sm
            ap = np.array([m.alpha_plus(a, b, vp_tilde=c, sol_type=CODE_SOL.get(int(t)))
                           if ok else np.nan for a, b, c, t, ok in zip(sc["wp"], sc["wm"], sc["vp_tilde"], code, solved)])
            ksum = sc["kappa"] + sc["omega"]
        return {
            "solved": solved, "solver_failed": st == 5, "no_solution_found": ~solved,
            "unphysical_alpha_plus": solved & np.isnan(ap),
            "negative_entropy_flux": solved & ((fp < 0) | (fm < 0) | (fm - fp < 0)),
            "negative_net_entropy_change": solved & (sc["va_entropy_density_diff"] < 0),
            "numerical_error": solved & ~np.isclose(ksum, 1, rtol=sum_rtol_error),
        }


def _per_point(models, n, attr):
    if isinstance(models, Model):
        return np.full(n, getattr(models, attr), dtype=float)
    return np.array([getattr(m, attr) for m in models], dtype=float)


@dataclasses.dataclass
class PointTable:
    """Host-side parameter rows of a sweep: what the model contributes per point (closed forms of models.py and the
    starting guesses of the fluid-reference table).  pg / pm are the inputs of ptt_sweep_spectra."""
    pg: np.ndarray            # [n,16]
    pm: np.ndarray            # [n,4]
    wn: np.ndarray
    codes: np.ndarray         # solution type codes; -1 where the point is rejected
    invalid: np.ndarray       # bool: rejected by the checks of the reference's Bubble constructor


def point_table(model, v_wall, alpha_n, sol_type=None, allow_invalid=False, device=False) -> PointTable:
    """Per-point inputs of the shell solver for Bubble(model, v_wall, alpha_n) (bubble.py:36-177, fluid.py:890-949).
    Points the reference's constructor rejects -- alpha_n < model.alpha_n_min (model.py:813-823), T_n > T_crit
    (bubble.py:97-101), no solution type (transition.py:47-69), wn not found -- get solution type -1: the device
    reports them as PTT_ST_INVALID_INPUT with NaN outputs.  device=True: `pg` and `pm` are assembled as CUDA tensors (the
    starting guesses come from the device table), the rest stays numpy."""
    vw = np.asarray(v_wall, dtype=float).reshape(-1)
    an = np.asarray(alpha_n, dtype=float).reshape(-1)
    n = vw.size
    if not isinstance(model, Model):
        # one model per point: the closed forms are vectorised per distinct model, then put back in place
        models = list(model)
        if len(models) != n:
            raise ValueError("one model per point is needed")
        pg, pm = np.zeros((n, 16)), np.zeros((n, 4))
        wn, codes, invalid = np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=bool)
        ids = np.array([id(m) for m in models])
        st = None if sol_type is None else np.asarray(sol_type, dtype=np.int32).reshape(-1)
        for uid in np.unique(ids):
            sel = np.nonzero(ids == uid)[0]
            sub = point_table(models[sel[0]], vw[sel], an[sel], None if st is None else st[sel], allow_invalid)
            pg[sel], pm[sel], wn[sel], codes[sel], invalid[sel] = sub.pg, sub.pm, sub.wn, sub.codes, sub.invalid
        if device:
            pg, pm = engine.dev_f64(pg), engine.dev_f64(pm)
        return PointTable(pg, pm, wn, codes, invalid)
    with np.errstate(invalid="ignore", divide="ignore"):
        wn = np.asarray(model.wn(an), dtype=float).reshape(-1)
        codes = model.solution_type_codes(vw, an, wn) if sol_type is None else \
            np.asarray(sol_type, dtype=np.int32).reshape(-1)
        v_cj = np.asarray(model.v_chapman_jouguet(an, wn), dtype=float).reshape(-1)
        tn = np.asarray(model.temp(wn, Phase.SYMMETRIC), dtype=float).reshape(-1)
        codes = np.array(codes, dtype=np.int32)
        invalid = ~np.isfinite(wn) | (codes < 0) | (vw < 0) | (vw > 1)
        if not allow_invalid:
            invalid |= (an < 0) | (an < model.alpha_n_min) | (tn > model.T_crit)
    codes[invalid] = -1
    # entropy density s(w) = coef * w^(1 - 1/mu) per phase (s = w / T): the retry search applies the reference's
    # entropy-flux test to its scan starts
    mus, mub = 1 + 1 / model.css2, 1 + 1 / model.csb2
    s_coef_s = (mus * model.a_s) ** (1 / mus) * model.T_ref ** (4 / mus - 1)
    s_coef_b = (mub * model.a_b) ** (1 / mub) * model.T_ref ** (4 / mub - 1)
    consts = (model.css2, model.csb2, model.V_s, model.V_b, 1.0 if model.name == "bag" else 0.0, s_coef_s, s_coef_b)
    guesses = fluid_reference.ref().get_batch(vw, an, codes)                 # device [n,6]: vp, vm, vp~, vm~, wp, wm
    if device:
        # assemble the rows where the starting guesses already are: on the device
        torch = engine._torch()
        up = torch.as_tensor(np.stack([vw, an, codes.astype(float), wn, v_cj]), device="cuda")     # [5, n]
        wn_d = up[3]
        wm_g = guesses[:, 5] * wn_d
        cols = [up[0], up[1], up[2], wn_d, up[4], guesses[:, 0], guesses[:, 2], guesses[:, 4] * wn_d,    # fluid.py:939-941
                torch.where(torch.isnan(wm_g), 0.3 * wn_d, wm_g)]                                      # fluid.py:942-949
        cols += [torch.full_like(wn_d, c) for c in consts]
        pg = torch.stack(cols, dim=1).contiguous()
        pm = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
        pm[:, 0], pm[:, 1], pm[:, 2] = model.a_s, model.a_b, model.T_ref
        return PointTable(pg, pm, wn, codes, invalid)
    guesses = guesses.cpu().numpy()
    pgT = np.empty((16, n))                               # filled row-wise (contiguous), transposed once
    pgT[0], pgT[1], pgT[2], pgT[3], pgT[4] = vw, an, codes, wn, v_cj
    pgT[5], pgT[6] = guesses[:, 0], guesses[:, 2]
    pgT[7] = guesses[:, 4] * wn                           # fluid.py:939-941
    wm_g = guesses[:, 5] * wn
    pgT[8] = np.where(np.isnan(wm_g), 0.3 * wn, wm_g)     # fluid.py:942-949
    for k, c in enumerate(consts):
        pgT[9 + k] = c
    pg = np.ascontiguousarray(pgT.T)
    pm = np.zeros((n, 4))
    pm[:, 0], pm[:, 1], pm[:, 2] = model.a_s, model.a_b, model.T_ref
    return PointTable(pg, pm, wn, codes, invalid)


def sweep_bubbles(model, v_wall, alpha_n, sol_type=None, n_xi=N_XI_DEFAULT, t_end=T_END_DEFAULT,
                  thin_shell_limit=THIN_SHELL_T_POINTS_MIN, wn_rtol=1e-4, thermo=True, retry="sync",
                  allow_invalid=False, use_bag_solver=False) -> BubbleBatch:
    """Bubble(model, v_wall, alpha_n).solve() for every point of a sweep (bubble.py:232-352, fluid.py:848-1052) through
    ptt_sweep_spectra (shells, retries, thermodynamic scalars; no spectra).  `model` is one Model or a sequence with
    one Model per point.  retry: "sync" = the reference's retries (fsolve_vary, backup scan) for the points whose first
    root search fails; "off" = report them as solver failures.  Profiles stay on the device (`shells`).
    use_bag_solver: for a BagModel the reference's OLD solver (sound_shell_bag) under the Bubble interface
    (fluid.py:892-918); ignored for other models, as in the reference."""
    torch = engine._torch()
    if retry not in ("sync", "off"):
        raise ValueError(f"retry={retry!r}")
    if use_bag_solver and isinstance(model, Model) and model.name == "bag":
        return _sweep_bubbles_bag_solver(model, v_wall, alpha_n, allow_invalid)
    pt = point_table(model, v_wall, alpha_n, sol_type, allow_invalid)
    n = pt.wn.size
    out = engine.sweep_c(None, engine.dev_f64(pt.pg), engine.dev_f64(pt.pm), n_xi=n_xi, t_end=t_end,
                         thin_shell_limit=thin_shell_limit, wn_rtol=wn_rtol, rows=True, retry=retry == "sync")
    pg = engine.dev_f64(pt.pg)
    sol = torch.as_tensor(pt.codes, dtype=torch.int32, device="cuda")
    shells = engine.ShellBatch(pg[:, 0].contiguous(), pg[:, 1].contiguous(), None, sol, out.status, out.rows[0],
                               out.rows[1], out.rows[2], out.off, out.npts, out.scalars, n_xi)
    scal = {name: out.scalars[:, k] for k, name in enumerate(_lib.SWEEP_SCALARS)}
    mu_s, mu_b = 1 + 1 / pg[:, 9], 1 + 1 / pg[:, 10]
    pp = engine.point_params(pg[:, 0], torch.sqrt(pg[:, 10]), 1 - 1 / mu_s, 1 - 1 / mu_b, pg[:, 11], pg[:, 12])
    return BubbleBatch(model, shells, shells.v_wall, shells.alpha_n, engine.dev_f64(pt.wn), sol, out.status, scal, pp,
                       n_xi, invalid=pt.invalid)


def _sweep_bubbles_bag_solver(model: Model, v_wall, alpha_n, allow_invalid=False) -> BubbleBatch:
    """Bubble(BagModel, v_wall, alpha_n, use_bag_solver=True).solve() for every point (fluid.py:892-918, bubble.py:250-
    352): sound_shell_bag(v_wall, alpha_n) at its default n_xi with w scaled by wn, the wall values of
    props.v_and_w_from_solution (:51-87, including its wm_sh = w[argmax(flip(w) > wn)]), v_sh / vm_sh / vm_tilde_sh = NaN
    (and with them the quantities that need v_sh), the thermodynamic integrals of the scaled profile.  Points within 0.05
    of alpha_n_max_deflagration_bag ("high alpha_n", bubble.py:253) come back as the reference returns them: solved with
    solver_failed and a one-sample NaN profile."""
    torch = engine._torch()
    pt = point_table(model, v_wall, alpha_n, None, allow_invalid)
    vw, an = engine.dev_f64(pt.pg[:, 0].copy()), engine.dev_f64(pt.pg[:, 1].copy())
    n = pt.wn.size
    wn = engine.dev_f64(pt.wn)
    an_max, _ = engine.bag_alpha_n_max_unique(vw)
    shells = engine.sweep_shells_bag(vw, an, n_xi=N_XI_DEFAULT, an_max=an_max)
    shells.rows_w.mul_(wn[:, None])                                   # the old solver works at wn = 1
    high = (an_max - an) < 0.05
    shells.sol_type[high] = -1                                         # fluid.py:893-897: SolutionType.ERROR
    invalid = torch.as_tensor(pt.invalid, device="cuda")
    failed = high | (shells.status != 0)
    wall = engine.shell_wall_values(shells)                           # vp, vm, vp_tilde, vm_tilde, wp, wm, w[-1]
    # wm_sh as the reference computes it: index of the first w > w[-1] in the FLIPPED array, used in the unflipped one
    cap = shells.rows_w.shape[1]
    off, npts = shells.off.long(), shells.npts.long()
    last = (off + npts - 1).clamp(0, cap - 1)
    rows = torch.arange(n, device="cuda")
    w_last = shells.rows_w[rows, last]
    wm_sh = torch.empty(n, dtype=torch.float64, device="cuda")
    col = torch.arange(cap, device="cuda")[None, :]
    for a in range(0, n, 2048):
        b = min(n, a + 2048)
        w = shells.rows_w[a:b]
        inside = (col >= off[a:b, None]) & (col <= last[a:b, None]) & (w > w_last[a:b, None])
        p_max = torch.where(inside, col, -1).max(dim=1).values
        j = torch.where(p_max >= 0, last[a:b] - p_max, torch.zeros_like(p_max))
        wm_sh[a:b] = w[torch.arange(b - a, device="cuda"), (off[a:b] + j).clamp(0, cap - 1)]
    mu_s, mu_b = 1 + 1 / model.css2, 1 + 1 / model.csb2
    pp2 = torch.zeros((n, 12), dtype=torch.float64, device="cuda")
    pp2[:, 0] = vw
    for k, c in enumerate((mu_s, mu_b, model.V_s, model.V_b, model.a_s, model.a_b, model.T_ref, 1.0)):
        pp2[:, 1 + k] = c
    I = engine.profile_integrals(shells, pp2)
    # thermo.py:42-221 from the integrals (the algebra of the sweep driver's thermo_finish_kernel); v_sh is NaN here
    nan = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
    four_pi_3, vw3 = 4 * np.pi / 3, vw ** 3
    va_kin, va_trace, va_thd, va_s, wbar = four_pi_3 * I[:, 0], four_pi_3 * I[:, 1], four_pi_3 * 0.75 * I[:, 2], \
        four_pi_3 * I[:, 3], I[:, 4]
    ebar = (1 - 1 / mu_s) * wn + model.V_s
    ked = 3 / (4 * np.pi * vw3) * va_kin
    kef = ked / ebar
    v_cj = engine.dev_f64(np.asarray(model.v_chapman_jouguet(pt.pg[:, 1], pt.wn), dtype=float).reshape(-1))
    cols = [wall[:, 0], wall[:, 1], wall[:, 2], wall[:, 3], nan, nan, nan, wall[:, 4], wall[:, 5], wm_sh, v_cj, wn,
            wall[:, 6], failed.double(), va_kin, va_trace, va_thd, va_s, wbar, ebar, ked, kef, nan, nan, nan,
            va_kin / va_trace.abs(), va_thd / va_trace.abs(), ked / I[:, 7], wbar / ebar, nan, nan, shells.npts.double()]
    table = torch.stack(cols, dim=1)
    assert table.shape[1] == len(_lib.SWEEP_SCALARS)
    table[failed | invalid] = float("nan")
    table[:, 13] = failed.double()
    status = torch.where(invalid, torch.full_like(shells.status, 4), torch.where(failed, torch.full_like(shells.status, 5),
                                                                                  torch.zeros_like(shells.status)))
    # failed points carry a one-sample NaN profile (const.nan_arr)
    bad = failed | invalid
    shells.off[bad] = 0
    shells.npts[bad] = 1
    for r in (shells.rows_v, shells.rows_w, shells.rows_xi):
        r[bad, 0] = float("nan")
    shells.status = status
    scal = {name: table[:, k] for k, name in enumerate(_lib.SWEEP_SCALARS)}
    cs_b = torch.full((n,), float(np.sqrt(model.csb2)), dtype=torch.float64, device="cuda")
    pp = engine.point_params(vw, cs_b, 1 - 1 / mu_s, 1 - 1 / mu_b, model.V_s, model.V_b)
    shells.scalars = table
    return BubbleBatch(model, shells, vw, an, wn, shells.sol_type, status, scal, pp, N_XI_DEFAULT, invalid=pt.invalid)


# ---------------------------------------------------------------------------------------------------------------
# Bubble object (one point)
# ---------------------------------------------------------------------------------------------------------------

class Bubble:
    """A solution of the hydrodynamic equations (bubble.py:34-632); a view of a one-point device batch."""

    _THERMO = ("kappa", "omega", "kinetic_energy_density", "kinetic_energy_fraction", "thermal_energy_density",
               "va_enthalpy_density", "va_kinetic_energy_density", "va_thermal_energy_density",
               "va_thermal_energy_density_diff", "va_trace_anomaly_diff", "va_entropy_density_diff", "ubarf2", "wbar",
               "ebar", "mean_adiabatic_index")

    def __init__(self, model: Model, v_wall: float, alpha_n: float, solve: bool = True, sol_type: SolutionType = None,
                 label_latex: str = None, label_unicode: str = None, wn_guess: float = None, wm_guess: float = None,
                 theta_bar: bool = False, t_end: float = T_END_DEFAULT, n_xi: int = N_XI_DEFAULT,
                 thin_shell_t_points_min: int = THIN_SHELL_T_POINTS_MIN, use_bag_solver: bool = False,
                 use_giese_solver: bool = False, log_success: bool = False, allow_invalid: bool = False,
                 log_invalid: bool = True):
        if use_giese_solver or theta_bar:
            raise NotImplementedError("use_giese_solver / theta_bar are outside the hot path")
        self.use_bag_solver = bool(use_bag_solver)
        if v_wall < 0 or v_wall > 1:
            raise ValueError(f"Invalid v_wall={v_wall}")
        model.validate_alpha_n(alpha_n, allow_invalid=allow_invalid, log_invalid=log_invalid)
        self.model, self.v_wall, self.alpha_n = model, v_wall, alpha_n
        self.wn = float(model.wn(alpha_n))
        self.alpha_theta_bar_n = float(model.alpha_theta_bar_n_from_alpha_n(alpha_n, wn=self.wn))
        if sol_type is None or sol_type is SolutionType.UNKNOWN:
            sol_type = model.solution_type(v_wall, alpha_n, wn=self.wn)
        if sol_type in (SolutionType.UNKNOWN, SolutionType.ERROR):
            raise ValueError(
                f"Could not determine solution type automatically for model={model.name}, v_wall={v_wall}, "
                f"alpha_n={alpha_n}. Got sol_type={sol_type}. Please choose it manually.")      # transition.py:47-69
        self.sol_type = sol_type
        self.t_end, self.n_xi, self.thin_shell_t_points_min = t_end, n_xi, thin_shell_t_points_min
        self.Tn = float(model.temp(self.wn, Phase.SYMMETRIC))
        if self.Tn > model.T_crit and not allow_invalid:
            raise ValueError(f"Bubbles form only when T_nuc < T_crit. Got: T_nuc={self.Tn}, T_crit={model.T_crit}")
        self.Psi_n = float(model.Psi_n(self.wn))
        self.solved = False
        self.solver_failed = self.no_solution_found = False
        self.negative_entropy_flux = self.negative_net_entropy_change = False
        self.numerical_error = self.unphysical_alpha_plus = False
        self.label_latex = label_latex or rf"{model.label_latex}, $v_w={v_wall}, \alpha_n={alpha_n}$"
        self.label_unicode = label_unicode or f"{model.label_unicode}, v_w={v_wall}, alpha_n={alpha_n}"
        self.notes: tp.List[str] = []
        self.v = self.w = self.xi = self.phase = None
        self._batch: tp.Optional[BubbleBatch] = None
        if solve:
            self.solve()

    def solve(self, sum_rtol_warning: float = 1.5e-2, sum_rtol_error: float = 5e-2, use_bag_solver: bool = False, **_):
        batch = sweep_bubbles(self.model, [self.v_wall], [self.alpha_n], [SOL_CODE[self.sol_type]], self.n_xi,
                              self.t_end, self.thin_shell_t_points_min,
                              use_bag_solver=self.use_bag_solver or use_bag_solver)
        self._batch = batch
        st = int(batch.status[0])
        if int(batch.sol_type[0]) < 0:
            self.sol_type = SolutionType.ERROR         # the old bag solver's answer to "high alpha_n" (fluid.py:893-897)
        if st not in (0, 5):
            self.no_solution_found = True              # bubble.py:290-295 (IndexError / RuntimeError in the solver)
            return
        self.solver_failed = st == 5
        self.v, self.w, self.xi = batch.shells.profile(0)
        sc = {k: float(v[0]) for k, v in batch.scal.items()}
        for k in ("vp", "vm", "vp_tilde", "vm_tilde", "v_sh", "vm_sh", "vm_tilde_sh", "wp", "wm", "wm_sh", "v_cj"):
            setattr(self, k, sc[k])
        self._sc = sc
        m = self.model
        self.sn = float(m.s(self.wn, Phase.SYMMETRIC))
        self.sm = float(m.s(self.wm, Phase.BROKEN))
        if self.sol_type == SolutionType.DETON:
            self.sp, self.sm_sh = self.sn, self.sm
        else:
            self.sp = float(m.s(self.wp, Phase.SYMMETRIC))
            self.sm_sh = float(m.s(self.wm_sh, Phase.SYMMETRIC))
        self.w_center = self.w[0]
        self.solved = True
        self.phase = find_phase(self.xi, self.v_wall)
        self.alpha_plus = m.alpha_plus(self.wp, self.wm, vp_tilde=self.vp_tilde, sol_type=self.sol_type)
        self.alpha_theta_bar_plus = m.alpha_theta_bar_plus(self.wp)
        if np.isnan(self.alpha_plus):
            self.alpha_plus = m.alpha_plus(self.wp, self.wm, nan_on_invalid=False)
            self.unphysical_alpha_plus = True
        g = lambda v: 1 / np.sqrt(1 - v ** 2)
        self.entropy_flux_p = g(self.vp_tilde) * self.vp_tilde * self.sp
        self.entropy_flux_m = g(self.vm_tilde) * self.vm_tilde * self.sm
        self.entropy_flux_diff = self.entropy_flux_m - self.entropy_flux_p
        if self.entropy_flux_p < 0 or self.entropy_flux_m < 0 or self.entropy_flux_diff < 0:
            self.negative_entropy_flux = True
        if sc["va_entropy_density_diff"] < 0:
            self.negative_net_entropy_change = True
        if not np.isclose(sc["kappa"] + sc["omega"], 1, rtol=sum_rtol_warning):
            if not np.isclose(sc["kappa"] + sc["omega"], 1, rtol=sum_rtol_error):
                self.numerical_error = True

    def __getattr__(self, name):
        if name in Bubble._THERMO:
            if not self.__dict__.get("solved", False):
                raise NotYetSolvedError
            return self.__dict__["_sc"][name]
        raise AttributeError(name)

    @property
    def vp_tilde_sh(self):
        return self.v_sh

    @property
    def kappa_giese(self) -> float:
        """bubble.py:539-543."""
        if not self.solved:
            raise NotYetSolvedError
        return 4 * self.kinetic_energy_density / (3 * self.alpha_theta_bar_n * self.wn)

    def export(self, path: str = None):
        data = {"model": self.model.export(), "v_wall": self.v_wall, "alpha_n": self.alpha_n, "sol_type": self.sol_type,
                "t_end": self.t_end, "n_xi": self.n_xi, "thin_shell_limit": self.thin_shell_t_points_min, "v": self.v,
                "w": self.w, "xi": self.xi, "notes": self.notes}
        for k in ("alpha_plus", "sp", "sm", "sm_sh", "sn", "Tn", "v_cj", "vp", "vm", "vp_tilde", "vm_tilde", "v_sh",
                  "vm_sh", "vm_tilde_sh", "wn", "wp", "wm", "wm_sh"):
            data[k] = getattr(self, k, None)
        if path is not None:
            import json
            with open(path, "w") as f:
                json.dump(data, f, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o))
        return data


def find_phase(xi: np.ndarray, v_wall: float) -> np.ndarray:
    """props.py:22-32."""
    i_wall = int(np.argmax(xi >= v_wall))
    phase = np.zeros_like(xi)
    if i_wall == 0:
        return phase
    phase[:i_wall - 1] = Phase.BROKEN.value
    if np.isclose(xi[i_wall], v_wall):
        phase[i_wall - 1] = Phase.BROKEN.value
    return phase


def sound_shell_generic(model: Model, v_wall: float, alpha_n: float, sol_type: SolutionType = None, wn: float = None,
                        t_end: float = T_END_DEFAULT, n_xi: int = N_XI_DEFAULT,
                        thin_shell_limit: int = THIN_SHELL_T_POINTS_MIN, wn_rtol: float = 1e-4, **_):
    """fluid.py:848-1052: the 17-tuple (v, w, xi, sol_type, vp, vm, vp_tilde, vm_tilde, v_sh, vm_sh, vm_tilde_sh, wp,
    wm, wm_sh, v_cj, solver_failed, elapsed)."""
    import time
    t0 = time.perf_counter()
    codes = None if sol_type is None else [SOL_CODE[sol_type]]
    b = sweep_bubbles(model, [v_wall], [alpha_n], codes, n_xi, t_end, thin_shell_limit, wn_rtol, thermo=False)
    st = int(b.status[0])
    if st not in (0, 5):
        raise RuntimeError(f"The fluid shell solver failed with status {st}")
    v, w, xi = b.shells.profile(0)
    sc = {k: float(x[0]) for k, x in b.scal.items()}
    return (v, w, xi, CODE_SOL[int(b.sol_type[0])], sc["vp"], sc["vm"], sc["vp_tilde"], sc["vm_tilde"], sc["v_sh"],
            sc["vm_sh"], sc["vm_tilde_sh"], sc["wp"], sc["wm"], sc["wm_sh"], sc["v_cj"], st == 5,
            time.perf_counter() - t0)


fluid_shell_generic = sound_shell_generic
