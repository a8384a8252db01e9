"""Sweep drivers with the reference's signatures (pttools/analysis/parallel.py:21-196, bubble_grid.py:13-80).

The reference maps one future per parameter point over a process pool and returns object arrays of pickled
Bubble / Spectrum instances; here the whole grid goes through the device pipeline once and the objects are thin,
lazily materialised views of rows of the result table."""
from __future__ import annotations

import typing as tp

import numpy as np

from . import bubble as bub
from . import sweep
from .models import CODE_SOL, Model, SolutionType


class BubbleView:
    """Row of a solved sweep with the attribute names of pttools.bubble.Bubble (bubble.py:143-168)."""

    def __init__(self, batch: bub.BubbleBatch, i: int, model: Model):
        self._b, self._i, self.model = batch, i, model
        self.v_wall, self.alpha_n = float(batch.v_wall[i]), float(batch.alpha_n[i])
        self.wn = float(batch.wn[i])
        self.sol_type = CODE_SOL[int(batch.sol_type[i])]
        st = int(batch.status[i])
        self.solved = st in (0, 5)
        self.solver_failed = st == 5
        self.no_solution_found = not self.solved
        self._prof = None

    def _profile(self):
        if self._prof is None:
            self._prof = self._b.shells.profile(self._i)
        return self._prof

    v = property(lambda self: self._profile()[0])
    w = property(lambda self: self._profile()[1])
    xi = property(lambda self: self._profile()[2])

    def __getattr__(self, name):
        scal = self.__dict__["_b"].scal
        if name in scal:
            return float(scal[name][self.__dict__["_i"]])
        raise AttributeError(name)


class SpectrumView:
    """Row of a spectrum sweep with the attribute names of pttools.omgw0.Spectrum."""

    def __init__(self, res: sweep.SweepResult, i: int, bubble: tp.Optional[BubbleView], r_star):
        self._r, self._i, self.bubble, self.r_star = res, i, bubble, r_star
        self.y = res.y

    pow_gw = property(lambda self: self._r.pow_gw[self._i].cpu().numpy())
    pow_v = property(lambda self: self._r.pow_v[self._i].cpu().numpy())

    def omgw0(self):
        return self._r.omgw0[self._i].cpu().numpy()


def create_bubbles(model: Model, v_walls: np.ndarray, alpha_ns: np.ndarray, func: callable = None,
                   log_progress_percentage: float = 5, max_workers: int = None, allow_bubble_failure: bool = False,
                   kwargs: tp.Dict[str, tp.Any] = None, bubble_kwargs: tp.Dict[str, tp.Any] = None,
                   bubble_func: callable = None):
    """parallel.py:89-154: object array [n_alpha, n_vw] of bubbles (+ the outputs of `func`, applied on the host)."""
    v_walls, alpha_ns = np.asarray(v_walls, dtype=float), np.asarray(alpha_ns, dtype=float)
    vw, an = np.meshgrid(v_walls, alpha_ns)                     # params[i_alpha, i_vw], parallel.py:126-130
    bk = bubble_kwargs or {}
    batch = bub.sweep_bubbles(model, vw.ravel(), an.ravel(), n_xi=bk.get("n_xi", bub.N_XI_DEFAULT),
                              t_end=bk.get("t_end", bub.T_END_DEFAULT))
    objs = np.empty(vw.shape, dtype=object)
    flat = objs.reshape(-1)
    for i in range(flat.size):
        view = BubbleView(batch, i, model)
        if not view.solved:
            if not allow_bubble_failure:
                raise RuntimeError(f"Solver crashed with model={model.label_unicode}, v_wall={view.v_wall}, "
                                   f"alpha_n={view.alpha_n}.")
            view = None
        flat[i] = view
    if func is None:
        return objs
    return objs, _apply(func, flat, vw.shape, kwargs)


def _apply(func, flat, shape, kwargs):
    """Post-functions carry .return_type / .fail_value attributes (bubble_quantities.py:6-39)."""
    kwargs = kwargs or {}
    fail = getattr(func, "fail_value", np.nan)
    vals = [fail if o is None else func(o, **kwargs) for o in flat]
    return np.array(vals, dtype=getattr(func, "return_type", float)).reshape(shape)


def create_spectra(model: Model, v_walls: np.ndarray, alpha_ns: np.ndarray, func: callable = None,
                   log_progress_percentage: float = 5, max_workers: int = None, allow_bubble_failure: bool = False,
                   kwargs: tp.Dict[str, tp.Any] = None, bubble_kwargs: tp.Dict[str, tp.Any] = None,
                   spectrum_kwargs: tp.Dict[str, tp.Any] = None):
    """parallel.py:157-187."""
    v_walls, alpha_ns = np.asarray(v_walls, dtype=float), np.asarray(alpha_ns, dtype=float)
    # device order: v_wall-major and ascending (points that share v_wall are contiguous); objects go back into the
    # reference's [i_alpha, i_vw] layout (parallel.py:126-130)
    order_vw = np.argsort(v_walls, kind="stable")
    vw = np.repeat(v_walls[order_vw], alpha_ns.size)
    an = np.tile(alpha_ns, v_walls.size)
    sk = dict(spectrum_kwargs or {})
    bk = bubble_kwargs or {}
    nuc = sk.get("nuc_type", "exponential")
    res = sweep.sweep_spectra(model, vw, an, y=sk.get("y"), r_star=sk.get("r_star"),
                              lifetime_multiplier=sk.get("lifetime_multiplier", 1.0),
                              nuc_type=str(getattr(nuc, "value", nuc)),
                              nt=sk.get("nt", 10000), n_z_lookup=sk.get("n_z_lookup", 10000),
                              z_st_thresh=sk.get("z_st_thresh", 50.0), n_xi=bk.get("n_xi", bub.N_XI_DEFAULT),
                              Tn=sk.get("Tn"), g_star=sk.get("g_star"), gs_star=sk.get("gs_star"), keep_bubbles=True)
    objs = np.empty((alpha_ns.size, v_walls.size), dtype=object)
    k = 0
    for bb in res.bubbles:
        for j in range(len(bb)):
            bview = BubbleView(bb, j, model)
            if not bview.solved and not allow_bubble_failure:
                raise RuntimeError(f"Solver crashed with v_wall={bview.v_wall}, alpha_n={bview.alpha_n}.")
            i_vw, i_alpha = order_vw[k // alpha_ns.size], k % alpha_ns.size
            objs[i_alpha, i_vw] = SpectrumView(res, k, bview, sk.get("r_star")) if bview.solved else None
            k += 1
    flat = objs.reshape(-1)
    vw = objs          # only .shape is used below
    if func is None:
        return objs
    return objs, _apply(func, flat, vw.shape, kwargs)


class BubbleGridVWAlpha:
    """bubble_grid.py:57-80."""

    def __init__(self, model: Model, v_walls: np.ndarray, alpha_ns: np.ndarray, func: callable = None,
                 use_bag_solver: bool = False, **kw):
        if use_bag_solver:
            raise NotImplementedError("use_bag_solver is outside the hot path")
        self.model, self.v_walls, self.alpha_ns = model, v_walls, alpha_ns
        out = create_bubbles(model, v_walls, alpha_ns, func, allow_bubble_failure=True, **kw)
        self.bubbles, self.data = (out if isinstance(out, tuple) else (out, None))

    def get_value(self, name: str) -> np.ndarray:
        return np.array([[np.nan if b is None else getattr(b, name) for b in row] for row in self.bubbles])

    def kappa(self):
        return self.get_value("kappa")

    def omega(self):
        return self.get_value("omega")

    def numerical_error(self):
        return np.array([[b is None for b in row] for row in self.bubbles])
