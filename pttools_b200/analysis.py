"""Sweep drivers with the reference's signatures (pttools/analysis/parallel.py:21-196, bubble_grid.py:13-80).

The reference maps one future per parameter point over a process pool and returns object arrays of pickled
Bubble / Spectrum instances; here the whole grid goes through the device pipeline once and the objects are thin,
lazily materialised views of rows of the result table."""
from __future__ import annotations

import typing as tp

import numpy as np

from . import bubble as bub
from . import sweep
from ._lib import SWEEP_SCALARS as _SWEEP_SCALARS
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
        self._flags = None

    NEEDS_SOLUTION = frozenset(_SWEEP_SCALARS) - {"wn", "v_cj"}

    def _validity(self):
        if self._flags is None:
            BubbleView._cache_validity(self._b)
            self._flags = {k: bool(v[self._i]) for k, v in self._b._validity.items()}
        return self._flags

    @staticmethod
    def _cache_validity(batch):
        if getattr(batch, "_validity", None) is None:
            batch._validity = batch.validity()

    numerical_error = property(lambda self: self._validity()["numerical_error"])
    unphysical_alpha_plus = property(lambda self: self._validity()["unphysical_alpha_plus"])
    negative_entropy_flux = property(lambda self: self._validity()["negative_entropy_flux"])
    negative_net_entropy_change = property(lambda self: self._validity()["negative_net_entropy_change"])

    def _profile(self):
        if self._prof is None:
            self._prof = self._b.shells.profile(self._i)
        return self._prof

    v = property(lambda self: self._profile()[0])
    w = property(lambda self: self._profile()[1])
    xi = property(lambda self: self._profile()[2])

    @property
    def alpha_theta_bar_n(self) -> float:
        return float(self.model.alpha_theta_bar_n_from_alpha_n(self.alpha_n, wn=self.wn))

    @property
    def kappa_giese(self) -> float:
        """bubble.py:539-543: 4 e_K / (3 alpha_theta_bar_n w_n)."""
        if not self.solved:
            raise bub.NotYetSolvedError
        return 4 * self.kinetic_energy_density / (3 * self.alpha_theta_bar_n * self.wn)

    def __getattr__(self, name):
        scal = self.__dict__["_b"].scal
        if name in scal:
            if not self.__dict__.get("solved", False) and name in BubbleView.NEEDS_SOLUTION:
                raise bub.NotYetSolvedError                       # bubble.py:377-632: cached properties of an unsolved bubble
            return float(scal[name][self.__dict__["_i"]])
        raise AttributeError(name)


class SpectrumView:
    """Row of a spectrum sweep with the attribute names of pttools.omgw0.Spectrum."""

    def __init__(self, res: sweep.SweepResult, i: int, bubble: tp.Optional[BubbleView], r_star, Tn=None, g_star=None,
                 gs_star=None):
        from . import omgw0 as om
        self._r, self._i, self.bubble, self.r_star = res, i, bubble, r_star
        self.y = res.y
        # both in-scope models have non-physical temperatures: the overrides of omgw0_ssm.py:62-112 apply
        self.Tn = om.T_default if Tn is None else Tn
        self.g_star = om.G_STAR_DEFAULT if g_star is None else gs_star           # sic: omgw0_ssm.py:74
        self.gs_star = self.g_star if gs_star is None else gs_star

    pow_gw = property(lambda self: self._r.pow_gw[self._i].cpu().numpy())
    pow_v = property(lambda self: self._r.pow_v[self._i].cpu().numpy())
    source_lifetime_factor = property(lambda self: float(self._r.scalars["source_lifetime_factor"][self._i]))

    def omgw0(self):
        return self._r.omgw0[self._i].cpu().numpy()

    @property
    def f_star0(self):
        from . import omgw0 as om
        return om.f_star0(Tn=self.Tn, g_star=self.g_star)

    def f(self, z=None):
        """omgw0_ssm.py:162-171."""
        from . import omgw0 as om
        return om.f(self.y if z is None else z, self.r_star, self.f_star0)

    def omgw0_peak(self):
        om_ = self.omgw0()
        i = int(np.argmax(om_))
        return self.f()[i], om_[i]

    def noise(self):
        from . import noise
        return noise.omega_noise(self.f())                    # omgw0_ssm.py:117-118

    def noise_ins(self):
        from . import noise
        return noise.omega_ins(self.f())                      # omgw0_ssm.py:120-121

    def signal_to_noise_ratio(self) -> float:
        from . import noise
        return float(noise.signal_to_noise_ratio(f=self.f(), signal=self.omgw0(), noise=self.noise()))      # :143-144

    def signal_to_noise_ratio_instrument(self) -> float:
        from . import noise
        return float(noise.signal_to_noise_ratio(f=self.f(), signal=self.omgw0(), noise=self.noise_ins()))  # :146-147


def _rejected(model: Model, v_wall: float, alpha_n: float) -> Exception:
    """The exception the reference's Bubble constructor raises for a point the sweep marked invalid (bubble.py:36-177)."""
    if alpha_n < 0 or alpha_n < model.alpha_n_min:
        return ValueError(f"Invalid alpha_n={alpha_n}. Minimum for the model: {model.alpha_n_min}")
    return ValueError(f"Bubble(model={model.label_unicode}, v_wall={v_wall}, alpha_n={alpha_n}) is rejected by the "
                      "constructor: no solution type, or T_nuc > T_crit")


def _bubble_objects(model, batch: bub.BubbleBatch, shape, allow_bubble_failure: bool):
    """Object array of BubbleView.  Like the reference's create_bubble (parallel.py:21-54): a point the Bubble constructor
    rejects raises, or becomes None under allow_bubble_failure; a point whose solve fails is still returned, with
    solved = False and no_solution_found = True."""
    objs = np.empty(int(np.prod(shape)), dtype=object)
    for i in range(objs.size):
        if batch.invalid is not None and batch.invalid[i]:
            if not allow_bubble_failure:
                raise _rejected(model, float(batch.v_wall[i]), float(batch.alpha_n[i]))
            objs[i] = None
        else:
            objs[i] = BubbleView(batch, i, model)
    return objs


def create_bubbles(model: Model, v_walls: np.ndarray, alpha_ns: np.ndarray, func: callable = None,
                   log_progress_percentage: float = 5, max_workers: int = None, single_thread: bool = False,
                   allow_bubble_failure: bool = False, kwargs: tp.Dict[str, tp.Any] = None,
                   bubble_kwargs: tp.Dict[str, tp.Any] = None, bubble_func: callable = None):
    """parallel.py:89-154: object array [n_alpha, n_vw] of bubbles (+ the outputs of `func`, applied on the host).
    `kwargs` may carry allow_bubble_failure / use_bag_solver the way BubbleGridVWAlpha passes them (bubble_grid.py:66-69);
    the remaining entries go to `func`."""
    v_walls, alpha_ns = np.asarray(v_walls, dtype=float), np.asarray(alpha_ns, dtype=float)
    kwargs = dict(kwargs or {})
    allow_bubble_failure = bool(kwargs.pop("allow_bubble_failure", allow_bubble_failure))
    use_bag_solver = bool(kwargs.pop("use_bag_solver", False))
    vw, an = np.meshgrid(v_walls, alpha_ns)                     # params[i_alpha, i_vw], parallel.py:126-130
    bk = bubble_kwargs or {}
    batch = bub.sweep_bubbles(model, vw.ravel(), an.ravel(), n_xi=bk.get("n_xi", bub.N_XI_DEFAULT),
                              t_end=bk.get("t_end", bub.T_END_DEFAULT), allow_invalid=bk.get("allow_invalid", False),
                              use_bag_solver=use_bag_solver or bool(bk.get("use_bag_solver", False)))
    flat = _bubble_objects(model, batch, vw.shape, allow_bubble_failure)
    objs = flat.reshape(vw.shape)
    if func is None:
        return objs
    return (objs,) + _apply(func, flat, vw.shape, kwargs)


def _apply(func, flat, shape, kwargs):
    """Post-functions carry .return_type / .fail_value attributes (bubble_quantities.py:6-39); a tuple return_type means
    several output arrays (parallel.py:103-110)."""
    if not hasattr(func, "return_type"):
        raise ValueError("The function should have a return_type attribute for output array initialization")
    kwargs = kwargs or {}
    multiple = isinstance(func.return_type, tuple)
    fail = getattr(func, "fail_value", (np.nan,) * len(func.return_type) if multiple else np.nan)
    vals = [fail if o is None else func(o, **kwargs) for o in flat]
    if not multiple:
        return (np.array(vals, dtype=func.return_type).reshape(shape),)
    return tuple(np.array([v[k] for v in vals], dtype=t).reshape(shape) for k, t in enumerate(func.return_type))


def create_spectra(model: Model, v_walls: np.ndarray, alpha_ns: np.ndarray, func: callable = None,
                   log_progress_percentage: float = 5, max_workers: int = None, single_thread: bool = False,
                   allow_bubble_failure: bool = False, kwargs: tp.Dict[str, tp.Any] = None,
                   bubble_kwargs: tp.Dict[str, tp.Any] = None, spectrum_kwargs: tp.Dict[str, tp.Any] = None):
    """parallel.py:157-187: object array [n_alpha, n_vw] of spectra (+ the outputs of `func`)."""
    v_walls, alpha_ns = np.asarray(v_walls, dtype=float), np.asarray(alpha_ns, dtype=float)
    vw, an = np.meshgrid(v_walls, alpha_ns)                     # params[i_alpha, i_vw]: the order of the returned arrays
    sk = dict(spectrum_kwargs or {})
    bk = bubble_kwargs or {}
    nuc = sk.get("nuc_type", "exponential")
    res = sweep.sweep_spectra(model, vw.ravel(), an.ravel(), y=sk.get("y"), r_star=sk.get("r_star"),
                              lifetime_multiplier=sk.get("lifetime_multiplier", 1.0),
                              nuc_type=str(getattr(nuc, "value", nuc)),
                              nt=sk.get("nt", 10000), n_z_lookup=sk.get("n_z_lookup", 10000),
                              z_st_thresh=sk.get("z_st_thresh", 50.0), n_xi=bk.get("n_xi", bub.N_XI_DEFAULT),
                              Tn=sk.get("Tn"), g_star=sk.get("g_star"), gs_star=sk.get("gs_star"), keep_bubbles=True,
                              allow_invalid=bk.get("allow_invalid", False))
    bubbles = _bubble_objects(model, res.bubbles, vw.shape, allow_bubble_failure)
    flat = np.empty(bubbles.size, dtype=object)
    for i, b in enumerate(bubbles):
        flat[i] = None if b is None else SpectrumView(res, i, b, sk.get("r_star"), sk.get("Tn"), sk.get("g_star"), sk.get("gs_star"))
    objs = flat.reshape(vw.shape)
    if func is None:
        return objs
    return (objs,) + _apply(func, flat, vw.shape, kwargs)


class BubbleGridVWAlpha:
    """bubble_grid.py:57-80."""

    def __init__(self, model: Model, v_walls: np.ndarray, alpha_ns: np.ndarray, func: callable = None,
                 use_bag_solver: bool = False, **kw):
        self.model, self.v_walls, self.alpha_ns = model, v_walls, alpha_ns
        kw = dict(kw)
        kw["kwargs"] = {**(kw.get("kwargs") or {}), "use_bag_solver": use_bag_solver}        # bubble_grid.py:66-69
        out = create_bubbles(model, v_walls, alpha_ns, func, allow_bubble_failure=True, **kw)
        if isinstance(out, tuple):
            self.bubbles = out[0]
            self.data = out[1] if len(out) == 2 else out[1:]               # bubble_grid.py:70-74
        else:
            self.bubbles, self.data = out, None

    def get_value(self, name: str, dtype=None) -> np.ndarray:
        """bubble_grid.py:18-34: None (NaN for float dtypes) where there is no bubble or it has not been solved."""
        def one(b):
            if b is None or (not b.solved and (name in BubbleView.NEEDS_SOLUTION or name == "kappa_giese")):
                return None
            return getattr(b, name)
        return np.array([[one(b) for b in row] for row in self.bubbles], dtype=dtype)

    def kappa(self):
        return self.get_value("kappa", dtype=np.float64)

    def omega(self):
        return self.get_value("omega", dtype=np.float64)

    def kappa_giese(self):
        return self.get_value("kappa_giese", dtype=np.float64)

    def numerical_error(self):
        return self.get_value("numerical_error", dtype=np.bool_)

    def solver_failed(self):
        return self.get_value("solver_failed", dtype=np.bool_)

    def unphysical_alpha_plus(self):
        return self.get_value("unphysical_alpha_plus", dtype=np.bool_)

    def negative_net_entropy_change(self):
        return self.get_value("negative_net_entropy_change", dtype=np.bool_)
