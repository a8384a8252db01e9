"""Generates tests/golden/*.npz by IMPORTING THE REFERENCE in the build container (it cannot travel to the GPU box).

Usage:   eval $(tools/refenv/setup_ref.sh) && python tools/make_golden.py [--only NAME]

Two kinds of fixtures are written:
  * ref_goldens.npz  -- the numeric golden vectors the reference's own tests pin (tests/test_data/*.txt there),
                        re-saved verbatim with their provenance, so that the oracle can be checked against them where
                        /root/reference does not exist;
  * ref_*.npz        -- outputs of the reference itself, run here on the inputs named in each entry
                        (profiles, alpha_+, A^2, spectra, omgw0).
Nothing from the reference's sources is copied.
"""
import argparse
import os
import sys

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF_TEST_DATA = "/root/reference/tests/test_data"


def goldens():
    d = {}
    for name in ("shell", "shells_weak", "shells_inter", "shells_esp", "bubble", "data_compare_nuc-test_reference",
                 "xi-v_plane"):
        d[name.replace("-", "_")] = np.loadtxt(os.path.join(REF_TEST_DATA, name + ".txt"))
    for name in ("v_plus_minus", "v_minus_plus"):
        d[name] = np.loadtxt(os.path.join(REF_TEST_DATA, name + ".txt"))
    np.savez_compressed(os.path.join(OUT, "ref_goldens.npz"), **d)
    print("ref_goldens.npz", {k: v.shape for k, v in d.items()})


BAG_POINTS = [(0.5, 0.1), (0.7, 0.1), (0.9, 0.1), (0.3, 0.2), (0.44, 0.0046), (0.7, 0.052), (0.58, 0.3),
              (0.92, 0.05), (0.731, 0.05), (0.05, 0.3), (0.2, 0.33), (0.6, 0.004), (0.56, 0.05), (0.8, 0.0046)]


def bag_shells():
    """sound_shell_bag profiles (decimated) + alpha_+ + scalars from the reference."""
    from pttools.bubble import fluid_bag, alpha, transition, quantities
    d = {"points": np.array(BAG_POINTS)}
    for i, (vw, an) in enumerate(BAG_POINTS):
        v, w, xi = fluid_bag.sound_shell_bag(vw, an)
        st = transition.identify_solution_type_bag(vw, an)
        ap = alpha.find_alpha_plus_bag(vw, an)
        d[f"v_{i}"], d[f"w_{i}"], d[f"xi_{i}"] = v, w, xi
        d[f"scal_{i}"] = np.array([ap, alpha.alpha_n_max_deflagration_bag(vw), quantities.ubarf_squared(v, w, xi, vw),
                                   {"Detonation": 2, "Hybrid": 1, "Subsonic deflagration": 0}[st.value]])
    np.savez_compressed(os.path.join(OUT, "ref_bag_shells.npz"), **d)
    print("ref_bag_shells.npz", len(BAG_POINTS))


SPEC_POINTS = [(0.5, 0.1), (0.7, 0.1), (0.9, 0.1), (0.44, 0.0046), (0.56, 0.05)]


def bag_spectra():
    """power_gw_scaled_bag / power_v_bag / a2_e_conserving_bag from the reference at reduced sizes."""
    import pttools.ssmtools as ssm
    z = np.logspace(-1, 3, 200)
    npt = (2000, 500, 1000)
    d = {"points": np.array(SPEC_POINTS), "z": z, "npt": np.array(npt)}
    zq = np.logspace(np.log10(0.2), np.log10(1000), 500)
    d["z_a2"] = zq
    for i, (vw, an) in enumerate(SPEC_POINTS):
        for nuc in ("exponential", "simultaneous"):
            d[f"pow_gw_{nuc[:3]}_{i}"] = ssm.power_gw_scaled_bag(z, (vw, an, ssm.NucType(nuc), (1.,)), npt=npt)
            d[f"pow_v_{nuc[:3]}_{i}"] = ssm.power_v_bag(z, (vw, an, ssm.NucType(nuc), (1.,)), npt=npt)
        a2, fp2, lam2 = ssm.a2_e_conserving_bag(zq, vw, an, npt=npt)
        d[f"a2_{i}"], d[f"fp2_{i}"], d[f"lam2_{i}"] = a2, fp2, lam2
    np.savez_compressed(os.path.join(OUT, "ref_bag_spectra.npz"), **d)
    print("ref_bag_spectra.npz", len(SPEC_POINTS))


def sin_transforms():
    """sin_transform on synthetic profiles (config 4 shape, tiny)."""
    import pttools.ssmtools as ssm
    rng = np.random.default_rng(0)
    z = np.logspace(-1, 3, 300)
    xi = np.linspace(0, 1, 513)
    d = {"z": z, "xi": xi}
    fs = []
    for p in range(4):
        phi, a, b = rng.uniform(0, 1), rng.uniform(0.1, 0.4), rng.uniform(0.5, 0.9)
        f = np.maximum(0, np.cos(xi + phi)) * ((xi > a) & (xi < b))
        fs.append(f)
        d[f"st_{p}"] = ssm.sin_transform(z, xi, f)
    d["f"] = np.array(fs)
    np.savez_compressed(os.path.join(OUT, "ref_sin_transform.npz"), **d)
    print("ref_sin_transform.npz")


TASKS = {"goldens": goldens, "bag_shells": bag_shells, "bag_spectra": bag_spectra, "sin_transforms": sin_transforms}

# ---------------------------------------------------------------------------------------------------------------
# Generic (model-independent) solver and SSMSpectrum / omgw0 fixtures
# ---------------------------------------------------------------------------------------------------------------

GENERIC_MODELS = {
    # name: (class, kwargs)
    "bag": ("BagModel", dict(a_s=1.1, a_b=1, V_s=1)),
    "cs_quarter_third": ("ConstCSModel", dict(css2=1 / 4, csb2=1 / 3, a_s=1.5, a_b=1, V_s=1)),
    "cs_third_quarter": ("ConstCSModel", dict(css2=1 / 3, csb2=1 / 4, a_s=5, a_b=1, V_s=1)),
    "cs_032_third": ("ConstCSModel", dict(css2=1 / 3 - 0.01, csb2=1 / 3, a_s=1.5, a_b=1, V_s=1)),
}
# per-model points with alpha_n above the model's alpha_n_min (0, 0.173, 0.2755, 0.1194)
GENERIC_POINTS = {
    "bag": [(0.5, 0.1), (0.7, 0.1), (0.9, 0.1), (0.3, 0.05), (0.3, 0.2), (0.62, 0.15), (0.8, 0.3), (0.55, 0.2),
            (0.44, 0.02), (0.95, 0.25)],
    "cs_quarter_third": [(0.5, 0.2), (0.7, 0.2), (0.9, 0.2), (0.3, 0.25), (0.62, 0.3), (0.8, 0.3), (0.55, 0.2),
                         (0.95, 0.25), (0.4, 0.35), (0.65, 0.18)],
    "cs_third_quarter": [(0.3, 0.3), (0.45, 0.3), (0.55, 0.35), (0.7, 0.3), (0.8, 0.3), (0.9, 0.4), (0.62, 0.4),
                         (0.95, 0.3)],
    "cs_032_third": [(0.5, 0.15), (0.7, 0.15), (0.9, 0.15), (0.3, 0.2), (0.62, 0.15), (0.8, 0.3), (0.55, 0.2),
                     (0.95, 0.25)],
}
SPECTRUM_POINTS = {"bag": [(0.5, 0.1), (0.7, 0.1), (0.9, 0.1)], "cs_quarter_third": [(0.5, 0.2), (0.8, 0.3)],
                   "cs_third_quarter": [(0.45, 0.3), (0.7, 0.3)]}
SCALARS = ["wn", "vp", "vm", "vp_tilde", "vm_tilde", "v_sh", "vm_sh", "vm_tilde_sh", "wp", "wm", "wm_sh", "v_cj",
           "kappa", "omega", "kinetic_energy_fraction", "va_enthalpy_density", "ubarf2", "wbar", "ebar",
           "va_kinetic_energy_density", "va_trace_anomaly_diff", "va_thermal_energy_density_diff",
           "va_entropy_density_diff", "alpha_plus", "Tn"]


def fluid_reference_table():
    from pttools.bubble import fluid_reference
    ref = fluid_reference.ref()
    import numpy as np
    raw = np.load(ref.path, allow_pickle=False)
    d = {k: raw[k] for k in raw.files}
    np.savez_compressed(os.path.join(OUT, "ref_fluid_reference.npz"), **d)
    print("ref_fluid_reference.npz", list(d))


def generic_bubbles():
    import pttools.models as models
    from pttools.bubble import Bubble
    from pttools.bubble.boundary import SolutionType
    codes = {SolutionType.SUB_DEF: 0, SolutionType.HYBRID: 1, SolutionType.DETON: 2}
    d = {"scalar_names": np.array(SCALARS)}
    for mname, (cls, kw) in GENERIC_MODELS.items():
        model = getattr(models, cls)(**kw)
        d[f"{mname}__model"] = np.array([getattr(model, "css2", 1 / 3), getattr(model, "csb2", 1 / 3), model.a_s, model.a_b,
                                         model.V_s, model.V_b, model.T_ref, model.T_crit, model.w_crit, model.alpha_n_min])
        d[f"{mname}__points"] = np.array(GENERIC_POINTS[mname])
        for i, (vw, an) in enumerate(GENERIC_POINTS[mname]):
            try:
                b = Bubble(model, vw, an)
            except Exception as e:  # invalid alpha_n for the model etc.
                d[f"{mname}__err_{i}"] = np.array([1.0])
                print(mname, vw, an, "ctor error", type(e).__name__, str(e)[:80])
                continue
            if not b.solved:
                d[f"{mname}__err_{i}"] = np.array([2.0])
                continue
            d[f"{mname}__v_{i}"], d[f"{mname}__w_{i}"], d[f"{mname}__xi_{i}"] = b.v, b.w, b.xi
            d[f"{mname}__phase_{i}"] = b.phase
            d[f"{mname}__flags_{i}"] = np.array([codes[b.sol_type], b.solver_failed, b.no_solution_found,
                                                 b.numerical_error, b.unphysical_alpha_plus, b.negative_entropy_flux,
                                                 b.negative_net_entropy_change], dtype=float)
            d[f"{mname}__scal_{i}"] = np.array([float(getattr(b, s)) for s in SCALARS])
            print(mname, vw, an, b.sol_type.value, b.v.size, "failed" if b.solver_failed else "ok")
    np.savez_compressed(os.path.join(OUT, "ref_generic_bubbles.npz"), **d)


def generic_spectra():
    import pttools.models as models
    from pttools.bubble import Bubble
    from pttools.omgw0 import Spectrum
    d = {}
    for mname, pts in SPECTRUM_POINTS.items():
        cls, kw = GENERIC_MODELS[mname]
        model = getattr(models, cls)(**kw)
        d[f"{mname}__points"] = np.array(pts)
        for i, (vw, an) in enumerate(pts):
            b = Bubble(model, vw, an)
            s = Spectrum(b, r_star=0.1)
            d[f"{mname}__pow_v_{i}"], d[f"{mname}__pow_gw_{i}"] = s.pow_v, s.pow_gw
            d[f"{mname}__omgw0_{i}"], d[f"{mname}__f_{i}"] = s.omgw0(), s.f()
            d[f"{mname}__a2_{i}"] = s.a2
            d[f"{mname}__misc_{i}"] = np.array([s.cs, s.source_lifetime_factor, s.F_gw0(), s.f_star0,
                                                s.suppression_factor(), b.kinetic_energy_fraction])
            print(mname, vw, an, "omgw0 peak", s.omgw0_peak())
        d["y"] = s.y
    np.savez_compressed(os.path.join(OUT, "ref_generic_spectra.npz"), **d)


def suppression_ssm():
    """The reference's stored results of calc_sup_ssm (tests/omgw0/test_suppression.py:32-59): integrals of
    power_gw_scaled_bag over ln z and get_ubarf2_bag for the simulation points, npt = (2000, 200, 320)."""
    import pttools.omgw0.suppression.suppression_ssm_data as pkg
    src = os.path.dirname(os.path.abspath(pkg.__file__))
    d = {}
    for f in ("suppression_2_ssm", "suppression_no_hybrids_ssm"):
        with np.load(os.path.join(src, f + ".npz")) as z:
            for k in z:
                d[f"{f}__{k}"] = z[k]
    np.savez_compressed(os.path.join(OUT, "ref_suppression_ssm.npz"), **d)


def model_tables():
    """The reference's EOS known answers (tests/test_data/models/{bag,const_cs}.json, tests/models/base_model.py:41-100)."""
    import json
    root = os.path.join(REF_TEST_DATA, "models")
    out = {n: json.load(open(os.path.join(root, n + ".json"))) for n in ("bag", "const_cs")}
    json.dump(out, open(os.path.join(OUT, "ref_models.json"), "w"))


# Points of the bench grid (BagModel(alpha_n_min=0.005), generic solver) whose FIRST root search fails in the device
# algorithm, so that its result depends on the retries (fsolve_vary / backup scan) being the reference's.
RETRY_POINTS = [(0.4545045045045045, 0.46085585585585587), (0.5211711711711712, 0.10310810810810812),
                (0.6707207207207208, 0.4578828828828829), (0.7472972972972972, 0.06545045045045046),
                (0.7554054054054054, 0.07436936936936937), (0.7617117117117117, 0.46779279279279284),
                (0.8382882882882883, 0.23144144144144146), (0.8851351351351351, 0.43905405405405407),
                (0.4815315315315315, 0.39594594594594595)]


def retry_points():
    """Bubble(model, v_wall, alpha_n) of the reference at RETRY_POINTS: flags, scalars, profile."""
    import logging
    from pttools.models import BagModel
    from pttools.bubble import Bubble
    logging.disable(logging.CRITICAL)
    model = BagModel(alpha_n_min=0.005)
    d = {"points": np.array(RETRY_POINTS), "model": np.array([model.a_s, model.a_b, model.V_s, model.V_b])}
    for i, (vw, an) in enumerate(RETRY_POINTS):
        b = Bubble(model, vw, an)
        d[f"flags_{i}"] = np.array([{"Subsonic deflagration": 0, "Hybrid": 1, "Detonation": 2}[b.sol_type.value],
                                    b.solver_failed, b.solved], dtype=float)
        d[f"scal_{i}"] = np.array([b.wn, b.wm, b.wp, b.v_sh, b.vp, b.vm], dtype=float)
        d[f"v_{i}"], d[f"w_{i}"], d[f"xi_{i}"] = b.v, b.w, b.xi
        print(vw, an, b.sol_type.value, "failed" if b.solver_failed else "ok", b.v.size)
    logging.disable(logging.NOTSET)
    np.savez_compressed(os.path.join(OUT, "ref_retry_points.npz"), **d)


CONSTCS_ALPHA_N_MIN = [  # (css2, csb2, a_s or None, a_b, V_s, alpha_n_min)
    (1 / 3, 1 / 3, 5, 1, 1, 0.05), (1 / 3, 1 / 4, 5, 1, 1, 0.05), (1 / 4, 1 / 3, 5, 1, 1, 0.05), (1 / 4, 1 / 4, 5, 1, 1, 0.05),
    (0.3, 0.28, 5, 1, 1, 0.01), (1 / 3 - 0.01, 1 / 3, 5, 1, 1, 0.1), (0.25, 0.3, None, 1, 1, 0.02), (1 / 3, 0.3, 5, 1, 1, 0.2),
    (0.29, 0.31, 5, 1, 1, 0.05), (0.32, 0.26, 3, 1.1, 1.3, 0.03), (1 / 3, 0.27, None, 1, 1, 0.15), (0.28, 0.28, 5, 1, 1, 0.08),
    (1 / 3, 1 / 3 - 0.01, 5, 1, 1, 0.005), (0.31, 1 / 3, 2, 1, 0.5, 0.12)]


def constcs_alpha_n_min():
    """ConstCSModel(css2, csb2, a_s, a_b, V_s, alpha_n_min=...): the parameters the reference's search
    (models/const_cs.py:225-397) ends with, and the resulting alpha_n_min, T_crit, w_crit."""
    import json
    import logging
    from pttools.models.const_cs import ConstCSModel
    logging.disable(logging.CRITICAL)
    out = []
    for css2, csb2, a_s, a_b, V_s, an in CONSTCS_ALPHA_N_MIN:
        kw = {} if a_s is None else {"a_s": a_s}
        entry = {"css2": css2, "csb2": csb2, "a_s": a_s, "a_b": a_b, "V_s": V_s, "alpha_n_min": an}
        try:
            m = ConstCSModel(css2=css2, csb2=csb2, a_b=a_b, V_s=V_s, alpha_n_min=an, **kw)
            entry["result"] = [float(m.a_s), float(m.a_b), float(m.V_s), float(m.V_b), float(m.alpha_n_min),
                               float(m.T_crit), float(m.w_crit)]
        except Exception as e:       # noqa: BLE001 - the kind of failure is part of the fixture
            entry["raises"] = type(e).__name__
        print(entry)
        out.append(entry)
    logging.disable(logging.NOTSET)
    json.dump(out, open(os.path.join(OUT, "ref_constcs_alpha_n_min.json"), "w"), indent=1)


TASKS.update({"retry_points": retry_points, "constcs_alpha_n_min": constcs_alpha_n_min, "model_tables": model_tables, "fluid_reference": fluid_reference_table, "generic_bubbles": generic_bubbles,
              "generic_spectra": generic_spectra, "suppression_ssm": suppression_ssm})


# ---------------------------------------------------------------------------------------------------------------
# Round 2: the bench grid itself (config 2/5) and config 3, default LSODA and tight LSODA side by side
# ---------------------------------------------------------------------------------------------------------------

GRID_SCALARS = ["wn", "vp", "vm", "vp_tilde", "vm_tilde", "v_sh", "wp", "wm", "kappa", "omega", "kinetic_energy_fraction",
                "va_enthalpy_density", "ubarf2", "wbar"]
Y_DEC = 4            # spectra are stored at y[::Y_DEC]
PROFILE_DEC = 24     # profiles are stored at every PROFILE_DEC-th sample (probes against the other curve)


class _tight_lsoda:
    """Runs the reference with scipy.integrate.odeint at rtol = atol = 1e-12 instead of its default 1.49e-8 (the
    reference passes no tolerances, bubble/integrate.py:180): the difference to the default run estimates the
    reference's own integrator noise."""

    def __enter__(self):
        import scipy.integrate as spi
        self.spi, self.orig = spi, spi.odeint

        def tight(func, y0, t, args=(), **kw):
            kw.setdefault("rtol", 1e-12)
            kw.setdefault("atol", 1e-12)
            kw.setdefault("mxstep", 50000)
            return self.orig(func, y0, t, args=args, **kw)
        spi.odeint = tight
        from pttools.bubble import boundary
        boundary.solve_junction_internal.cache_clear()
        return self

    def __exit__(self, *exc):
        self.spi.odeint = self.orig
        return False


def _decimate(a, step=PROFILE_DEC):
    idx = np.unique(np.concatenate([np.arange(0, a.size, step), np.arange(min(4, a.size)), np.arange(max(0, a.size - 4), a.size)]))
    return idx


def _one_point(model, vw, an, spectrum_kwargs):
    """The reference's Bubble + Spectrum at one point -> dict (or {"err": code})."""
    from pttools.bubble import Bubble
    from pttools.omgw0 import Spectrum
    codes = {"Subsonic deflagration": 0, "Hybrid": 1, "Detonation": 2}
    try:
        b = Bubble(model, vw, an)
    except Exception as e:      # noqa: BLE001 - the constructor rejects the point (alpha_n range, solution type)
        return {"err": 1.0, "msg": type(e).__name__}
    if not b.solved:
        return {"err": 2.0}
    s = Spectrum(b, **spectrum_kwargs)
    return {"flags": np.array([codes[b.sol_type.value], b.solver_failed, b.no_solution_found, b.numerical_error,
                               b.unphysical_alpha_plus, b.negative_entropy_flux, b.negative_net_entropy_change,
                               b.v.size], dtype=float),
            "scal": np.array([float(getattr(b, k)) for k in GRID_SCALARS]),
            "v": b.v, "w": b.w, "xi": b.xi, "pow_v": s.pow_v, "pow_gw": s.pow_gw, "omgw0": s.omgw0(),
            "misc": np.array([s.cs, s.source_lifetime_factor, s.F_gw0(), s.f_star0, s.suppression_factor()])}


def _grid_fixture(model, points, spectrum_kwargs, prefix, d):
    """Default-LSODA and tight-LSODA runs of the reference at `points`; stores the default run (decimated) and, per
    point, how far the tight run is from it (= the reference's own integrator noise)."""
    import logging
    sys.path.insert(0, os.path.join(os.path.dirname(OUT), ""))
    from util_compare import curve_errors
    logging.disable(logging.CRITICAL)
    n = len(points)
    d[f"{prefix}points"] = np.array(points)
    d[f"{prefix}scalar_names"] = np.array(GRID_SCALARS)
    err = np.zeros(n)
    flags = np.full((n, 8), np.nan)
    scal = np.full((n, len(GRID_SCALARS)), np.nan)
    misc = np.full((n, 5), np.nan)
    noise = np.full((n, 8), np.nan)   # pow_v, pow_gw, omgw0, curve v, curve w, scalars(max), npts differs, flags differ
    for i, (vw, an) in enumerate(points):
        r = _one_point(model, vw, an, spectrum_kwargs)
        if "err" in r:
            err[i] = r["err"]
            print(prefix, i, vw, an, "no bubble", r)
            continue
        with _tight_lsoda():
            t = _one_point(model, vw, an, spectrum_kwargs)
        from pttools.bubble import boundary
        boundary.solve_junction_internal.cache_clear()
        flags[i], scal[i], misc[i] = r["flags"], r["scal"], r["misc"]
        idx = _decimate(r["v"])
        for k in ("v", "w", "xi"):
            d[f"{prefix}{k}_{i}"] = r[k][idx]
        for k in ("pow_v", "pow_gw", "omgw0"):
            d[f"{prefix}{k}_{i}"] = r[k][::Y_DEC]
        if "err" in t:
            noise[i, 7] = 1.0
            print(prefix, i, vw, an, "tight run has no bubble")
            continue
        rel = lambda a, b: float(np.nanmax(np.abs(a / b - 1))) if np.all(np.isfinite(b)) else np.nan
        ev, ew = curve_errors(t["xi"], t["v"], t["w"], r["xi"], r["v"], r["w"], vw)
        with np.errstate(invalid="ignore", divide="ignore"):
            noise[i] = [rel(t["pow_v"], r["pow_v"]), rel(t["pow_gw"], r["pow_gw"]), rel(t["omgw0"], r["omgw0"]), ev, ew,
                        float(np.nanmax(np.abs(t["scal"] / r["scal"] - 1))), float(t["flags"][7] != r["flags"][7]),
                        float(np.any(t["flags"][:7] != r["flags"][:7]))]
            d[f"{prefix}scal_noise_{i}"] = np.abs(t["scal"] / r["scal"] - 1)
        print(prefix, i, f"{vw:.4f} {an:.4f}", int(r["flags"][0]), int(r["flags"][7]), "noise", " ".join(f"{x:.1e}" for x in noise[i]),
              flush=True)
    d[f"{prefix}err"], d[f"{prefix}flags"], d[f"{prefix}scal"], d[f"{prefix}misc"], d[f"{prefix}noise"] = err, flags, scal, misc, noise
    logging.disable(logging.NOTSET)


def headline_grid():
    """256 points of the bench grid (BASELINE.json configs 2/5): v_wall = linspace(0.05, 0.95, 1000)[iv] x
    alpha_n = linspace(0.005, 0.5, 1000)[ia], a 16 x 16 sub-lattice that includes the edges of the grid;
    BagModel(alpha_n_min=0.005), Bubble + Spectrum(r_star=0.1) at the reference's default sizes."""
    from pttools.models import BagModel
    model = BagModel(alpha_n_min=0.005)
    v_walls, alpha_ns = np.linspace(0.05, 0.95, 1000), np.linspace(0.005, 0.5, 1000)
    iv = np.linspace(0, 999, 16).astype(int)
    ia = np.linspace(0, 999, 16).astype(int)
    pts = [(float(v_walls[a]), float(alpha_ns[b])) for a in iv for b in ia]
    d = {"iv": iv, "ia": ia, "y_dec": np.array([Y_DEC]), "model": np.array([model.a_s, model.a_b, model.V_s, model.V_b])}
    _grid_fixture(model, pts, {"r_star": 0.1}, "", d)
    np.savez_compressed(os.path.join(OUT, "ref_headline_grid.npz"), **d)
    print("ref_headline_grid.npz", len(pts))


def config3_grid():
    """Config 3 as the reference's example builds it (examples/const_cs/const_cs_gw.py:138-184):
    ConstCSModel(css2, csb2, a_s=5, a_b=1, V_s=1, alpha_n_min=0.05) for css2, csb2 in {1/3, 1/4}, Spectrum(r_star=0.1,
    Tn=200); 4 models x 16 points of the 128 x 128 grid v_wall = linspace(0.05, 0.95, 128) x alpha_n = linspace(0.05,
    0.5, 128)."""
    from pttools.models import ConstCSModel
    v_walls, alpha_ns = np.linspace(0.05, 0.95, 128), np.linspace(0.05, 0.5, 128)
    iv = np.array([3, 40, 77, 120])
    ia = np.array([0, 30, 70, 127])
    d = {"iv": iv, "ia": ia, "y_dec": np.array([Y_DEC])}
    for im, (css2, csb2) in enumerate(((1 / 3, 1 / 3), (1 / 3, 1 / 4), (1 / 4, 1 / 3), (1 / 4, 1 / 4))):
        model = ConstCSModel(css2=css2, csb2=csb2, a_s=5, a_b=1, V_s=1, alpha_n_min=0.05)
        d[f"m{im}__model"] = np.array([css2, csb2, model.a_s, model.a_b, model.V_s, model.V_b, model.T_ref, model.alpha_n_min])
        pts = [(float(v_walls[a]), float(alpha_ns[b])) for a in iv for b in ia]
        _grid_fixture(model, pts, {"r_star": 0.1, "Tn": 200}, f"m{im}__", d)
    np.savez_compressed(os.path.join(OUT, "ref_config3_grid.npz"), **d)
    print("ref_config3_grid.npz")


def noise_snr():
    """LISA noise curves (omgw0/noise.py) on a frequency grid and Spectrum.signal_to_noise_ratio[_instrument]
    (omgw0_ssm.py:143-147) for bag bubbles, Spectrum(r_star=0.1) and Spectrum(r_star=0.01, Tn=500)."""
    from pttools.omgw0 import noise, Spectrum
    from pttools.models import BagModel
    from pttools.bubble import Bubble
    f = np.logspace(-5, 0, 200)
    d = {"f": f}
    for name in ("omega_noise", "omega_ins", "omega_gb", "omega_eb", "S_AE", "S_gb", "P_acc", "S_AE_approx"):
        d[name] = getattr(noise, name)(f)
    d["consts"] = np.array([noise.FT_LISA, noise.F2_LISA, noise.P_oms(), noise.N_acc()])
    model = BagModel(a_s=1.1, a_b=1, V_s=1)
    pts = [(0.5, 0.1), (0.7, 0.1), (0.9, 0.1), (0.62, 0.15)]
    d["points"] = np.array(pts)
    rows = []
    for vw, an in pts:
        b = Bubble(model, vw, an)
        for kw in ({"r_star": 0.1}, {"r_star": 0.01, "Tn": 500}):
            s = Spectrum(b, **kw)
            rows.append([vw, an, kw["r_star"], kw.get("Tn", 100), s.signal_to_noise_ratio(), s.signal_to_noise_ratio_instrument(),
                         s.f()[0], s.f()[-1]])
            print(rows[-1])
    d["snr"] = np.array(rows)
    np.savez_compressed(os.path.join(OUT, "ref_noise_snr.npz"), **d)


def bag_solver_bubbles():
    """Bubble(BagModel, v_wall, alpha_n, use_bag_solver=True): the reference's Bubble on top of its OLD bag solver
    (fluid.py:892-918: sound_shell_bag, w scaled by wn, props.v_and_w_from_solution; v_sh, vm_sh, vm_tilde_sh are NaN;
    alpha_n within 0.05 of alpha_n_max_deflagration_bag: NaN arrays with solver_failed).  Profiles decimated; scalars
    in full."""
    from pttools.models import BagModel
    from pttools.bubble import Bubble
    from pttools.bubble.boundary import SolutionType
    codes = {SolutionType.SUB_DEF: 0, SolutionType.HYBRID: 1, SolutionType.DETON: 2, SolutionType.ERROR: -1}
    model = BagModel(a_s=1.1, a_b=1, V_s=1)
    pts = [(0.3, 0.05), (0.5, 0.1), (0.44, 0.04), (0.62, 0.1), (0.7, 0.15), (0.8, 0.05), (0.92, 0.2), (0.55, 0.3),
           (0.4, 0.33)]
    from pttools.bubble.alpha import alpha_n_max_deflagration_bag
    pts.append((0.3, float(alpha_n_max_deflagration_bag(0.3)) - 0.02))      # "high alpha_n": NaN arrays, solver_failed
    names = ["wn", "vp", "vm", "vp_tilde", "vm_tilde", "wp", "wm", "wm_sh", "v_cj", "kappa", "omega", "kinetic_energy_fraction",
             "kinetic_energy_density", "thermal_energy_density", "va_enthalpy_density", "ubarf2", "wbar", "ebar",
             "va_trace_anomaly_diff", "va_thermal_energy_density_diff", "va_entropy_density_diff", "mean_adiabatic_index"]
    d = {"points": np.array(pts), "scalar_names": np.array(names),
         "model": np.array([model.a_s, model.a_b, model.V_s, model.V_b])}
    for i, (vw, an) in enumerate(pts):
        b = Bubble(model, vw, an, use_bag_solver=True)
        d[f"flags_{i}"] = np.array([codes[b.sol_type], b.solved, b.solver_failed, b.no_solution_found, b.numerical_error,
                                    b.unphysical_alpha_plus, b.negative_entropy_flux, b.negative_net_entropy_change,
                                    b.v.size], dtype=float)
        if b.solved and not b.solver_failed:
            idx = _decimate(b.v)
            d[f"v_{i}"], d[f"w_{i}"], d[f"xi_{i}"] = b.v[idx], b.w[idx], b.xi[idx]
            d[f"scal_{i}"] = np.array([float(getattr(b, k)) for k in names])
            d[f"nan_scal_{i}"] = np.array([float(b.v_sh), float(b.vm_sh), float(b.vm_tilde_sh)])
        print(i, vw, an, b.sol_type, "solved", b.solved, "failed", b.solver_failed, b.v.size)
    np.savez_compressed(os.path.join(OUT, "ref_bag_solver_bubbles.npz"), **d)
    print("ref_bag_solver_bubbles.npz")


TASKS.update({"headline_grid": headline_grid, "config3_grid": config3_grid, "noise_snr": noise_snr,
              "bag_solver_bubbles": bag_solver_bubbles})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    try:
        import pttools  # noqa: F401
    except ImportError:
        sys.exit("the reference is not importable: eval $(tools/refenv/setup_ref.sh) first")
    for name, fn in TASKS.items():
        if args.only in (None, name):
            fn()
