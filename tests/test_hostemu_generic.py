"""CPU: the device-side model-independent solver (csrc/ptt_generic.cuh through the TEST-ONLY host emulation) against
fixtures produced by the reference itself (tests/golden/ref_generic_bubbles.npz) and against the oracle."""
import os
import sys
import warnings

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu"))
import emu  # noqa: E402
from oracle import models, shell_generic as sg  # noqa: E402
from util_compare import curve_errors  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = {"bag": models.bag_model(1.1, 1, 1), "cs_quarter_third": models.const_cs_model(1 / 4, 1 / 3, 1.5, 1, 1),
          "cs_third_quarter": models.const_cs_model(1 / 3, 1 / 4, 5, 1, 1),
          "cs_032_third": models.const_cs_model(1 / 3 - 0.01, 1 / 3, 1.5, 1, 1)}


@pytest.fixture(scope="module")
def fixtures():
    d = np.load(os.path.join(GOLDEN, "ref_generic_bubbles.npz"))
    ref = sg.FluidReference(dict(np.load(os.path.join(GOLDEN, "ref_fluid_reference.npz"))))
    return d, ref


def test_oracle_reproduces_reference_runs(fixtures):
    """The oracle's generic solver is pinned by outputs of the reference itself (bit-level agreement)."""
    d, ref = fixtures
    names = list(d["scalar_names"])
    warnings.filterwarnings("ignore")
    for mname in ("bag", "cs_third_quarter"):
        m = MODELS[mname]
        for i, (vw, an) in enumerate(d[f"{mname}__points"][:6]):
            if f"{mname}__v_{i}" not in d:
                continue
            sol = sg.sound_shell_generic(m, vw, an, ref)
            assert sol["v"].size == d[f"{mname}__v_{i}"].size
            np.testing.assert_allclose(sol["v"], d[f"{mname}__v_{i}"], rtol=1e-8, atol=1e-12)
            np.testing.assert_allclose(sol["w"], d[f"{mname}__w_{i}"], rtol=1e-8)
            np.testing.assert_allclose(sol["xi"], d[f"{mname}__xi_{i}"], rtol=1e-8, atol=1e-12)
            sc = dict(zip(names, d[f"{mname}__scal_{i}"]))
            bs = sg.bubble_scalars(m, sol, vw)
            for k in ("kappa", "omega", "kinetic_energy_fraction", "va_enthalpy_density", "ubarf2", "wbar"):
                np.testing.assert_allclose(bs[k], sc[k], rtol=1e-8)
            np.testing.assert_allclose(sol["wm"], sc["wm"], rtol=1e-8)


@pytest.mark.parametrize("mname", list(MODELS))
def test_device_algorithm_matches_reference_runs(fixtures, mname):
    d, ref = fixtures
    names = list(d["scalar_names"])
    m = MODELS[mname]
    for i, (vw, an) in enumerate(d[f"{mname}__points"]):
        if f"{mname}__v_{i}" not in d:
            continue
        v, w, xi = d[f"{mname}__v_{i}"], d[f"{mname}__w_{i}"], d[f"{mname}__xi_{i}"]
        sc = dict(zip(names, d[f"{mname}__scal_{i}"]))
        st = int(d[f"{mname}__flags_{i}"][0])
        wn = m.wn(an)
        g = ref.get(vw, an, st)
        rc, gv, gw, gx, gs = emu.generic_solve((m.css2, m.csb2, m.V_s, m.V_b, m.kind == "bag"), vw, an, st, wn,
                                               sg.v_chapman_jouguet(m, an, wn), (g[0], g[2], g[4] * wn, g[5] * wn))
        assert rc == 0 and gs["solver_failed"] == 0
        ev, ew = curve_errors(gx, gv, gw, xi, v, w, vw)
        assert ev < 1e-6 and ew < 1e-6                      # north_star: profiles within 1e-6 relative
        for k in ("wm", "wp", "v_sh", "vp", "vm", "vp_tilde", "vm_tilde", "wm_sh"):
            if sc[k] != 0:
                np.testing.assert_allclose(gs[k], sc[k], rtol=1e-6)
        np.testing.assert_allclose(gw[0], w[0], rtol=1e-6)
        if st != 2:
            assert gv.size == v.size or abs(v[-4]) < 1e-8     # same tau-grid unless the shock is located in numerical noise


def test_vanishing_shock_points_follow_the_reference():
    """Slow walls with strong transitions (v_wall < 0.1, alpha_n ~ 0.3): the flow runs into the fixed point
    (xi, v) = (cs, 0) and the shock has vanishing strength.  The reference finds its index through LSODA noise; the
    device algorithm takes the first sample that np.isclose()s the fixed point (shock.py:114-121 criterion).  Same root
    and same curve; before this rule these points ended as solver failures."""
    from pttools_b200.models import BagModel
    pm = BagModel(alpha_n_min=0.005)                       # the model of bench.py
    mo = models.bag_model(pm.a_s, pm.a_b, pm.V_s, pm.V_b)
    ref = sg.FluidReference(dict(np.load(os.path.join(GOLDEN, "ref_fluid_reference.npz"))))
    warnings.filterwarnings("ignore")
    for vw, an in ((0.08783783783783783, 0.26414414414414417), (0.08783783783783783, 0.33400900900900904),
                   (0.051341075157952554, 0.29217078105609173)):
        sol = sg.sound_shell_generic(mo, vw, an, ref)
        wn = mo.wn(an)
        g = ref.get(vw, an, 0)
        rc, gv, gw, gx, gs = emu.generic_solve((mo.css2, mo.csb2, mo.V_s, mo.V_b, True), vw, an, 0, wn,
                                               sg.v_chapman_jouguet(mo, an, wn), (g[0], g[2], g[4] * wn, g[5] * wn))
        assert rc == 0 and gs["solver_failed"] == 0 and not sol["solver_failed"]
        np.testing.assert_allclose(gs["wm"], sol["wm"], rtol=1e-6)
        ev, ew = curve_errors(gx, gv, gw, sol["xi"], sol["v"], sol["w"], vw)
        assert ev < 1e-6 and ew < 1e-6


def test_retry_search_follows_the_reference():
    """Points whose first root search fails in the device algorithm (1e-4 of the bench grid).  The reference then varies
    the starting guess (fsolve_vary, speedup/solvers.py:12-55) and, for hybrids, scans 20 starts (fluid.py:731-795);
    the device retry search (generic_rescue_kernel, here its lane-by-lane emulation) makes the same attempts in the same
    order.  Fixture: the reference's own Bubble at these points (tools/make_golden.py::retry_points)."""
    from pttools_b200.models import BagModel
    d = np.load(os.path.join(GOLDEN, "ref_retry_points.npz"))
    pm = BagModel(alpha_n_min=0.005)
    mo = models.bag_model(pm.a_s, pm.a_b, pm.V_s, pm.V_b)
    warnings.filterwarnings("ignore")
    np.testing.assert_allclose([pm.a_s, pm.a_b, pm.V_s, pm.V_b], d["model"], rtol=1e-14)
    ref = sg.FluidReference(dict(np.load(os.path.join(GOLDEN, "ref_fluid_reference.npz"))))
    eos = (pm.css2, pm.csb2, pm.V_s, pm.V_b, True)
    s_coef = ((pm.mu_s * pm.a_s) ** (1 / pm.mu_s), (pm.mu_b * pm.a_b) ** (1 / pm.mu_b))      # s = w / T, T_ref = 1
    np.testing.assert_allclose(pm.s(7.0, 0.0), s_coef[0] * 7.0 ** (1 - 1 / pm.mu_s), rtol=1e-14)
    n_first_failed = 0
    for i, (vw, an) in enumerate(d["points"]):
        code, ref_failed, _ = d[f"flags_{i}"]
        code = int(code)
        wn_ref, wm_ref, wp_ref, vsh_ref = d[f"scal_{i}"][:4]
        wn = float(pm.wn(an))
        np.testing.assert_allclose(wn, wn_ref, rtol=1e-13)
        g = ref.get(vw, an, code)
        guess = (g[0], g[2], g[4] * wn, (0.3 if np.isnan(g[5]) else g[5]) * wn)
        v_cj = float(pm.v_chapman_jouguet(an, wn))
        rc, _, _, _, gs = emu.generic_solve(eos, vw, an, code, wn, v_cj, guess)
        n_first_failed += int(rc != 0 or gs["solver_failed"] == 1)
        win, st, gv, gw, gx, sc = emu.generic_rescue(eos, vw, an, code, wn, v_cj, guess, s_coef)
        if ref_failed:
            # the reference keeps the result of its first search and flags it; so does the device path
            assert win < 0 or sc["solver_failed"] == 1
            continue
        assert win >= 0 and st == 0 and sc["solver_failed"] == 0, (vw, an, win, st)
        assert gv.size == d[f"v_{i}"].size                     # same tau-grid index of the shock
        np.testing.assert_allclose([sc["wm"], sc["wp"], sc["v_sh"]], [wm_ref, wp_ref, vsh_ref], rtol=2e-6)
        ev, ew = curve_errors(gx, gv, gw, d[f"xi_{i}"], d[f"v_{i}"], d[f"w_{i}"], vw)
        assert ev < 1e-6 and ew < 1e-6, (vw, an, ev, ew)
    assert n_first_failed >= 7          # the fixture is about points that need the retries
    # Hybrids near v_wall = v_CJ where every start of the reference's backup scan is rejected by its entropy-flux test
    # (the wall junction is solved with allow_negative_entropy_flux_change = False there, fluid.py:745-750) and the
    # reference ends with solver_failed (profiles/r01_retry_flag_agreement.json): no candidate may win here either.
    for vw, an in ((0.7364864864864865, 0.05554054054054054), (0.6454954954954955, 0.007972972972972973),
                   (0.7644144144144144, 0.08427927927927929)):
        wn = float(pm.wn(an))
        g = ref.get(vw, an, 1)
        guess = (g[0], g[2], g[4] * wn, (0.3 if np.isnan(g[5]) else g[5]) * wn)
        sol = sg.sound_shell_generic(mo, vw, an, ref)
        assert sol["solver_failed"]
        win, st, *_ = emu.generic_rescue(eos, vw, an, 1, wn, float(pm.v_chapman_jouguet(an, wn)), guess, s_coef)
        assert win < 0, (vw, an, win)


def test_oracle_reproduces_reference_retries():
    """The oracle's fsolve_vary and backup scan (oracle/shell_generic.py) against the reference's bubbles at the retry
    points: same flags, same arrays; and the reference-failed hybrids of profiles/r01_retry_flag_agreement.json stay
    failed with the same array length (their backup scan rejects every start)."""
    from pttools_b200.models import BagModel
    d = np.load(os.path.join(GOLDEN, "ref_retry_points.npz"))
    pm = BagModel(alpha_n_min=0.005)
    mo = models.bag_model(pm.a_s, pm.a_b, pm.V_s, pm.V_b)
    ref = sg.FluidReference(dict(np.load(os.path.join(GOLDEN, "ref_fluid_reference.npz"))))
    warnings.filterwarnings("ignore")
    for i, (vw, an) in enumerate(d["points"]):
        sol = sg.sound_shell_generic(mo, vw, an, ref)
        assert bool(sol["solver_failed"]) == bool(d[f"flags_{i}"][1])
        assert sol["v"].size == d[f"v_{i}"].size
        np.testing.assert_allclose(sol["v"], d[f"v_{i}"], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(sol["w"], d[f"w_{i}"], rtol=1e-8)
    for vw, an, n_ref in ((0.7364864864864865, 0.05554054054054054, 5005), (0.8806306306306306, 0.4167567567567568, 5166)):
        sol = sg.sound_shell_generic(mo, vw, an, ref)
        assert sol["solver_failed"] and sol["v"].size == n_ref


def test_shock_search_inside_a_step_is_the_ordered_walk(fixtures):
    """integrate_grid lets the shock locators find their sample inside an integrator step by bisection (and walk the
    approach to the fixed point (cs, 0) on xi alone) instead of visiting every tau-grid sample in order.  Same decisions,
    same samples: bit-identical profiles and scalars against the build with -DPTT_NO_STEP_SEARCH, on thin hybrids next to
    v_CJ, slow walls with vanishing shocks, ordinary points of all three types, for the bag and a ConstCS model, and for the
    old bag solver's machine."""
    _, ref = fixtures
    rng = np.random.default_rng(5)
    cases = []
    for mname in ("bag", "cs_third_quarter"):
        m = MODELS[mname]
        pts = [(vw, an) for vw in (0.06, 0.1266, 0.2977, 0.45, 0.56, 0.62, 0.7, 0.8113, 0.9) for an in (0.0055, 0.05, 0.1601, 0.2993, 0.45)]
        pts += list(zip(rng.uniform(0.05, 0.95, 40), rng.uniform(0.005, 0.5, 40)))
        for vw, an in pts:
            wn = m.wn(an)
            if not np.isfinite(wn):
                continue
            try:
                st = int(sg.solution_type(m, vw, an, wn))
            except Exception:
                continue
            if st < 0:
                continue
            g = ref.get(vw, an, st)
            wm_g = g[5] * wn if np.isfinite(g[5]) else 0.3 * wn
            cases.append(((m.css2, m.csb2, m.V_s, m.V_b, m.kind == "bag"), vw, an, st, wn,
                          sg.v_chapman_jouguet(m, an, wn), (g[0], g[2], g[4] * wn, wm_g)))
    bag_pts = [(vw, an) for vw in (0.3, 0.5, 0.62, 0.75, 0.9) for an in (0.01, 0.1, 0.3)]
    out = {}
    try:
        for variant in ((), ("PTT_NO_STEP_SEARCH",)):
            emu.use_variant(*variant)
            res = [emu.generic_solve(*c) for c in cases]
            bag = []
            for vw, an in bag_pts:
                am, _ = emu.anmax(vw)
                bag.append(emu.bag_machine(vw, an, 5000, am))
            out[variant] = (res, bag)
    finally:
        emu.use_variant()
    (a, abag), (b, bbag) = out[()], out[("PTT_NO_STEP_SEARCH",)]
    assert len(a) > 120 and sum(r[0] == 0 for r in a) > 100
    for (rc0, v0, w0, x0, s0), (rc1, v1, w1, x1, s1) in zip(a, b):
        assert rc0 == rc1
        for p, q in ((v0, v1), (w0, w1), (x0, x1)):
            assert p.shape == q.shape and np.array_equal(p, q, equal_nan=True)
        for k in s0:
            assert s0[k] == s1[k] or (np.isnan(s0[k]) and np.isnan(s1[k]))
    for r0, r1 in zip(abag, bbag):
        for p, q in zip(r0, r1):
            if isinstance(p, np.ndarray):
                assert p.shape == q.shape and np.array_equal(p, q, equal_nan=True)
            else:
                assert p == q or (np.isnan(p) and np.isnan(q))
