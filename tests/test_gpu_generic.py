"""GPU parity: model-independent solver, thermodynamic scalars and the Bubble -> Spectrum -> omgw0 chain against
fixtures produced by the reference itself."""
import os

import numpy as np
import pytest

from util_compare import curve_errors

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _models():
    from pttools_b200.models import BagModel, ConstCSModel
    return {"bag": BagModel(a_s=1.1, a_b=1, V_s=1), "cs_quarter_third": ConstCSModel(1 / 4, 1 / 3, a_s=1.5, a_b=1, V_s=1),
            "cs_third_quarter": ConstCSModel(1 / 3, 1 / 4, a_s=5, a_b=1, V_s=1),
            "cs_032_third": ConstCSModel(1 / 3 - 0.01, 1 / 3, a_s=1.5, a_b=1, V_s=1)}


def test_fluid_reference_table_matches_reference():
    from pttools_b200 import fluid_reference
    tab = np.load(os.path.join(GOLDEN, "ref_fluid_reference.npz"))
    ref = fluid_reference.ref()
    for k in ("vp", "vm", "vp_tilde", "vm_tilde", "wp", "wm"):
        a, b = getattr(ref, k), tab[k]
        both = np.isfinite(a) & np.isfinite(b)
        assert both.sum() > 0.9 * np.isfinite(b).sum()
        if k in ("vm", "vm_tilde", "wm"):
            continue
        np.testing.assert_allclose(a[both], b[both], rtol=2e-6, atol=1e-7)
    assert (np.isfinite(ref.vp) != np.isfinite(tab["vp"])).mean() < 0.01
    # vm, vm_tilde, wm of hybrids are read one tau-step behind the wall, where the curve leaves the sonic point with
    # infinite slope: the reference's own LSODA noise is largest there.  Rule: 2e-6 where the table is not noise-limited;
    # elsewhere within twice the distance between the table and the reference algorithm run with tight LSODA tolerances
    # (measured here on the 40 entries that deviate most).
    from oracle import shell_bag as sb
    from oracle import shell_generic as sg
    for k in ("vm", "vm_tilde", "wm"):
        a, b = getattr(ref, k), tab[k]
        both = np.isfinite(a) & np.isfinite(b)
        dev = np.where(both, np.abs(a - b) / np.maximum(np.abs(b), 1e-7 / 2e-6), 0.0)
        worst = np.argsort(dev.ravel())[::-1][:40]
        for flat in worst:
            ia, iv = np.unravel_index(flat, dev.shape)
            if dev[ia, iv] <= 2e-6:
                break
            vw, an = float(tab["v_wall"][iv]), float(tab["alpha_n"][ia])
            with sb.lsoda_tolerances(1e-12, 1e-13):
                v_, w_, xi_ = sb.sound_shell_bag(vw, an)
                vals = sg.v_and_w_from_solution(v_, w_, xi_, vw, sb.identify_solution_type(vw, an))
            tight = dict(zip(("vp", "vm", "vp_tilde", "vm_tilde", "wp", "wm"), vals))[k]
            noise = abs(tight / b[ia, iv] - 1)
            assert dev[ia, iv] <= max(2e-6, 2 * noise), (k, vw, an, dev[ia, iv], noise)
        rest = np.ones(dev.size, dtype=bool)
        rest[worst] = False
        assert dev.ravel()[rest].max() <= max(2e-6, dev.ravel()[worst[-1]])


@pytest.mark.parametrize("mname", ["bag", "cs_quarter_third", "cs_third_quarter", "cs_032_third"])
def test_generic_bubbles_match_reference_runs(mname):
    from pttools_b200 import bubble
    d = np.load(os.path.join(GOLDEN, "ref_generic_bubbles.npz"))
    names = list(d["scalar_names"])
    model = _models()[mname]
    mm = d[f"{mname}__model"]
    np.testing.assert_allclose([model.T_crit, model.w_crit, model.alpha_n_min], mm[7:10], rtol=1e-12)
    pts = [(i, p) for i, p in enumerate(d[f"{mname}__points"]) if f"{mname}__v_{i}" in d]
    vw = np.array([p[0] for _, p in pts])
    an = np.array([p[1] for _, p in pts])
    batch = bubble.sweep_bubbles(model, vw, an)
    assert np.all(batch.status.cpu().numpy() == 0)
    # the reference's own integrator noise on the difference-type integrals, measured with the restatement of its
    # algorithm run with tight LSODA tolerances (oracle pinned to the reference: tests/test_hostemu_generic.py)
    from oracle import models as omodels, shell_bag as sb, shell_generic as sg
    import warnings
    omodel = {"bag": omodels.bag_model(1.1, 1, 1), "cs_quarter_third": omodels.const_cs_model(1 / 4, 1 / 3, 1.5, 1, 1),
              "cs_third_quarter": omodels.const_cs_model(1 / 3, 1 / 4, 5, 1, 1),
              "cs_032_third": omodels.const_cs_model(1 / 3 - 0.01, 1 / 3, 1.5, 1, 1)}[mname]
    oref = sg.FluidReference(dict(np.load(os.path.join(GOLDEN, "ref_fluid_reference.npz"))))
    tight_cache = {}

    def tight_scalars(k):
        if k not in tight_cache:
            with warnings.catch_warnings(), sb.lsoda_tolerances(1e-12, 1e-13):
                warnings.simplefilter("ignore")
                sol = sg.sound_shell_generic(omodel, vw[k], an[k], oref)
                tight_cache[k] = sg.bubble_scalars(omodel, sol, vw[k])
        return tight_cache[k]

    for k, (i, _) in enumerate(pts):
        v, w, xi = d[f"{mname}__v_{i}"], d[f"{mname}__w_{i}"], d[f"{mname}__xi_{i}"]
        sc = dict(zip(names, d[f"{mname}__scal_{i}"]))
        assert int(batch.sol_type[k]) == int(d[f"{mname}__flags_{i}"][0])
        gv, gw, gx = batch.shells.profile(k)
        ev, ew = curve_errors(gx, gv, gw, xi, v, w, vw[k])
        assert ev < 1e-6 and ew < 1e-6
        np.testing.assert_allclose(float(batch.wn[k]), sc["wn"], rtol=1e-12)
        for name in ("vp", "vm", "vp_tilde", "vm_tilde", "v_sh", "vm_sh", "vm_tilde_sh", "wp", "wm", "wm_sh", "v_cj"):
            if sc[name] != 0:
                # atol: for very weak shocks the reference locates the shock where v ~ 1e-10 (integrator noise)
                np.testing.assert_allclose(float(batch.scal[name][k]), sc[name], rtol=1e-6, atol=2e-8, err_msg=name)
        for name in ("kappa", "omega", "kinetic_energy_fraction", "va_enthalpy_density", "ubarf2", "wbar", "ebar",
                     "va_kinetic_energy_density", "va_trace_anomaly_diff", "va_thermal_energy_density_diff",
                     "va_entropy_density_diff"):
            # the entropy difference integrates s - s_n, a small difference of large numbers: it amplifies the ~1e-7
            # profile differences (the reference's own LSODA noise) more than the other integrals
            # omega, the thermal-energy and the entropy differences integrate w - w_n (or s - s_n), small differences
            # of large numbers: they amplify the ~1e-7 profile differences (the reference's own LSODA noise) by
            # w_n / (w - w_n) ~ 30-100
            got = float(batch.scal[name][k])
            if abs(got / sc[name] - 1) <= 2e-6:
                continue
            assert name in ("va_entropy_density_diff", "omega", "va_thermal_energy_density_diff"), (name, got, sc[name])
            # over the bar: allowed up to twice the reference's own noise on this quantity at this point
            noise = abs(tight_scalars(k)[name] / sc[name] - 1)
            assert abs(got / sc[name] - 1) <= max(2e-6, 2 * noise), (name, vw[k], an[k], got, sc[name], noise)


def test_bubble_object_config1():
    """BASELINE config 1: BagModel(a_s=1.1, a_b=1, V_s=1), v_wall=0.5, alpha_n=0.1 (values measured on the reference)."""
    from pttools_b200.bubble import Bubble
    from pttools_b200.models import SolutionType
    b = Bubble(_models()["bag"], 0.5, 0.1)
    assert b.solved and not b.solver_failed and b.sol_type == SolutionType.SUB_DEF
    assert b.v.size == 203
    np.testing.assert_allclose(b.kappa, 0.2819506859971521, rtol=1e-6)
    np.testing.assert_allclose(b.kinetic_energy_fraction, 0.025616507567699634, rtol=1e-6)


@pytest.mark.parametrize("mname", ["bag", "cs_quarter_third", "cs_third_quarter"])
def test_spectrum_chain_matches_reference_runs(mname):
    """Spectrum(Bubble(model, v_wall, alpha_n), r_star=0.1) at the reference's default sizes: pow_v, pow_gw, omgw0."""
    from pttools_b200 import sweep
    d = np.load(os.path.join(GOLDEN, "ref_generic_spectra.npz"))
    model = _models()[mname]
    pts = d[f"{mname}__points"]
    res = sweep.sweep_spectra(model, pts[:, 0], pts[:, 1], y=d["y"], r_star=0.1).numpy()
    for i in range(len(pts)):
        np.testing.assert_allclose(res.pow_v[i], d[f"{mname}__pow_v_{i}"], rtol=1e-5)
        np.testing.assert_allclose(res.pow_gw[i], d[f"{mname}__pow_gw_{i}"], rtol=1e-5)
        np.testing.assert_allclose(res.omgw0[i], d[f"{mname}__omgw0_{i}"], rtol=1e-5)


def test_spectrum_object_config1():
    from pttools_b200.bubble import Bubble
    from pttools_b200.omgw0 import Spectrum
    d = np.load(os.path.join(GOLDEN, "ref_generic_spectra.npz"))
    b = Bubble(_models()["bag"], 0.5, 0.1)
    s = Spectrum(b, r_star=0.1)
    np.testing.assert_allclose(s.pow_gw, d["bag__pow_gw_0"], rtol=1e-5)
    ref_a2 = d["bag__a2_0"]
    big = ref_a2 > 1e-3 * ref_a2.max()
    np.testing.assert_allclose(s.a2[big], ref_a2[big], rtol=1e-5)
    # far below the peak |A|^2 is a near-cancellation of f' and l (the "z^0 piece", ssm.py:95-98) and carries the
    # profile noise amplified
    np.testing.assert_allclose(s.a2, ref_a2, rtol=2e-4, atol=1e-30)
    np.testing.assert_allclose(s.f(), d["bag__f_0"], rtol=1e-12)
    f_peak, om_peak = s.omgw0_peak()
    np.testing.assert_allclose([f_peak, om_peak], [1.2725153729897776e-04, 4.289465881249297e-11], rtol=1e-5)
    misc = d["bag__misc_0"]
    np.testing.assert_allclose([s.cs, s.source_lifetime_factor], misc[:2], rtol=1e-6)


def test_create_spectra_grid_api():
    from pttools_b200 import analysis
    model = _models()["bag"]
    objs = analysis.create_spectra(model, np.array([0.5, 0.7]), np.array([0.1, 0.15]),
                                   spectrum_kwargs={"r_star": 0.1, "nt": 500, "n_z_lookup": 500,
                                                    "y": np.logspace(-1, 3, 100)}, bubble_kwargs={"n_xi": 2000})
    assert objs.shape == (2, 2)
    assert objs[0, 0].bubble.v_wall == 0.5 and objs[1, 1].bubble.alpha_n == 0.15
    assert np.all(np.isfinite(objs[0, 1].pow_gw)) and objs[0, 1].omgw0().shape == (100,)


def test_grid_facade_values_against_the_reference():
    """create_spectra / create_bubbles / BubbleGridVWAlpha (analysis/parallel.py:89-187, bubble_grid.py:57-80) on a 4 x 4
    sub-lattice of the bench grid, against the reference's own Bubble + Spectrum objects at those points
    (ref_headline_grid.npz): array layout [i_alpha_n, i_v_wall], omgw0 1e-5 (or twice the reference's own LSODA noise
    recorded in the fixture), kappa / omega 1e-6 (same rule), post-functions, None for points the constructor rejects."""
    from pttools_b200 import analysis
    from pttools_b200.models import BagModel
    d = np.load(os.path.join(GOLDEN, "ref_headline_grid.npz"))
    model = BagModel(alpha_n_min=0.005)
    v_all, a_all = np.linspace(0.05, 0.95, 1000), np.linspace(0.005, 0.5, 1000)
    sel_v, sel_a = [2, 6, 9, 14], [0, 3, 7, 12]                 # positions in the fixture's 16 x 16 lattice
    v_walls, alpha_ns = v_all[d["iv"][sel_v]], a_all[d["ia"][sel_a]]
    ydec = int(d["y_dec"][0])
    names = list(d["scalar_names"])

    def peak(s):
        return float(np.nanmax(s.omgw0()))
    peak.return_type, peak.fail_value = np.float64, np.nan

    objs, peaks = analysis.create_spectra(model, v_walls, alpha_ns, func=peak, allow_bubble_failure=True,
                                          spectrum_kwargs={"r_star": 0.1})
    assert objs.shape == peaks.shape == (4, 4)
    grid = analysis.BubbleGridVWAlpha(model, v_walls, alpha_ns)
    kappa, omega, kg = grid.kappa(), grid.omega(), grid.kappa_giese()
    n_checked = 0
    for ja, ka in enumerate(sel_a):
        for jv, kv in enumerate(sel_v):
            i = kv * 16 + ka                                      # fixture order: v_wall-major
            np.testing.assert_allclose(d["points"][i], [v_walls[jv], alpha_ns[ja]], rtol=1e-14)
            s = objs[ja, jv]
            if d["err"][i] == 1:                                  # the reference's constructor raises
                assert s is None and np.isnan(peaks[ja, jv]) and grid.bubbles[ja, jv] is None and np.isnan(kappa[ja, jv])
                continue
            assert s.bubble.v_wall == v_walls[jv] and s.bubble.alpha_n == alpha_ns[ja]
            if d["err"][i] == 2 or d["flags"][i][1]:
                continue
            noise = d["noise"][i]
            ref = d[f"omgw0_{i}"]
            got = s.omgw0()[::ydec]
            if np.all(np.isfinite(ref)):
                assert np.max(np.abs(got / ref - 1)) <= max(1e-5, 2 * noise[2]), (v_walls[jv], alpha_ns[ja])
                np.testing.assert_allclose(peaks[ja, jv], np.nanmax(s.omgw0()), rtol=0)
            sc = dict(zip(names, d["scal"][i]))
            sn = dict(zip(names, d[f"scal_noise_{i}"]))
            for nm, arr in (("kappa", kappa), ("omega", omega)):
                assert abs(arr[ja, jv] / sc[nm] - 1) <= max(1e-6, 2 * sn[nm]), (nm, v_walls[jv], alpha_ns[ja])
            b = grid.bubbles[ja, jv]
            np.testing.assert_allclose(kg[ja, jv], 4 * b.kinetic_energy_density / (3 * b.alpha_theta_bar_n * b.wn), rtol=1e-14)
            n_checked += 1
    assert n_checked >= 10
    with pytest.raises(ValueError):                               # without allow_bubble_failure the rejection propagates
        analysis.create_bubbles(model, v_walls[:1], np.array([0.001]))


def test_spectrum_chain_through_grouped_matrix_product():
    """Same fixtures as above with the banded-GEMM spec_den_v forced on (runs of equal v_wall of any length)."""
    from pttools_b200 import engine, sweep
    d = np.load(os.path.join(GOLDEN, "ref_generic_spectra.npz"))
    model = _models()["bag"]
    pts = d["bag__points"]
    old = engine.GROUP_MIN
    engine.GROUP_MIN = 1
    try:
        res = sweep.sweep_spectra(model, pts[:, 0], pts[:, 1], y=d["y"], r_star=0.1).numpy()
    finally:
        engine.GROUP_MIN = old
    for i in range(len(pts)):
        np.testing.assert_allclose(res.pow_v[i], d[f"bag__pow_v_{i}"], rtol=1e-5)
        np.testing.assert_allclose(res.omgw0[i], d[f"bag__omgw0_{i}"], rtol=1e-5)


def test_unsorted_input_order_is_restored():
    from pttools_b200 import sweep
    model = _models()["bag"]
    vw = np.array([0.7, 0.5, 0.9, 0.5])
    an = np.array([0.1, 0.1, 0.1, 0.2])
    y = np.logspace(-1, 3, 60)
    kw = dict(y=y, r_star=0.1, nt=300, n_z_lookup=400, n_xi=1000)
    res = sweep.sweep_spectra(model, vw, an, **kw).numpy()
    np.testing.assert_array_equal(res.v_wall, vw)
    for i in range(4):
        one = sweep.sweep_spectra(model, vw[i:i + 1], an[i:i + 1], **kw).numpy()
        np.testing.assert_allclose(res.pow_gw[i], one.pow_gw[0], rtol=1e-12)
        np.testing.assert_allclose(res.scalars["kappa"][i], one.scalars["kappa"][0], rtol=1e-12)


def test_retry_search_matches_the_reference_on_points_whose_first_search_fails():
    """Fixture: the reference's Bubble at points of the bench grid where the first root search of the device algorithm
    fails (tools/make_golden.py::retry_points).  With the reference's retries (fsolve_vary's varied starts, the hybrid
    backup scan; generic_rescue_kernel) validity flags, tau-grid sizes and curves agree; without them the points are
    solver failures.  Also through sweep_spectra, where the retries run on a side stream."""
    import torch
    from pttools_b200 import bubble, sweep
    from pttools_b200.models import BagModel
    d = np.load(os.path.join(GOLDEN, "ref_retry_points.npz"))
    model = BagModel(alpha_n_min=0.005)
    vw, an = d["points"][:, 0].copy(), d["points"][:, 1].copy()
    ref_failed = np.array([bool(d[f"flags_{i}"][1]) for i in range(vw.size)])
    off = bubble.sweep_bubbles(model, vw, an, retry="off")
    assert (off.status.cpu().numpy() != 0).sum() >= 7
    bb = bubble.sweep_bubbles(model, vw, an)
    st = bb.status.cpu().numpy()
    np.testing.assert_array_equal(st == 5, ref_failed)
    assert np.all((st == 0) | (st == 5))
    for i in np.nonzero(~ref_failed)[0]:
        v, w, xi = bb.shells.profile(i)
        assert v.size == d[f"v_{i}"].size
        wn_ref, wm_ref, wp_ref, vsh_ref = d[f"scal_{i}"][:4]
        got = [float(bb.scal[k][i]) for k in ("wm", "wp", "v_sh")]
        np.testing.assert_allclose(got, [wm_ref, wp_ref, vsh_ref], rtol=2e-6)
        ev, ew = curve_errors(xi, v, w, d[f"xi_{i}"], d[f"v_{i}"], d[f"w_{i}"], vw[i])
        assert ev < 1e-6 and ew < 1e-6, (vw[i], an[i], ev, ew)
    # thermodynamic scalars of the rescued points are those of their new profiles
    one = bubble.Bubble(model, float(vw[3]), float(an[3]))
    assert one.solved and not one.solver_failed
    np.testing.assert_allclose(float(bb.scal["kappa"][3]), one.kappa, rtol=1e-12)
    # sweep_spectra: retries queued on a side stream, folded in after the chunks
    y = np.logspace(-1, 3, 60)
    res = sweep.sweep_spectra(model, vw, an, y=y, r_star=0.1, nt=600, n_z_lookup=700)
    np.testing.assert_array_equal(res.status.cpu().numpy() == 5, ref_failed)
    ok = torch.as_tensor(~ref_failed, device="cuda")
    assert torch.isfinite(res.pow_gw[ok]).all() and (res.pow_gw[ok] > 0).all()
    np.testing.assert_allclose(res.scalars["kappa"].cpu().numpy(), bb.scal["kappa"].cpu().numpy(), rtol=1e-12)
    from pttools_b200.omgw0 import Spectrum
    s = Spectrum(one, y=y, r_star=0.1, nt=600, n_z_lookup=700)
    np.testing.assert_allclose(res.pow_gw[3].cpu().numpy(), s.pow_gw, rtol=1e-9)
    # several chunks: the chunk with the fewest retried points runs first, the others see the rescued profiles in place
    big_vw = np.concatenate([vw, np.full(7, 0.5), vw[:4] + 1e-3])
    big_an = np.concatenate([an, np.linspace(0.05, 0.2, 7), an[:4]])
    whole = sweep.sweep_spectra(model, big_vw, big_an, y=y, r_star=0.1, nt=600, n_z_lookup=700)
    parts = sweep.sweep_spectra(model, big_vw, big_an, y=y, r_star=0.1, nt=600, n_z_lookup=700, chunk=5)
    np.testing.assert_array_equal(whole.status.cpu().numpy(), parts.status.cpu().numpy())
    fin = torch.isfinite(whole.pow_gw)
    assert torch.equal(fin, torch.isfinite(parts.pow_gw)) and fin[:vw.size][ok].all()
    np.testing.assert_allclose(parts.pow_gw[fin].cpu().numpy(), whole.pow_gw[fin].cpu().numpy(), rtol=1e-9)
    np.testing.assert_allclose(parts.omgw0[fin].cpu().numpy(), whole.omgw0[fin].cpu().numpy(), rtol=1e-9)
    np.testing.assert_allclose(parts.scalars["kappa"].cpu().numpy(), whole.scalars["kappa"].cpu().numpy(), rtol=1e-12,
                               equal_nan=True)
