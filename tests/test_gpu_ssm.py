"""GPU parity: Sound Shell Model chain (CUDA through the C ABI) against the oracle and the reference-run fixtures."""
import os

import numpy as np
import pytest

from oracle import shell_bag as sb
from oracle import ssm

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SPEC_RTOL = 1e-5      # north_star: spectra and omgw0 amplitudes within 1e-5 relative


def _rows_from_profiles(profiles, torch):
    cap = max(p[0].size for p in profiles)
    P = len(profiles)
    rows = [torch.zeros((P, cap), dtype=torch.float64, device="cuda") for _ in range(3)]
    for i, prof in enumerate(profiles):
        for r, a in zip(rows, prof):
            r[i, : a.size] = torch.as_tensor(a, device="cuda")
    off = torch.zeros(P, dtype=torch.int32, device="cuda")
    npts = torch.as_tensor([p[0].size for p in profiles], dtype=torch.int32, device="cuda")
    return rows, off, npts


def test_sin_transform_reference_fixture():
    import torch
    from pttools_b200 import engine
    d = np.load(os.path.join(GOLDEN, "ref_sin_transform.npz"))
    f = d["f"]
    P, n = f.shape
    xi = np.tile(d["xi"], (P, 1))
    out = engine.sin_transform_batch(d["z"], xi, f, np.zeros(P, dtype=np.int32), np.full(P, n, dtype=np.int32))
    out = out.cpu().numpy()
    for p in range(P):
        want = d[f"st_{p}"]
        np.testing.assert_allclose(out[p], want, rtol=1e-9, atol=1e-13 * np.abs(want).max())


def test_sin_transform_ragged_and_edge_cases():
    import torch
    from pttools_b200 import engine
    rng = np.random.default_rng(3)
    profs, zs = [], np.logspace(-1, np.log10(49.0), 64)      # all-exact branch (no z above the threshold)
    for n in (17, 130, 1025):
        xi = np.sort(rng.uniform(0, 1, n))
        f = np.where((xi > 0.2) & (xi < 0.8), rng.uniform(0.1, 1, n), 0.0)
        profs.append((f, f, xi))
    rows, off, npts = _rows_from_profiles(profs, torch)
    out = engine.sin_transform_batch(zs, rows[2], rows[0], off, npts).cpu().numpy()
    for p, (f, _, xi) in enumerate(profs):
        np.testing.assert_allclose(out[p], ssm.sin_transform(zs, xi, f), rtol=1e-9, atol=1e-14)
    # all-asymptotic branch
    zs_hi = np.logspace(2, 3, 16)
    out = engine.sin_transform_batch(zs_hi, rows[2], rows[0], off, npts).cpu().numpy()
    for p, (f, _, xi) in enumerate(profs):
        np.testing.assert_allclose(out[p], ssm.sin_transform(zs_hi, xi, f), rtol=1e-9, atol=1e-14)


@pytest.mark.parametrize("mode", ["oracle_profiles", "gpu_profiles"])
def test_a2_spec_den_and_gw_bag_chain(mode):
    """power_gw_scaled_bag / power_v_bag / a2_e_conserving_bag at the fixture sizes (npt = 2000, 500, 1000)."""
    import torch
    from pttools_b200 import engine
    d = np.load(os.path.join(GOLDEN, "ref_bag_spectra.npz"))
    z, npt = d["z"], tuple(int(x) for x in d["npt"])
    pts = d["points"]
    vw, an = pts[:, 0], pts[:, 1]
    if mode == "gpu_profiles":
        shells = engine.sweep_shells_bag(vw, an, n_xi=npt[0])
        assert np.all(shells.status.cpu().numpy() == 0)
        tol = SPEC_RTOL
    else:
        # decouple spectrum parity from shell parity: feed the oracle's own profiles
        profs = [sb.sound_shell_bag(a, b, npt[0]) for a, b in pts]
        rows, off, npts_t = _rows_from_profiles(profs, torch)
        shells = engine.ShellBatch(engine.dev_f64(vw), engine.dev_f64(an), None, None, None, rows[0], rows[1], rows[2],
                                   off, npts_t, None, npt[0])
        tol = 1e-8
    theta_s = 0.75 * an          # e = 3/4 w + theta_s (1 - phase), w_n = 1  (quantities.py:38-54)
    pp = engine.point_params(vw, sb.CS0, 0.75, 0.75, theta_s, 0.0)
    # The reference's own LSODA noise (rtol 1.5e-8) moves the trim indices of weak shells, e.g. (0.44, 0.0046):
    # re-running the reference algorithm with tight LSODA tolerances changes ITS pow_gw by 3.1e-5 there.  So the GPU
    # is held to 1e-5 against the noise-free restatement, and to the reference fixture within the fixture's own noise.
    tight = {}
    if mode == "gpu_profiles":
        with sb.lsoda_tolerances(1e-12, 1e-13):
            for i, (a, b) in enumerate(pts):
                prof = sb.sound_shell_bag(a, b, npt[0])
                tight[i] = {nuc: ssm.power_gw_scaled_bag(z, (a, b, nuc, (1.,)), npt, profile=prof)
                            for nuc in ("exponential", "simultaneous")}
                tight[i]["pow_v"] = ssm.power_v_bag(z, (a, b, "exponential", (1.,)), npt, profile=prof)
    for nuc in ("exponential", "simultaneous"):
        plan = engine.SsmPlan(z, cs=sb.CS0, nt=npt[1], n_z_lookup=npt[2], nxi=npt[0], nuc_type=nuc, mode="bag")
        res = engine.spectra_from_shells(plan, shells, pp, keep=True)
        pow_gw = res.pow_gw.cpu().numpy()
        for i in range(len(pts)):
            ref = d[f"pow_gw_{nuc[:3]}_{i}"]
            if mode == "gpu_profiles":
                np.testing.assert_allclose(pow_gw[i], tight[i][nuc], rtol=SPEC_RTOL)
                noise = np.max(np.abs(tight[i][nuc] / ref - 1))
                np.testing.assert_allclose(pow_gw[i], ref, rtol=max(SPEC_RTOL, 2 * noise))
            else:
                np.testing.assert_allclose(pow_gw[i], ref, rtol=tol)
    # power_v_bag: spec_den_v evaluated on z itself
    plan = engine.SsmPlan(z, cs=sb.CS0, nt=npt[1], n_z_lookup=npt[2], nxi=npt[0], mode="ssm", phase_rule=0)
    res = engine.spectra_from_shells(plan, shells, pp, keep=True)
    pow_v = res.pow_v.cpu().numpy()
    for i in range(len(pts)):
        ref = d[f"pow_v_exp_{i}"]
        if mode == "gpu_profiles":
            # same rule as for pow_gw: 1e-5 against the noise-free restatement, the fixture within twice its own noise
            np.testing.assert_allclose(pow_v[i], tight[i]["pow_v"], rtol=SPEC_RTOL)
            noise = np.max(np.abs(tight[i]["pow_v"] / ref - 1))
            np.testing.assert_allclose(pow_v[i], ref, rtol=max(SPEC_RTOL, 2 * noise))
        else:
            np.testing.assert_allclose(pow_v[i], ref, rtol=tol)


def test_a2_parts_against_reference_fixture():
    import torch
    from pttools_b200 import engine, _lib
    import ctypes as C
    d = np.load(os.path.join(GOLDEN, "ref_bag_spectra.npz"))
    npt = tuple(int(x) for x in d["npt"])
    pts = d["points"]
    profs = [sb.sound_shell_bag(a, b, npt[0]) for a, b in pts]
    rows, off, npts_t = _rows_from_profiles(profs, torch)
    shells = engine.ShellBatch(engine.dev_f64(pts[:, 0]), None, None, None, None, rows[0], rows[1], rows[2], off, npts_t,
                               None, npt[0])
    pp = engine.point_params(pts[:, 0], sb.CS0, 0.75, 0.75, 0.75 * pts[:, 1], 0.0)
    # a plan whose single A^2 grid is the fixture's z_a2: build through the generic helper
    zq = d["z_a2"]
    plan = engine.A2OnlyPlan(zq, nxi=npt[0], phase_rule=0)
    a2, fp2, lam2 = engine.a2_batch(plan, shells, pp, want_parts=True)
    a2, fp2, lam2 = a2.cpu().numpy(), fp2.cpu().numpy(), lam2.cpu().numpy()
    for i in range(len(pts)):
        np.testing.assert_allclose(a2[i], d[f"a2_{i}"], rtol=1e-8, atol=1e-30)
        np.testing.assert_allclose(fp2[i], d[f"fp2_{i}"], rtol=1e-8, atol=1e-30)
        np.testing.assert_allclose(lam2[i], d[f"lam2_{i}"], rtol=1e-8, atol=1e-30)


def test_failed_shell_gives_nan_spectrum():
    from pttools_b200 import engine
    z = np.logspace(-1, 3, 50)
    shells = engine.sweep_shells_bag([0.2, 0.5], [0.45, 0.1], n_xi=500)
    pp = engine.point_params([0.2, 0.5], sb.CS0, 0.75, 0.75, [0.75 * 0.45, 0.075], 0.0)
    plan = engine.SsmPlan(z, nt=200, n_z_lookup=300, nxi=500, mode="bag")
    res = engine.spectra_from_shells(plan, shells, pp)
    pw = res.pow_gw.cpu().numpy()
    assert np.all(np.isnan(pw[0]))
    assert np.all(np.isfinite(pw[1])) and np.all(pw[1] > 0)


def test_moment_and_direct_sine_transforms_agree():
    """The block-moment evaluation of the trapezoid sum against one sine per (z, xi) pair, on real shells."""
    from pttools_b200 import engine
    vw = np.array([0.5, 0.7, 0.9, 0.44, 0.62])
    an = np.array([0.1, 0.1, 0.1, 0.0046, 0.15])
    shells = engine.sweep_shells_bag(vw, an, n_xi=5000)
    pp = engine.point_params(vw, sb.CS0, 0.75, 0.75, 0.75 * an, 0.0)
    z = np.logspace(-3, 4, 3000)
    a = engine.a2_batch(engine.A2OnlyPlan(z, phase_rule=0), shells, pp, want_parts=True)
    b = engine.a2_batch(engine.A2OnlyPlan(z, phase_rule=0, force_direct=True), shells, pp, want_parts=True)
    for x, y in zip(a, b):
        x, y = x.cpu().numpy(), y.cpu().numpy()
        # |A|^2 involves f' = df/dz on a log grid (amplification ~1/dlog z) and spans 20 decades: compare against the
        # row maximum as well
        for r in range(x.shape[0]):
            np.testing.assert_allclose(x[r], y[r], rtol=1e-7, atol=1e-10 * np.abs(y[r]).max())


def test_grouped_gemm_spec_den_v_matches_gather_kernel():
    """K6' (banded FP64 matrix product for profiles that share v_wall) against the gather kernel K6."""
    import torch
    from pttools_b200 import engine
    rng = np.random.default_rng(5)
    n_p = 200
    y = np.logspace(-1, 3, 300)
    plan = engine.SsmPlan(y, nt=1500, n_z_lookup=1800, mode="ssm")
    a2 = torch.as_tensor(np.abs(rng.normal(size=(n_p, plan.n_a2))) * np.logspace(0, -12, plan.n_a2)[None, :], device="cuda")
    for v_wall in (0.05, 0.5, 0.95):
        vw = torch.full((n_p,), v_wall, dtype=torch.float64, device="cuda")
        for ipass in (0, 1):
            ref = engine.spec_den_v_batch(plan, ipass, a2, vw).cpu().numpy()
            out = torch.empty((n_p, plan.sdv_plans[ipass].n_z), dtype=torch.float64, device="cuda")
            engine.spec_den_v_group(plan, ipass, a2, v_wall, out)
            np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-11)
    # mixed batch: two big groups and some singletons through the dispatcher
    vw = np.concatenate([np.full(120, 0.3), [0.31, 0.32, 0.33], np.full(77, 0.6)])
    got = engine.spec_den_v_auto(plan, 1, a2, engine.dev_f64(vw)).cpu().numpy()
    ref = engine.spec_den_v_batch(plan, 1, a2, engine.dev_f64(vw)).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-11)


@pytest.mark.parametrize("sizes", [(300, 1800, None), (64, 700, 333), (1000, 10000, None)])
def test_gw_cell_formulation_matches_per_sample_kernel(sizes):
    """K7' (cells of constant np.interp segments, lanes over profiles) against K7 (one lookup pair per sample) and, at
    the small sizes, against the oracle's spec_den_gw_scaled; includes a ragged batch and a NaN profile."""
    import torch
    from pttools_b200 import engine
    n_y, n_zl, nz_int = sizes
    rng = np.random.default_rng(11)
    y = np.logspace(-1, 3, n_y)
    z_lookup = ssm.gen_lookup(y, sb.CS0, n_zl)
    plan = engine.GwOnlyPlan(z_lookup, y, cs=sb.CS0, nz_int=nz_int)
    n_p = 150 if n_zl < 5000 else 70
    shape = z_lookup[None, :] ** rng.uniform(1, 4, (n_p, 1)) / (1 + (z_lookup[None, :] / rng.uniform(1, 30, (n_p, 1))) ** 8)
    pv_h = shape * (1 + 0.3 * np.sin(3 * np.log(z_lookup))[None, :] * rng.uniform(0, 1, (n_p, 1)))
    pv_h[7] = np.nan
    pv = torch.as_tensor(pv_h, device="cuda")
    slf = torch.as_tensor(rng.uniform(0.5, 2, n_p), device="cuda")
    scale = torch.as_tensor(rng.uniform(1e-6, 1e-5, n_p), device="cuda")
    a = engine.spec_den_gw_batch(plan, pv, slf, scale, cells=False)
    b = engine.spec_den_gw_batch(plan, pv, slf, scale, cells=True)
    for x, c in zip(a, b):
        x, c = x.cpu().numpy(), c.cpu().numpy()
        assert np.all(np.isnan(c[7]))
        ok = np.arange(n_p) != 7
        np.testing.assert_allclose(c[ok], x[ok], rtol=1e-11)
    if n_zl < 5000:
        sd = b[0].cpu().numpy()
        for i in (0, 3, 149):
            ref = ssm.spec_den_gw_scaled(z_lookup, pv_h[i], y, sb.CS0, source_lifetime_factor=float(slf[i]),
                                         nz_int=nz_int)[0]
            np.testing.assert_allclose(sd[i], ref, rtol=1e-10)


def test_sin_transform_operator_moments_against_direct_and_oracle():
    """Config-4 style inputs: f_p(xi) = max(0, cos(xi + phi_p)) on [xi_a, xi_b], shared uniform grid and a
    ragged non-uniform variant; z all-exact and mixed exact/asymptotic.  The block-moment operator against the
    one-sine-per-pair kernel (same sum: 1e-10 of the row maximum) and against the oracle's sin_transform."""
    import torch
    from pttools_b200 import engine
    rng = np.random.default_rng(0)
    P, n_xi = 48, 2048
    phi, xa = rng.uniform(0, 1, P), rng.uniform(0.05, 0.5, P)
    xb = xa + rng.uniform(0.05, 0.45, P)
    xi_u = np.linspace(0, 1, n_xi)
    rows_xi = np.zeros((P, n_xi + 7))
    rows_f = np.zeros_like(rows_xi)
    off = rng.integers(0, 7, P)
    npts = np.full(P, n_xi)
    for p in range(P):
        if p % 2:       # ragged, non-uniform
            npts[p] = rng.integers(300, n_xi)
            x = np.sort(rng.uniform(0, 1, npts[p]))
        else:
            x = xi_u
        rows_xi[p, off[p]:off[p] + npts[p]] = x
        rows_f[p, off[p]:off[p] + npts[p]] = np.maximum(0, np.cos(x + phi[p])) * ((x > xa[p]) & (x < xb[p]))
        rows_xi[p, :off[p]] = -5.0                     # padding must not matter
    for z in (np.logspace(-1, np.log10(50), 512), np.logspace(-1, 3, 512)):
        a = engine.sin_transform_batch(z, rows_xi, rows_f, off, npts, method="direct").cpu().numpy()
        b = engine.sin_transform_batch(z, rows_xi, rows_f, off, npts, method="auto").cpu().numpy()
        c = engine.sin_transform_batch(z, rows_xi, rows_f, off, npts, method="moments").cpu().numpy()
        assert np.array_equal(b, c)                    # auto picked the moment path
        for p in range(P):
            np.testing.assert_allclose(c[p], a[p], rtol=1e-9, atol=1e-10 * np.abs(a[p]).max())
        for p in (0, 1, 7):
            x, f = rows_xi[p, off[p]:off[p] + npts[p]], rows_f[p, off[p]:off[p] + npts[p]]
            np.testing.assert_allclose(c[p], ssm.sin_transform(z, x, f), rtol=1e-8, atol=1e-10 * np.abs(a[p]).max())
    # a profile outside [0, 1] must not be silently wrong: auto falls back to the direct kernel, forced moments -> NaN
    rows_xi[3, off[3] + npts[3] - 1] = 1.5
    z = np.logspace(-1, 1, 64)
    d = engine.sin_transform_batch(z, rows_xi, rows_f, off, npts, method="moments").cpu().numpy()
    assert np.all(np.isnan(d[3])) and np.all(np.isfinite(d[2]))
    e = engine.sin_transform_batch(z, rows_xi, rows_f, off, npts, method="auto").cpu().numpy()
    x, f = rows_xi[3, off[3]:off[3] + npts[3]], rows_f[3, off[3]:off[3] + npts[3]]
    np.testing.assert_allclose(e[3], ssm.sin_transform(z, x, f), rtol=1e-8, atol=1e-12)


def test_pow_specs_reference_table_all_rows():
    """data_compare_nuc-test_reference.txt, the table the reference pins in tests/paper/test_pow_specs.py:27-69 at
    rtol 1.72e-6: the integrals of P_v and P_gw over ln z for 10 shells x 2 nucleation types (Np = (10000, 1000),
    z = logspace(log10 0.2, 3, 5000), spec_den_gw_scaled on z itself).  The oracle reproduces two rows at 1e-9
    (tests/test_oracle_golden.py); here all ten through the kernels, at the reference's own tolerance (measured
    worst case 6.4e-7, printed)."""
    from pttools_b200 import engine
    ref = np.load(os.path.join(GOLDEN, "ref_goldens.npz"))["data_compare_nuc_test_reference"]
    alpha, vw = ref[:, 0].copy(), ref[:, 1].copy()
    z = np.logspace(np.log10(0.2), np.log10(1000), 5000)
    cs = sb.CS0
    y = ssm.logspace(np.log10(z.min() * 2 * cs / (1 - cs)), np.log10(z.max() * 2 * cs / (1 + cs)), z.size)
    shells = engine.sweep_shells_bag(vw, alpha, n_xi=10000)
    assert np.all(shells.status.cpu().numpy() == 0)
    pp = engine.point_params(vw, cs, 0.75, 0.75, 0.75 * alpha, 0.0)
    gplan = engine.GwOnlyPlan(z, y, cs=cs)
    worst = 0.0
    for col, nuc in ((0, "simultaneous"), (1, "exponential")):
        plan = engine.SsmPlan(z, cs=cs, nt=1000, nxi=10000, mode="ssm", phase_rule=0, nuc_type=nuc, only_first_pass=True)
        a2 = engine.a2_batch(plan, shells, pp)
        sdv = engine.spec_den_v_auto(plan, 0, a2, engine.dev_f64(vw))
        pv = (z ** 3 / (2 * np.pi ** 2))[None, :] * sdv.cpu().numpy()
        sd_gw = engine.spec_den_gw_batch(gplan, sdv.contiguous())[0].cpu().numpy()
        pgw = (y ** 3 / (2 * np.pi ** 2))[None, :] * sd_gw
        v2 = np.trapezoid(pv / z, z, axis=1)
        omgw = np.trapezoid(pgw / y, y, axis=1)
        for got, want in ((v2, ref[:, 2 + col]), (omgw, ref[:, 4 + col])):
            worst = max(worst, float(np.abs(got / want - 1).max()))
            np.testing.assert_allclose(got, want, rtol=1.72e-6)
    print("worst relative deviation from data_compare_nuc-test_reference.txt:", worst)


def test_bubble_txt_through_the_kernels():
    """bubble.txt (tests/paper/test_bubble.py:23-37 of the reference, rtol 1e-7 there): sums over z of A^2, f'^2/2 and
    (cs l)^2/2 for three weak (alpha 0.0046) and three intermediate (0.05) shells, Np = (10000, 1000).  Intermediate
    shells: 2e-6 against the table.  Weak shells carry the reference's own LSODA noise (tie policy; up to
    1.5e-5 in these sums): 3e-5 against the table, and 2e-6 against the restatement run with tight LSODA tolerances."""
    from pttools_b200 import engine
    ref = np.load(os.path.join(GOLDEN, "ref_goldens.npz"))["bubble"]
    vw = np.array([0.92, 0.56, 0.44, 0.92, 0.56, 0.44])
    alpha = np.array([0.0046] * 3 + [0.05] * 3)
    z = np.logspace(np.log10(0.2), np.log10(1000), 5000)
    shells = engine.sweep_shells_bag(vw, alpha, n_xi=10000)
    pp = engine.point_params(vw, sb.CS0, 0.75, 0.75, 0.75 * alpha, 0.0)
    a2, fp2, lam2 = engine.a2_batch(engine.A2OnlyPlan(z, nxi=10000, phase_rule=0), shells, pp, want_parts=True)
    got = np.stack([np.full(6, z.sum()), a2.sum(1).cpu().numpy(), fp2.sum(1).cpu().numpy(), lam2.sum(1).cpu().numpy()], axis=1)
    print("relative deviations from bubble.txt:", np.abs(got / ref - 1).max(axis=1))
    np.testing.assert_allclose(got[3:], ref[3:], rtol=2e-6)
    for i in range(3):
        # weak shells: 2e-6 against the restatement run with tight LSODA tolerances, and the table within twice the
        # distance of that restatement from it (the reference's own integrator noise, measured here)
        with sb.lsoda_tolerances(1e-12, 1e-13):
            prof = sb.sound_shell_bag(vw[i], alpha[i], 10000)
            t = ssm.a2_e_conserving_bag(z, vw[i], alpha[i], npt=(10000, 1000), profile=prof)
        tight = np.array([x.sum() for x in t])
        np.testing.assert_allclose(got[i, 1:], tight, rtol=2e-6)
        noise = np.abs(tight / ref[i, 1:] - 1)
        assert np.all(np.abs(got[i, 1:] / ref[i, 1:] - 1) <= np.maximum(2e-6, 2 * noise)), (i, noise)


def test_spec_den_gw_scaled_on_an_arbitrary_lookup_grid():
    """spec_den_gw_scaled takes any ascending z_lookup (np.interp, spectrum.py:361-363): a non-uniform grid and unordered
    y through the binary-search kernel, against the oracle; and the same kernel against the arithmetic-index kernel on a
    log-uniform grid."""
    import torch
    from pttools_b200 import engine, ssmtools
    rng = np.random.default_rng(21)
    cs = 0.55
    z_lookup = np.sort(np.concatenate([np.logspace(-2, 4, 700), rng.uniform(0.5, 50, 300)]))
    z_lookup = z_lookup[np.concatenate(([True], np.diff(z_lookup) > 0))]
    pv = z_lookup ** 3 / (1 + (z_lookup / 7) ** 7) * (1 + 0.2 * np.sin(2 * np.log(z_lookup)))
    y = np.array([30.0, 0.7, 5.0, 120.0, 2.2])                 # not ordered
    got, y_out = ssmtools.spec_den_gw_scaled(z_lookup, pv, y, cs=cs, source_lifetime_factor=0.8, nz_int=900)
    want = ssm.spec_den_gw_scaled(z_lookup, pv, y, cs, source_lifetime_factor=0.8, nz_int=900)[0]
    np.testing.assert_allclose(got, want, rtol=1e-10)
    np.testing.assert_array_equal(y_out, y)
    # y = None: the reference builds it from the lookup range
    got2, y2 = ssmtools.spec_den_gw_scaled(z_lookup, pv, cs=cs)
    want2, y2_ref = ssm.spec_den_gw_scaled(z_lookup, pv, None, cs)
    np.testing.assert_allclose(y2, y2_ref, rtol=1e-13)
    np.testing.assert_allclose(got2, want2, rtol=1e-9)
    # both kernels on a log-uniform grid
    zl = np.logspace(-2, 4, 1500)
    pv2 = torch.as_tensor((zl ** 3 / (1 + (zl / 7) ** 7))[None, :].repeat(3, 0), device="cuda")
    yy = np.logspace(-0.5, 2, 40)
    a = engine.spec_den_gw_batch(engine.GwOnlyPlan(zl, yy, cs=cs), pv2, cells=False)[0].cpu().numpy()
    plan_g = engine.GwOnlyPlan(zl, yy, cs=cs)
    plan_g.general = True
    b = engine.spec_den_gw_batch(plan_g, pv2)[0].cpu().numpy()
    np.testing.assert_allclose(b, a, rtol=1e-10)
