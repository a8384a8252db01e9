"""GPU: seam B as ONE call of the C ABI (ptt_sweep_spectra, csrc/ptt_sweep.cu) and seam A / the shell sweeps with HOST
buffers (PTT_MEM_HOST), the way INTEGRATION.md binds them (numpy arrays in, numpy arrays out, no torch)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bag():
    from pttools_b200.models import BagModel
    return BagModel(a_s=1.1, a_b=1, V_s=1)


def _points(n_rows=5, per_row=40, seed=0):
    rng = np.random.default_rng(seed)
    v_walls = np.sort(rng.uniform(0.1, 0.92, n_rows))
    vw = np.repeat(v_walls, per_row)
    an = np.tile(np.linspace(0.02, 0.45, per_row), n_rows)
    return vw, an


SMALL = dict(y=np.logspace(-1, 3, 80), r_star=0.1, nt=400, n_z_lookup=500, n_xi=1500)


def test_sweep_operator_chain_equals_operator_by_operator():
    """The one-call sweep against the same kernels driven operator by operator from Python (seam C): shells, thermodynamic
    integrals, A^2, spec_den_v (gather kernel), spec_den_gw -- same arithmetic, so agreement to rounding."""
    import torch
    from pttools_b200 import bubble, engine, sweep
    model = _bag()
    vw, an = _points()
    res = sweep.sweep_spectra(model, vw, an, **SMALL)
    bb = bubble.sweep_bubbles(model, vw, an, n_xi=SMALL["n_xi"])
    np.testing.assert_array_equal(res.status.cpu().numpy(), bb.status.cpu().numpy())
    plan = engine.SsmPlan(SMALL["y"], cs=float(np.sqrt(model.csb2)), nt=SMALL["nt"], n_z_lookup=SMALL["n_z_lookup"], mode="ssm")
    slf = bb.scal["source_lifetime_factor"]                   # computed without r_star by sweep_bubbles: recompute
    nu = bb.scal["nu_gdh2024"]
    slf = 1 / (1 + 2 * nu) * (1 - (1 + 2 * 0.1 / torch.sqrt(bb.scal["kinetic_energy_fraction"])) ** (-1 - 2 * nu))
    scale = engine.dev_f64(sweep.omgw0_scale(vw, an, 0.1))
    old = engine.USE_GROUPED_SDV
    engine.USE_GROUPED_SDV = False
    try:
        ref = engine.spectra_from_shells(plan, bb.shells, bb.pp, slf=slf.contiguous(), scale=scale)
    finally:
        engine.USE_GROUPED_SDV = old
    ok = torch.isfinite(ref.pow_gw[:, 0])
    assert ok.sum() > 100 and torch.equal(ok, torch.isfinite(res.pow_gw[:, 0]))
    for a, b in ((res.pow_gw, ref.pow_gw), (res.pow_v, ref.pow_v), (res.omgw0, ref.omgw0)):
        fin = torch.isfinite(b)
        assert torch.equal(fin, torch.isfinite(a))
        np.testing.assert_allclose(a[fin].cpu().numpy(), b[fin].cpu().numpy(), rtol=1e-10)
    np.testing.assert_allclose(res.scalars["source_lifetime_factor"][ok].cpu().numpy(), slf[ok].cpu().numpy(), rtol=1e-13)


def test_host_buffers_chunking_and_input_order():
    """PTT_MEM_HOST (numpy in, pinned numpy out) == PTT_MEM_DEVICE; small shell / spectrum chunks (several shell chunks,
    both workspace slots, output double buffering) == one chunk; shuffled input order comes back in the caller's order."""
    from pttools_b200 import sweep
    model = _bag()
    vw, an = _points(n_rows=6, per_row=30, seed=1)
    dev = sweep.sweep_spectra(model, vw, an, **SMALL).numpy()
    host = sweep.sweep_spectra(model, vw, an, mem="host", **SMALL)
    assert isinstance(host.pow_gw, np.ndarray) and isinstance(host.status, np.ndarray)
    small = sweep.sweep_spectra(model, vw, an, chunk=25, shell_chunk=50, mem="host", **SMALL)
    perm = np.random.default_rng(2).permutation(vw.size)
    shuf_d = sweep.sweep_spectra(model, vw[perm], an[perm], chunk=64, **SMALL).numpy()
    shuf_h = sweep.sweep_spectra(model, vw[perm], an[perm], chunk=64, mem="host", **SMALL)
    for other, idx in ((host, None), (small, None), (shuf_d, perm), (shuf_h, perm)):
        sel = slice(None) if idx is None else idx
        np.testing.assert_array_equal(np.asarray(other.status), dev.status[sel])
        for name in ("pow_v", "pow_gw", "omgw0"):
            a, b = np.asarray(getattr(other, name)), getattr(dev, name)[sel]
            np.testing.assert_allclose(a, b, rtol=1e-10, equal_nan=True)
        for name in ("kappa", "wm", "v_sh", "source_lifetime_factor", "npts"):
            np.testing.assert_allclose(np.asarray(other.scalars[name]), dev.scalars[name][sel], rtol=1e-12, equal_nan=True)
    assert np.isfinite(dev.pow_gw).all(axis=1).sum() > 100 and (dev.status != 0).any()


def test_chunk_callback_ranges_cover_the_sweep():
    from pttools_b200 import sweep
    model = _bag()
    vw, an = _points(n_rows=6, per_row=30, seed=3)
    seen = []
    res = sweep.sweep_spectra(model, vw, an, chunk=32, on_chunk=lambda a, b, s: seen.append((a, b, bool(s))), **SMALL)
    assert sorted(x[:2] for x in seen) == [(a, min(a + 32, vw.size)) for a in range(0, vw.size, 32)]
    assert all(x[2] for x in seen)
    assert res.stats.launches > 0 and res.stats.a2_profiles == vw.size


def test_plan_built_by_the_library_gives_the_same_spectra():
    """ptt_ssm_plan_create (C++ plan builder) vs the numpy plan of engine.SsmPlan through the same sweep."""
    from pttools_b200 import engine, sweep
    model = _bag()
    vw, an = _points(n_rows=3, per_row=40, seed=4)
    kw = {k: v for k, v in SMALL.items() if k not in ("nt", "n_z_lookup")}
    cs = float(np.sqrt(model.csb2))
    a = sweep.sweep_spectra(model, vw, an, plan=engine.SsmPlan(SMALL["y"], cs=cs, nt=400, n_z_lookup=500), **kw).numpy()
    b = sweep.sweep_spectra(model, vw, an, plan=engine.CPlan(SMALL["y"], cs=cs, nt=400, n_z_lookup=500), **kw).numpy()
    for name in ("pow_v", "pow_gw", "omgw0"):
        np.testing.assert_allclose(getattr(b, name), getattr(a, name), rtol=1e-9, equal_nan=True)


def test_invalid_points_are_marked_not_solved():
    """Points the reference's Bubble constructor rejects (alpha_n < alpha_n_min: model.py:813-823; no solution type:
    transition.py:47-69) come back as PTT_ST_INVALID_INPUT with NaN rows instead of finite spectra."""
    from pttools_b200 import _lib, analysis, sweep
    from pttools_b200.models import BagModel
    model = BagModel(a_s=1.1, a_b=1, V_s=1, alpha_n_min=0.05) if False else BagModel(alpha_n_min=0.05)
    vw = np.array([0.5, 0.5, 0.2, 0.7])
    an = np.array([0.01, 0.1, 0.45, 0.2])          # below alpha_n_min | fine | no solution type | fine
    res = sweep.sweep_spectra(model, vw, an, **SMALL).numpy()
    np.testing.assert_array_equal(res.invalid, [True, False, True, False])
    np.testing.assert_array_equal(res.status == _lib.ST_INVALID_INPUT, [True, False, True, False])
    assert np.isnan(res.pow_gw[[0, 2]]).all() and np.isfinite(res.pow_gw[[1, 3]]).all()
    with pytest.raises(ValueError):
        analysis.create_spectra(model, np.array([0.5]), np.array([0.01, 0.1]), spectrum_kwargs={k: v for k, v in SMALL.items() if k != "n_xi"})
    objs = analysis.create_spectra(model, np.array([0.5, 0.7]), np.array([0.01, 0.1]), allow_bubble_failure=True,
                                   spectrum_kwargs={k: v for k, v in SMALL.items() if k != "n_xi"},
                                   bubble_kwargs={"n_xi": SMALL["n_xi"]})
    assert objs.shape == (2, 2) and objs[0, 0] is None and objs[0, 1] is None and objs[1, 0].bubble.solved
    np.testing.assert_allclose(objs[1, 0].pow_gw, res.pow_gw[1], rtol=1e-10)
    assert objs[1, 0].bubble.v_wall == 0.5 and objs[1, 1].bubble.v_wall == 0.7 and objs[1, 1].bubble.alpha_n == 0.1


def _host_call(name, *args):
    from pttools_b200 import _lib
    _lib.check(getattr(_lib.load(), name)(*args), name)


def test_seam_a_and_shell_sweeps_with_host_buffers():
    """ptt_integrate_batch / ptt_sweep_shells_bag / ptt_sweep_shells_generic called exactly as INTEGRATION.md shows:
    numpy host buffers, PTT_MEM_HOST, default stream; against the PTT_MEM_DEVICE results."""
    import torch
    from pttools_b200 import _lib, bubble, engine
    lib = _lib.load()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    # seam A
    n, n_xi = 6, 700
    v0 = np.linspace(0.05, 0.3, n); w0 = np.ones(n); xi0 = np.linspace(0.4, 0.65, n)
    phase = np.zeros(n); t_end = np.full(n, -30.0)
    out = [np.empty((n, n_xi)) for _ in range(3)]
    status = np.empty(n, dtype=np.int32)
    eos = _lib.Eos(kind=0, css2=1 / 3, csb2=1 / 3, a_s=0, a_b=0, V_s=0, V_b=0, T_ref=1)
    _host_call("ptt_integrate_batch", n, p(v0), p(w0), p(xi0), p(phase), p(t_end), n_xi, C.byref(eos), p(out[0]), p(out[1]),
               p(out[2]), p(status), _lib.MEM_HOST, None)
    dv, dw, dx, dst = engine.integrate_batch(v0, w0, xi0, phase, t_end, n_xi)
    assert np.all(status == 0)
    for h, d in zip(out, (dv, dw, dx)):
        np.testing.assert_array_equal(h, d.cpu().numpy())
    # the scalar API of the reference on top of it: fresh arrays, RuntimeError on failure (integrate.py:81-135)
    v, w, xi, t = bubble.fluid_integrate_param(0.2, 1.0, 0.5, phase=0.0, t_end=-30.0, n_xi=n_xi)
    assert v.shape == (n_xi,) and t[0] == 0.0 and t[-1] == -30.0
    with pytest.raises(RuntimeError, match="Integration failed"):
        bubble.fluid_integrate_param(np.nan, 1.0, 0.5, phase=0.0, t_end=-30.0, n_xi=n_xi)
    # bag shell sweep
    vw = np.array([0.5, 0.7, 0.9, 0.2]); an = np.array([0.1, 0.1, 0.1, 0.45])
    n, n_xi = vw.size, 1200
    cap = lib.ptt_profile_row_capacity(n_xi)
    rows = [np.empty((n, cap)) for _ in range(3)]
    off, npts, sol, st = (np.empty(n, dtype=np.int32) for _ in range(4))
    ap = np.empty(n); scal = np.empty((n, 4))
    _host_call("ptt_sweep_shells_bag", n, p(vw), p(an), n_xi, None, p(rows[0]), p(rows[1]), p(rows[2]), cap, p(off), p(npts),
               p(ap), p(sol), p(scal), p(st), _lib.MEM_HOST, None)
    sh = engine.sweep_shells_bag(vw, an, n_xi=n_xi)
    np.testing.assert_array_equal(st, sh.status.cpu().numpy())
    np.testing.assert_array_equal(npts, sh.npts.cpu().numpy())
    assert st[3] == _lib.ST_NO_SOLUTION and np.all(st[:3] == 0)
    for i in range(3):
        gv, gw, gx = sh.profile(i)
        for h, d in zip(rows, (gv, gw, gx)):
            np.testing.assert_array_equal(h[i, off[i]:off[i] + npts[i]], d)
    # generic shell sweep
    pt = bubble.point_table(_bag(), vw[:3], an[:3])
    cap = lib.ptt_generic_row_capacity(n_xi)
    rows = [np.empty((3, cap)) for _ in range(3)]
    off, npts, st = (np.empty(3, dtype=np.int32) for _ in range(3))
    scal = np.empty((3, 16))
    _host_call("ptt_sweep_shells_generic", 3, p(pt.pg), n_xi, 50.0, 100, 1e-4, p(rows[0]), p(rows[1]), p(rows[2]), cap, p(off),
               p(npts), p(scal), p(st), _lib.MEM_HOST, None)
    bb = bubble.sweep_bubbles(_bag(), vw[:3], an[:3], n_xi=n_xi)
    assert np.all(st == 0)
    np.testing.assert_array_equal(npts, bb.shells.npts.cpu().numpy())
    np.testing.assert_allclose(scal[:, :13], bb.shells.scalars[:, :13].cpu().numpy(), rtol=0, atol=0)
    for i in range(3):
        gv, gw, gx = bb.shells.profile(i)
        np.testing.assert_array_equal(rows[0][i, off[i]:off[i] + npts[i]], gv)


def test_sweep_argument_errors():
    from pttools_b200 import _lib, engine
    lib = _lib.load()
    o = _lib.SweepOut()
    opts = _lib.SweepOpts()
    assert lib.ptt_sweep_spectra(None, C.byref(opts), 0, None, None, None, C.byref(o), None, None, _lib.MEM_HOST, None) == 0
    rc = lib.ptt_sweep_spectra(None, C.byref(opts), 5, None, None, None, C.byref(o), None, None, _lib.MEM_HOST, None)
    assert rc == -1 and b"ptt_sweep_spectra" in lib.ptt_last_error()
    pg = np.zeros((2, 16)); pm = np.zeros((2, 4)); st = np.zeros(2, dtype=np.int32); pw = np.zeros((2, 10))
    o.status = st.ctypes.data; o.pow_gw = pw.ctypes.data
    rc = lib.ptt_sweep_spectra(None, C.byref(opts), 2, pg.ctypes.data, pm.ctypes.data, None, C.byref(o), None, None,
                               _lib.MEM_HOST, None)
    assert rc == -1 and b"plan" in lib.ptt_last_error()
    assert lib.ptt_sweep_release() == 0


def test_sweeps_queued_back_to_back_and_their_stage_times():
    """Device-table sweeps given the host copies of v_wall and the solution type (ptt_sweep_opts_t.host_v_wall /
    host_sol_type) are queued behind one another without waiting; the stage times of sweep k are read with
    ptt_sweep_timing_collect(serial) while later sweeps are in flight.  Same tables as the sweep that fetches pg itself."""
    import torch
    from pttools_b200 import _lib, bubble, engine
    from pttools_b200.sweep import omgw0_scale
    model = _bag()
    vw, an = _points(n_rows=4, per_row=30, seed=3)
    plan = engine.SsmPlan(SMALL["y"], cs=float(np.sqrt(model.csb2)), nt=SMALL["nt"], n_z_lookup=SMALL["n_z_lookup"],
                          mode="ssm", only_lookup_pass=True)
    pt = bubble.point_table(model, vw, an, device=True)
    scale = engine.dev_f64(omgw0_scale(vw, an, 0.1))
    kw = dict(n_xi=SMALL["n_xi"], r_star=0.1, want=("pow_gw", "omgw0"), timing=True)
    plain = engine.sweep_c(plan, pt.pg, pt.pm, scale, **kw)
    queued = [engine.sweep_c(plan, pt.pg, pt.pm, scale, host_cols=(vw, pt.codes), **kw) for _ in range(3)]
    serials = [int(q.stats.timing_serial) for q in queued]
    assert serials == list(range(serials[0], serials[0] + 3)) and int(plain.stats.timing_serial) == serials[0] - 1
    lib = _lib.load()
    st = _lib.SweepStats()
    # the middle one, while the last one may still be running; then everything
    _lib.check(lib.ptt_sweep_timing_collect(serials[1], C.byref(st)), "ptt_sweep_timing_collect")
    assert int(st.timing_serial) == serials[1] and st.stage_ms[0] > 0 and st.stage_calls[0] == 1
    engine.finish_timing()
    for q in queued + [plain]:
        assert q.stats.stage_ms[0] > 0 and int(q.stats.stage_calls[0]) == 1
    torch.cuda.synchronize()
    for q in queued:
        assert torch.equal(q.status, plain.status)
        assert torch.equal(torch.isnan(q.omgw0), torch.isnan(plain.omgw0))
        # (the moment sums of the sine transform are accumulated with shared-memory atomics: equal to rounding)
        np.testing.assert_allclose(torch.nan_to_num(q.omgw0).cpu().numpy(), torch.nan_to_num(plain.omgw0).cpu().numpy(),
                                   rtol=1e-11)
    assert lib.ptt_sweep_timing_collect(10 ** 9, C.byref(st)) != 0            # unknown serial: an error, not a hang
    with pytest.raises(ValueError):
        engine.sweep_c(plan, pt.pg, pt.pm, scale, host_cols=(vw[:-1], pt.codes), **kw)
