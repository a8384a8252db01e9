"""GPU: edge cases of the sweep operators - empty and single-point sweeps, a sweep in which no point has a solution,
unsorted and duplicated points, a ragged last chunk - checked for shape, NaN policy and order-independence."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

Y = np.logspace(-1, 3, 60)
KW = dict(y=Y, r_star=0.1, nt=400, n_z_lookup=500)


def _model():
    from pttools_b200.models import BagModel
    return BagModel(a_s=1.1, a_b=1, V_s=1)


def test_empty_sweeps():
    from pttools_b200 import engine, sweep
    res = sweep.sweep_spectra(_model(), np.zeros(0), np.zeros(0), **KW)
    assert res.pow_gw.shape == (0, Y.size) and res.omgw0.shape == (0, Y.size) and res.status.numel() == 0
    sh = engine.sweep_shells_bag(np.zeros(0), np.zeros(0))
    assert len(sh) == 0
    res = sweep.power_gw_scaled_bag_sweep(np.zeros(0), np.zeros(0), Y, npt=(500, 200, 300))
    assert res.pow_gw.shape == (0, Y.size)


def test_single_point_equals_batch_row():
    from pttools_b200 import sweep
    m = _model()
    vw, an = np.array([0.44, 0.44, 0.62, 0.8]), np.array([0.05, 0.1, 0.1, 0.07])
    full = sweep.sweep_spectra(m, vw, an, **KW)
    for i in range(vw.size):
        one = sweep.sweep_spectra(m, vw[i:i + 1], an[i:i + 1], **KW)
        np.testing.assert_allclose(one.omgw0.cpu().numpy()[0], full.omgw0.cpu().numpy()[i], rtol=1e-12)
        assert int(one.status[0]) == int(full.status[i]) == 0


def test_sweep_without_any_solution_and_mixed_validity():
    from pttools_b200 import sweep
    m = _model()
    # alpha_n above alpha_n_max_deflagration(v_wall) (0.338 at 0.05, 0.47 at 0.3): no solution anywhere
    res = sweep.sweep_spectra(m, np.full(40, 0.05), np.linspace(0.36, 0.5, 40), **KW)
    assert np.all(res.status.cpu().numpy() != 0) and np.all(np.isnan(res.pow_gw.cpu().numpy()))
    # a v_wall row whose tail has no solution, long enough for the grouped path: NaN exactly on the failed points
    an = np.linspace(0.2, 0.45, 64)
    res = sweep.sweep_spectra(m, np.full(64, 0.05), an, **KW)
    st = res.status.cpu().numpy()
    om = res.pow_gw.cpu().numpy()        # (omgw0 is NaN here anyway: v_wall = 0.05 is outside the suppression data's hull)
    assert 0 < (st == 0).sum() < 64
    assert np.all(np.isfinite(om[st == 0])) and np.all(np.isnan(om[st != 0]))
    # the same points one by one (gather kernel, no trimming of the failed tail) give the same numbers
    for i in (0, 5, int(np.nonzero(st == 0)[0][-1])):
        one = sweep.sweep_spectra(m, np.array([0.05]), an[i:i + 1], **KW)
        np.testing.assert_allclose(one.pow_gw.cpu().numpy()[0], om[i], rtol=1e-9)


def test_unsorted_duplicated_points_and_ragged_chunks():
    from pttools_b200 import sweep
    m = _model()
    rng = np.random.default_rng(1)
    vw = rng.choice([0.3, 0.5, 0.7, 0.9], size=150)
    an = rng.uniform(0.02, 0.3, size=150)
    vw[10], an[10] = vw[3], an[3]                                   # duplicate
    a = sweep.sweep_spectra(m, vw, an, chunk=64, **KW)              # ragged last chunk, v_wall groups cut by chunks
    order = np.argsort(vw, kind="stable")
    b = sweep.sweep_spectra(m, vw[order], an[order], chunk=4096, **KW)
    oa, ob = a.omgw0.cpu().numpy(), b.omgw0.cpu().numpy()
    np.testing.assert_allclose(oa[order], ob, rtol=1e-10, equal_nan=True)
    np.testing.assert_array_equal(oa[10], oa[3])
    np.testing.assert_array_equal(a.status.cpu().numpy()[order], b.status.cpu().numpy())
