"""CPU: the plan builder of the C ABI (ptt_ssm_plan_create_host, csrc/ptt_plan.cu - what a C caller of ptt_sweep_spectra
uses) against the numpy plan of engine.py, whose expressions are the reference's (spectrum.py:169-186, 503-515, 473;
calculators.py:100, 136-153).  Host arrays only: no device is touched."""
import numpy as np
import pytest

from pttools_b200 import engine


@pytest.mark.parametrize("case", [
    dict(y=np.logspace(-1, 3, 1000), cs=1 / np.sqrt(3), nt=10000, nzl=10000, nuc="exponential"),
    dict(y=np.logspace(-1, 3, 120), cs=0.5, nt=800, nzl=900, nuc="simultaneous"),
    dict(y=np.logspace(0, 2.5, 77), cs=0.55, nt=333, nzl=501, nuc="exponential", nz_int=777),
])
def test_c_plan_matches_numpy_plan(case):
    y, cs, nt, nzl, nuc = case["y"], case["cs"], case["nt"], case["nzl"], case["nuc"]
    cp = engine.CPlan(y, cs=cs, nt=nt, n_z_lookup=nzl, nuc_type=nuc, nz_int=case.get("nz_int", 0), upload=False)
    info = cp.info()
    eps = 1e-8
    lo, hi = y.min() * 0.5 * (1.0 - cs) / cs - eps, y.max() * 0.5 * (1.0 + cs) / cs + eps
    z_lookup = 10.0 ** np.linspace(np.log10(lo), np.log10(hi), nzl)
    np.testing.assert_allclose(cp.array("z_lookup"), z_lookup, rtol=1e-15)
    passes = [engine.SdvPass(y, nt, nuc, 1.0), engine.SdvPass(z_lookup, nt, nuc, 1.0)]
    assert info["n_pass"] == 2 and info["seg"] == [0, passes[0].qT.size, passes[0].qT.size + passes[1].qT.size]
    for i, p in enumerate(passes):
        assert info["z_tile"][i] == p.z_tile and info["win_entries"][i] == p.win_entries
        np.testing.assert_allclose(info["l0_dl"][2 * i:2 * i + 2], [p.l0, p.dl], rtol=1e-14)
        for k, rtol in (("qT", 1e-15), ("z", 1e-15), ("zterm", 2e-15), ("T", 1e-15), ("LT", 1e-15)):
            a = cp.array(f"p{i}_{k}")
            assert a.shape == getattr(p, k).shape, k
            np.testing.assert_allclose(a, getattr(p, k), rtol=rtol, atol=1e-13, err_msg=k)
        # W = trapezoid weight * T^6 nu(T): the weights difference neighbouring T's (1 ulp of T -> 1e-13 of the weight)
        np.testing.assert_allclose(cp.array(f"p{i}_W"), p.W, rtol=2e-12)
        # (alpha, beta) of the closed-form weights of K6': sums of W over a few samples, some cancel to ~0
        ab, ref = cp.array(f"p{i}_omega_ab").reshape(-1, 2), p.omega_ab
        np.testing.assert_allclose(ab, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())
    gw = engine.build_gw_tables(z_lookup, y, cs, nz_int=case.get("nz_int"))
    for k in ("L1", "L2", "Z1", "Z2", "G", "yterm"):
        np.testing.assert_allclose(cp.array("gw_" + k), gw[k], rtol=1e-10, atol=1e-13, err_msg=k)
    np.testing.assert_allclose(info["gw_prefactor"], gw["prefactor"], rtol=1e-14)
    assert info["gw_win_entries"] == gw["win_entries"]
    frac = np.concatenate([engine.blend_fraction(p.qT) for p in passes])
    np.testing.assert_allclose(cp.array("frac_all"), frac, atol=1e-13)
    np.testing.assert_allclose(info["z_exact_max"], engine.z_exact_max(np.concatenate([p.qT for p in passes]), frac), rtol=1e-14)
    np.testing.assert_array_equal(cp.array("xi_re"), np.linspace(0, 1 - 1 / 2000, 2000))


def test_c_plan_argument_errors():
    from pttools_b200 import _lib
    with pytest.raises(_lib.PttError, match="positive"):
        engine.CPlan(np.array([1.0, -2.0, 3.0]), upload=False)
    with pytest.raises(ValueError):
        engine.CPlan(np.logspace(-1, 1, 10), nuc_type="other", upload=False)
    with pytest.raises(_lib.PttError):
        engine.CPlan(np.logspace(-1, 1, 10), cs=1.5, upload=False)
    cp = engine.CPlan(np.logspace(-1, 1, 10), nt=50, n_z_lookup=40, upload=False)
    with pytest.raises(KeyError):
        cp.array("nope")
    with pytest.raises(_lib.PttError):
        cp.sweep_view()                     # not uploaded: no device view
