"""CPU: LISA noise curves and signal_to_noise_ratio (pttools_b200/noise.py, host closed forms) against values produced by
the reference's pttools/omgw0/noise.py (tools/make_golden.py::noise_snr), and the trapezoid-weight form the device
reduction uses against the reference formula."""
import os

import numpy as np

from pttools_b200 import noise

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_noise_curves_match_the_reference():
    d = np.load(os.path.join(GOLDEN, "ref_noise_snr.npz"))
    f = d["f"]
    for name in ("omega_noise", "omega_ins", "omega_gb", "omega_eb", "S_AE", "S_gb", "P_acc", "S_AE_approx"):
        np.testing.assert_allclose(getattr(noise, name)(f), d[name], rtol=1e-13, err_msg=name)
    np.testing.assert_allclose([noise.FT_LISA, noise.F2_LISA, noise.P_oms(), noise.N_acc()], d["consts"], rtol=1e-15)


def test_snr_weight_form_equals_the_reference_formula():
    rng = np.random.default_rng(0)
    f = np.logspace(-5, -1, 300)
    sig = 1e-11 * np.exp(-np.log(f / 1e-3) ** 2) * rng.uniform(0.5, 1.5, f.size)
    n1, n2 = noise.omega_noise(f), noise.omega_ins(f)
    w = noise.snr_weights(f, (n1, n2))
    for k, n in enumerate((n1, n2)):
        want = noise.signal_to_noise_ratio(f, sig, n)
        np.testing.assert_allclose(np.sqrt(np.sum(w[k] * sig ** 2)), want, rtol=1e-13)
    # f_min / f_max and a separate noise grid: the reference's index conventions (noise.py:31-50)
    a = noise.signal_to_noise_ratio(f, sig, n1, f_min=1e-4, f_max=1e-2)
    i0, i1 = np.argmax(f >= 1e-4), np.argmax(f >= 1e-2)
    np.testing.assert_allclose(a, np.sqrt(noise.LISA_OBS_TIME * np.trapezoid((sig / n1)[i0:i1] ** 2, f[i0:i1])), rtol=1e-14)
    fn = np.logspace(-4.5, -1.5, 100)
    b = noise.signal_to_noise_ratio(f, sig, noise.omega_noise(fn), f_noise=fn)
    assert np.isfinite(b) and b > 0


def test_suppression_triangulation_reproduces_scipy_linear_interpolation():
    """The triangle tables the device kernel walks (Suppression._triangles) against LinearNDInterpolator (= griddata
    linear, suppression.py:82-87), evaluated on the host with the kernel's own formula."""
    from pttools_b200 import omgw0
    sup = omgw0.DEFAULT_SUPPRESSION
    tri, val = sup._triangles()
    rng = np.random.default_rng(1)
    vw, an = rng.uniform(0.2, 1.0, 400), rng.uniform(0.0, 0.7, 400)
    want = sup.points(vw, an)
    got = np.full(vw.size, np.nan)
    best = np.full(vw.size, -np.inf)
    for q, v in zip(tri, val):
        x0, y0, x1, y1, x2, y2 = q
        det = (y1 - y2) * (x0 - x2) + (x2 - x1) * (y0 - y2)
        l0 = ((y1 - y2) * (vw - x2) + (x2 - x1) * (an - y2)) / det
        l1 = ((y2 - y0) * (vw - x2) + (x0 - x2) * (an - y2)) / det
        l2 = 1 - l0 - l1
        lmin = np.minimum(l0, np.minimum(l1, l2))
        take = (lmin >= -1e-12) & (lmin > best)
        best[take] = lmin[take]
        got[take] = (l0 * v[0] + l1 * v[1] + l2 * v[2])[take]
    assert np.isfinite(want).sum() > 50 and np.isnan(want).sum() > 50
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    ok = np.isfinite(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-12, atol=1e-15)
