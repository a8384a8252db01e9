"""GPU: post-processing epilogues of a sweep: suppression factor on the device and LISA SNR of every
point, against the reference's Spectrum.signal_to_noise_ratio[_instrument] (fixture from the reference) and against the
host closed forms."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_suppression_on_device_matches_host_interpolation():
    from pttools_b200 import omgw0
    rng = np.random.default_rng(2)
    vw, an = rng.uniform(0.1, 1.0, 5000), rng.uniform(0.0, 0.8, 5000)
    for sup in (omgw0.NO_HYBRIDS_EXT, omgw0.WITH_HYBRIDS):
        for method in (omgw0.SuppressionMethod.NO_EXT, omgw0.SuppressionMethod.EXT_CONSTANT):
            want = sup.points(vw, an, method)
            got = sup.points_device(vw, an, method).cpu().numpy()
            np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
            ok = np.isfinite(want)
            assert ok.sum() > 500
            np.testing.assert_allclose(got[ok], want[ok], rtol=1e-12, atol=1e-15)


def test_snr_of_a_sweep_matches_the_reference():
    """Spectrum(Bubble(bag, v_wall, alpha_n), r_star, Tn).signal_to_noise_ratio() / _instrument() of the reference at its
    default sizes; through sweep_spectra(snr=True) (device reduction) and through the Spectrum object (host)."""
    from pttools_b200 import sweep
    from pttools_b200.bubble import Bubble
    from pttools_b200.models import BagModel
    from pttools_b200.omgw0 import Spectrum
    d = np.load(os.path.join(GOLDEN, "ref_noise_snr.npz"))
    rows = d["snr"]
    model = BagModel(a_s=1.1, a_b=1, V_s=1)
    for r_star, tn in ((0.1, 100), (0.01, 500)):
        sel = rows[(rows[:, 2] == r_star) & (rows[:, 3] == tn)]
        res = sweep.sweep_spectra(model, sel[:, 0], sel[:, 1], r_star=r_star, Tn=tn, snr=True)
        snr = res.snr.cpu().numpy()
        assert snr.shape == (len(sel), 2)
        # SNR is linear in omgw0: the 1e-5 bar of the spectra carries over
        np.testing.assert_allclose(snr[:, 0], sel[:, 4], rtol=1e-5)
        np.testing.assert_allclose(snr[:, 1], sel[:, 5], rtol=1e-5)
        host = sweep.sweep_spectra(model, sel[:, 0], sel[:, 1], r_star=r_star, Tn=tn, snr=True, mem="host")
        np.testing.assert_allclose(host.snr, snr, rtol=1e-12)
    s = Spectrum(Bubble(model, 0.5, 0.1), r_star=0.1)
    np.testing.assert_allclose([s.signal_to_noise_ratio(), s.signal_to_noise_ratio_instrument(), s.f()[0], s.f()[-1]],
                               rows[0, 4:8], rtol=1e-5)
    # shuffled input order, small chunks
    sel = rows[rows[:, 2] == 0.1]
    perm = np.array([2, 0, 3, 1])
    res = sweep.sweep_spectra(model, sel[perm, 0], sel[perm, 1], r_star=0.1, snr=True, chunk=2)
    np.testing.assert_allclose(res.snr.cpu().numpy()[:, 0], sel[perm, 4], rtol=1e-5)
