"""CPU: host-side plan logic of the spectrum kernels (pttools_b200/engine.py, numpy only) against the oracle.

K6' evaluates spec_den_v as a banded matrix product P_v(z_i) = 2/(b_R v_w)^6 sum_e Omega[i, e] A2[e]; the weights
Omega come from a closed form per piece (engine._omega_pieces) and the band from engine._group_band.  Both are host
computations a reference maintainer would have to reproduce, so they are pinned here
against the reference's own formulation: np.interp + np.trapezoid (oracle.ssm.spec_den_v_core = spectrum.py:451-486)."""
import numpy as np
import pytest

from oracle import ssm
from pttools_b200 import engine


def _omega_dense(p: engine.SdvPass, v_wall: float):
    """Omega[n_z, n_q] the way sdv_weights_kernel builds it: closed form on interior knots, sample loop on the two
    clamp knots (np.interp's left / right fill)."""
    b_r = (8.0 * np.pi) ** (1.0 / 3.0)
    bv = b_r * v_wall
    shift = -np.log10(bv) / p.dl
    n_q, nt = p.qT.size, p.nt
    ab = engine._omega_pieces(p.LT, p.T, p.W, p.dl)
    lt0 = p.LT[0]
    inv_dlt = (nt - 1) / (p.LT[-1] - lt0)
    s = p.z / bv
    ub = p.zterm + shift
    e = np.arange(n_q)
    x = e[None, :] - ub[:, None]
    piece = np.zeros(x.shape, dtype=np.int64)
    for sgn in (-1, 0, 1):
        yy = (x - sgn - lt0) * inv_dlt
        piece += np.clip(np.ceil(np.where(yy > 0, yy, 0.0)).astype(np.int64), 0, nt)
    om = ab[piece, 0] + ab[piece, 1] * (s[:, None] / p.qT[None, :])
    # clamp knots: brute force over the samples
    for i in range(p.z.size):
        q = s[i] * p.T
        er = np.floor(ub[i] + p.LT).astype(np.int64)
        for knot in (0, n_q - 1):
            acc = 0.0
            for j in range(nt):
                if er[j] < 0:
                    eff, w0, w1 = 0, p.W[j], 0.0
                elif er[j] >= n_q - 1:
                    eff, w0, w1 = n_q - 2, 0.0, p.W[j]
                else:
                    eff = er[j]
                    t = (q[j] - p.qT[eff]) / (p.qT[eff + 1] - p.qT[eff])
                    w0, w1 = p.W[j] * (1 - t), p.W[j] * t
                if eff == knot:
                    acc += w0
                elif eff == knot - 1:
                    acc += w1
            om[i, knot] = acc
    return om


@pytest.mark.parametrize("v_wall", [0.07, 0.44, 0.92])
def test_closed_form_weights_reproduce_interp_trapezoid(v_wall):
    rng = np.random.default_rng(5)
    z = np.logspace(-1, 2.5, 70)
    nt = 180
    p = engine.SdvPass(z, nt, "exponential", 1.0)
    a2 = np.exp(rng.normal(size=p.qT.size)) * p.qT ** 2 / (1 + p.qT ** 6)          # positive, many decades
    want = ssm.spec_den_v_core(z, a2, p.qT, v_wall, nt=nt)
    om = _omega_dense(p, v_wall)
    b_r = (8.0 * np.pi) ** (1.0 / 3.0)
    got = 2.0 / (b_r * v_wall) ** 6 * om @ a2
    np.testing.assert_allclose(got, want, rtol=2e-12)
    # the band handed to the GEMM holds every non-zero weight of its 128-row tile
    e_base, kw = engine._group_band(p, v_wall)
    assert kw % 16 == 0
    for t, e0 in enumerate(e_base):
        rows = om[128 * t:128 * (t + 1)]
        nz = np.nonzero(np.any(rows != 0.0, axis=0))[0]
        assert nz.size and nz.min() >= e0 and nz.max() < e0 + kw, (t, e0, kw, nz.min(), nz.max())


def test_qt_lookup_grid_and_blend_fraction_match_the_oracle():
    z = np.logspace(-1, 3, 257)
    qt, l0, dl = engine.qt_lookup_grid(z)
    np.testing.assert_array_equal(qt, ssm.qt_lookup_grid(z))
    np.testing.assert_allclose(np.log10(qt[1:] / qt[:-1]), dl, rtol=1e-9)
    # sin_transform's blend (calculators.py:136-153): exact below thresh - pi, asymptotic above thresh, cubic in between
    fr = engine.blend_fraction(z)
    assert np.all(fr[z <= 50 - np.pi] == 0) and np.all(fr[z > 50] == 1)
    mid = (z > 50 - np.pi) & (z <= 50)
    assert mid.sum() >= 2 and np.all(np.diff(fr[mid]) > 0) and fr[mid][0] == 0 and fr[mid][-1] == 1
    xi = np.linspace(0, 1, 400)
    f = np.where((xi > 0.5) & (xi < 0.7), np.sin(8 * xi), 0.0)
    full = ssm.sin_transform(z, xi, f)
    exact = ssm.sin_transform_core(xi, f, z)
    np.testing.assert_allclose(full[fr == 0], exact[fr == 0], rtol=1e-13, atol=1e-16)
