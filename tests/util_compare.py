"""Profile comparison helpers shared by the CPU (host-emulation) and GPU parity tests.

Two shells computed by different integrators do not share a tau-grid (the trim indices flip under 1e-8
perturbations), so the curves are compared geometrically: piecewise between the
discontinuities (behind the wall / wall -> shock), never across the wall or the shock, with cubic interpolation.
Behind the wall of a hybrid the rarefaction wave leaves the wall with infinite slope dv/dxi (the integration starts
on the sonic fixed point), so v(xi) is not interpolable there; the smooth parametrisation is xi(v), w(v).  The error
of a sample is therefore the smaller of the two coordinate distances to the other curve in the (xi, v) plane.
"""
import numpy as np
from scipy.interpolate import CubicSpline


def _segment(xi, v, w, v_wall, side):
    m = ((xi < v_wall) if side < 0 else (xi >= v_wall)) & (v > 0)
    return xi[m], v[m], w[m]


def _strict(x, *ys):
    """Sort by x and keep strictly increasing abscissae."""
    o = np.argsort(x, kind="stable")
    x = x[o]
    keep = np.concatenate(([True], np.diff(x) > 0))
    return (x[keep],) + tuple(y[o][keep] for y in ys)


def curve_errors(xi_g, v_g, w_g, xi_o, v_o, w_o, v_wall, margin=3):
    """Max distance of the oracle's moving-fluid samples from the GPU curve: (err_v relative to max v, err_w relative)."""
    vmax = np.max(v_o)
    err_v, err_w = 0.0, 0.0
    for side in (-1, 1):
        xg, vg, wg = _segment(xi_g, v_g, w_g, v_wall, side)
        xo, vo, wo = _segment(xi_o, v_o, w_o, v_wall, side)
        if xg.size < 4 * margin or xo.size < 4 * margin:
            continue
        # parametrised by v
        pv, px, pw = _strict(vg, xg, wg)
        sx, sw = CubicSpline(pv, px), CubicSpline(pv, pw)
        # parametrised by xi
        qx, qv = _strict(xg, vg)
        sv = CubicSpline(qx, qv)
        sel = (vo > pv[margin]) & (vo < pv[-margin - 1]) & (xo > qx[margin]) & (xo < qx[-margin - 1])
        if not np.any(sel):
            continue
        d_xi = np.abs(sx(vo[sel]) - xo[sel])
        d_v = np.abs(sv(xo[sel]) - vo[sel]) / vmax
        err_v = max(err_v, float(np.max(np.minimum(d_xi, d_v))))
        err_w = max(err_w, float(np.max(np.abs(sw(vo[sel]) / wo[sel] - 1.0))))
    return err_v, err_w
