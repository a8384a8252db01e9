"""CPU: the per-thread device algorithm (csrc/*.cuh compiled with g++ into a TEST-ONLY emulation library) against
the oracle.  This exercises exactly the source the CUDA kernels compile, without a GPU; the product never loads it."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu"))
import emu  # noqa: E402
from oracle import shell_bag as sb  # noqa: E402
from util_compare import curve_errors  # noqa: E402


def test_integrator_accuracy():
    from scipy.integrate import solve_ivp
    for (v0, w0, xi0, te) in [(0.2519, 1.0, 0.5, -50), (0.42, 1.0, 0.7, -50), (0.2058, 0.5376, 0.7, -50), (0.2519, 1.0, 0.5, -3.3)]:
        rc, v, w, xi = emu.integrate(v0, w0, xi0, 1 / 3, te, 2000)
        assert rc == 2000
        t = np.linspace(0, te, 2000)
        ref = solve_ivp(lambda t_, y_: sb.rhs(y_, t_, 1 / 3), (0, te), [v0, w0, xi0], method="DOP853", rtol=1e-13, atol=1e-15, t_eval=t)
        assert np.abs(v - ref.y[0]).max() < 1e-9
        assert np.abs(w / ref.y[1] - 1).max() < 1e-9
        assert np.abs(xi - ref.y[2]).max() < 1e-9


def test_alpha_n_max():
    for vw in (0.05, 0.3, 0.5, 0.577, 0.58, 0.7, 0.9):
        got, st = emu.anmax(vw)
        assert st == 0
        np.testing.assert_allclose(got, sb.alpha_n_max_deflagration(vw), rtol=5e-7)


def test_alpha_plus_random_points():
    rng = np.random.default_rng(2)
    for _ in range(25):
        vw, an = rng.uniform(0.05, 0.95), rng.uniform(0.005, 0.5)
        am, _ = emu.anmax(vw)
        rc, ap, st, it = emu.find_alpha_plus(vw, an, 5000, am)
        st_o = sb.identify_solution_type(vw, an)
        assert st == st_o
        if st_o == sb.ERROR:
            assert rc == 1
            continue
        assert rc == 0 and it < 20
        np.testing.assert_allclose(ap, sb.find_alpha_plus(vw, an), rtol=5e-7)


@pytest.mark.parametrize("v_wall,alpha_n", [(0.5, 0.1), (0.7, 0.1), (0.9, 0.1), (0.44, 0.0046), (0.58, 0.3), (0.05, 0.3)])
def test_profiles(v_wall, alpha_n):
    am, _ = emu.anmax(v_wall)
    rc, ap, st, _ = emu.find_alpha_plus(v_wall, alpha_n, 5000, am)
    rc2, v, w, xi, nw = emu.bag_profile(v_wall, ap, st, 5000)
    assert rc == 0 and rc2 == 0
    vo, wo, xio = sb.sound_shell_bag(v_wall, alpha_n)
    ev, ew = curve_errors(xi, v, w, xio, vo, wo, v_wall)
    assert ev < 1e-6 and ew < 1e-6
    np.testing.assert_allclose(xi[-2], xio[-2], rtol=1e-6)
    np.testing.assert_allclose(w[0], wo[0], rtol=1e-6)
    assert nw == sb.find_v_index(xi, v_wall)


def test_bag_machine_is_bit_identical_to_the_nested_formulation():
    """BagMachine (the resumable form the kernels run: one integration call site) against bag_find_alpha_plus +
    bag_emit_profile (the nested form) - same arithmetic in the same order, so identical bits."""
    rng = np.random.default_rng(7)
    pts = [(0.5, 0.1), (0.7, 0.1), (0.9, 0.1), (0.44, 0.0046), (0.62, 0.15), (0.92, 0.0046), (0.3, 0.4), (0.05, 0.3),
           (0.2, 0.6)] + [(rng.uniform(0.05, 0.95), rng.uniform(0.005, 0.5)) for _ in range(16)]
    for vw, an in pts:
        am, _ = emu.anmax(vw)
        rc_r, ap, typ, it = emu.find_alpha_plus(vw, an, 1500, am)
        rs, rp, ap2, typ2, it2, v2, w2, x2, nw2, _, _ = emu.bag_machine(vw, an, 1500, am)
        assert (rc_r, typ, it) == (rs, typ2, it2) and (ap == ap2 or (ap != ap and ap2 != ap2))
        if rc_r == 0 and typ >= 0:
            rc_p, v, w, x, nw = emu.bag_profile(vw, ap, typ, 1500)
            assert rc_p == rp and nw == nw2
            for a, b in ((v, v2), (w, w2), (x, x2)):
                assert np.array_equal(a, b, equal_nan=True)
            _, rp3, _, _, _, v3, w3, x3, nw3, _, _ = emu.bag_machine(vw, an, 1500, am, mode=1, ap_in=ap, sol_type_in=typ, w_n=1.7)
            rc_p4, v4, w4, x4, nw4 = emu.bag_profile(vw, ap, typ, 1500, 1.7)
            assert rp3 == rc_p4 and nw3 == nw4 and np.array_equal(w3, w4, equal_nan=True) and np.array_equal(v3, v4, equal_nan=True)
