"""GPU: the reference's own known-answer tables for Bubble thermodynamics through the model-independent solver
(tests/bubble/test_thermo.py:66-134 of the reference: lecture-notes values, Hindmarsh-Hijazi values, and regression
tables for BagModel and ConstCSModel), at the reference's tolerances (1.3e-2 ... 1.8e-2, 6.8e-3 for ubarf)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TABLES = {
    "lecture_notes": dict(
        model=("bag", dict(a_s=1.1, a_b=1, V_s=1)),
        alpha_n=[0.1, 0.1, 0.1], v_wall=[0.4, 0.7, 0.8],
        kappa=[0.189, 0.452, 0.235], omega=[0.815, 0.559, 0.769], ke_frac=[0.0172, 0.0411, 0.0213],
        ubarf=[0.119, 0.184, 0.133]),
    "hindmarsh_hijazi": dict(
        model=("bag", dict(a_s=1.1, a_b=1, V_s=1)),
        alpha_n=[0.578, 0.151, 0.091], v_wall=[0.5, 0.7, 0.77],
        kappa=[0.610, 0.522, 0.264], omega=[0.395, 0.491, 0.744], ke_frac=[0.223, 0.0684, 0.022]),
    "bag": dict(
        model=("bag", dict(a_s=1.1, a_b=1, V_s=1)),
        alpha_n=np.repeat([0.1, 0.2, 0.3], 3), v_wall=np.tile([0.3, 0.7, 0.8], 3),
        kappa=[0.1227, 0.4512, 0.2346, 0.2141, 0.5645, 0.4630, 0.2881, 0.6245, 0.5786],
        omega=[0.8773, 0.5574, 0.7683, 0.7854, 0.4408, 0.5496, 0.7113, 0.3794, 0.4305],
        ke_frac=[1.12012152e-02, 4.10228822e-02, 2.13351667e-02, 3.58402474e-02, 9.40854610e-02, 7.71674313e-02,
                 6.67623755e-02, 1.44117154e-01, 1.33526554e-01]),
    "const_cs": dict(
        model=("const_cs", dict(css2=1 / 3 - 0.01, csb2=1 / 3, a_s=1.5, a_b=1, V_s=1)),
        alpha_n=np.repeat([0.15, 0.2, 0.3], 3), v_wall=np.tile([0.3, 0.7, 0.8], 3),
        kappa=[1.75263237e-01, 5.29811721e-01, 3.60041399e-01, 2.18903275e-01, 5.74135884e-01, 4.66461842e-01,
               2.93811401e-01, 6.33146311e-01, 5.83362278e-01],
        omega=[8.24598271e-01, 4.77202010e-01, 6.57197385e-01, 7.80795161e-01, 4.31561444e-01, 5.47128825e-01,
               7.05720787e-01, 3.71190282e-01, 4.26371978e-01],
        ke_frac=[2.24107249e-02, 6.85518085e-02, 4.69331415e-02, 3.57922063e-02, 9.49107619e-02, 7.75841920e-02,
                 6.65943334e-02, 1.44939257e-01, 1.34176160e-01]),
}


@pytest.mark.parametrize("name", list(TABLES))
def test_reference_thermo_tables(name):
    from pttools_b200.bubble import Bubble
    from pttools_b200.models import BagModel, ConstCSModel
    t = TABLES[name]
    kind, kw = t["model"]
    model = BagModel(**kw) if kind == "bag" else ConstCSModel(**kw)
    bubbles = [Bubble(model, v_wall=v, alpha_n=a) for v, a in zip(t["v_wall"], t["alpha_n"])]
    np.testing.assert_allclose([b.kappa for b in bubbles], t["kappa"], rtol=1.5e-2)
    np.testing.assert_allclose([b.omega for b in bubbles], t["omega"], rtol=1.3e-2)
    np.testing.assert_allclose([b.kappa + b.omega for b in bubbles], 1, rtol=1.8e-2)
    np.testing.assert_allclose([b.kinetic_energy_fraction for b in bubbles], t["ke_frac"], rtol=1.5e-2)
    if "ubarf" in t:
        np.testing.assert_allclose([np.sqrt(b.ubarf2) for b in bubbles], t["ubarf"], rtol=6.8e-3)


def test_batched_sweep_reproduces_the_tables():
    """The same tables through the batched operator (one launch per model), tighter against the per-point objects."""
    from pttools_b200 import bubble as bub
    from pttools_b200.models import BagModel, ConstCSModel
    for name in ("bag", "const_cs"):
        t = TABLES[name]
        kind, kw = t["model"]
        model = BagModel(**kw) if kind == "bag" else ConstCSModel(**kw)
        order = np.argsort(t["v_wall"], kind="stable")
        bb = bub.sweep_bubbles(model, np.asarray(t["v_wall"], float)[order], np.asarray(t["alpha_n"], float)[order])
        np.testing.assert_allclose(bb.scal["kappa"].cpu().numpy(), np.asarray(t["kappa"])[order], rtol=1.5e-2)
        np.testing.assert_allclose(bb.scal["omega"].cpu().numpy(), np.asarray(t["omega"])[order], rtol=1.3e-2)
        np.testing.assert_allclose(bb.scal["kinetic_energy_fraction"].cpu().numpy(), np.asarray(t["ke_frac"])[order],
                                   rtol=1.5e-2)
