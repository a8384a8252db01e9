"""CPU: the reference's own EOS known answers (tests/test_data/models/{bag,const_cs}.json, checked there by
tests/models/base_model.py:41-100) against the host-side models of the product (pttools_b200/models.py) and against
the oracle's models; same inputs, same tolerances as the reference's test."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N = 10
ALPHA_N = np.linspace(0.15, 0.5, 10)
TEMP = np.linspace(1, 100, N)
W1, W2 = TEMP ** 4, TEMP ** 3.9
PHASE = np.array([(1 + (-1) ** i) / 2 for i in range(N)])
# (key, tolerance) as in base_model.py
TOL = {"cs2": dict(rtol=3.4e-4), "p": dict(rtol=2.3e-4), "theta": dict(rtol=1e-7, atol=1.2e-8), "p_temp": dict(rtol=5e-5)}


@pytest.fixture(scope="module")
def tables():
    return json.load(open(os.path.join(GOLDEN, "ref_models.json")))


def _check(model, ref, skip=()):
    calls = {
        "alpha_n": lambda: model.alpha_n(W1), "alpha_plus": lambda: model.alpha_plus(W1, W2),
        "cs2": lambda: model.cs2(W1, PHASE), "e": lambda: model.e(W1, PHASE), "p": lambda: model.p(W1, PHASE),
        "s": lambda: model.s(W1, PHASE), "theta": lambda: model.theta(W1, PHASE),
        "e_temp": lambda: model.e_temp(TEMP, PHASE), "p_temp": lambda: model.p_temp(TEMP, PHASE),
        "s_temp": lambda: model.s_temp(TEMP, PHASE), "temp": lambda: model.temp(W1, PHASE),
        "w": lambda: model.w(TEMP, PHASE), "w_n": lambda: model.wn(ALPHA_N), "critical_temp": lambda: model.critical_temp(),
    }
    for key, fn in calls.items():
        if key in skip:
            continue
        got = np.asarray(fn(), dtype=float)
        want = np.asarray(ref[key], dtype=float)
        np.testing.assert_allclose(np.broadcast_to(got, want.shape), want, **TOL.get(key, dict(rtol=1e-7)), err_msg=key)


def test_product_models_against_reference_tables(tables):
    from pttools_b200.models import BagModel, ConstCSModel
    _check(BagModel(a_s=1.2, a_b=1.1, V_s=1.3), tables["bag"])
    _check(ConstCSModel(a_s=1.2, a_b=1.1, V_s=1.3, css2=1 / 3, csb2=1 / 3), tables["bag"])      # TestConstCSLikeBag
    _check(ConstCSModel(a_s=1.2, a_b=1.1, V_s=1.3, css2=0.4 ** 2, csb2=1 / 3), tables["const_cs"])


def test_oracle_models_against_reference_tables(tables):
    from oracle import models
    _check(models.bag_model(1.2, 1.1, 1.3), tables["bag"])
    _check(models.const_cs_model(0.4 ** 2, 1 / 3, 1.2, 1.1, 1.3), tables["const_cs"])


def test_constcs_alpha_n_min_parameter_search():
    """ConstCSModel(..., alpha_n_min=x) runs the parameter search of const_cs.py:225-397.  Fixture: what the reference
    itself ends with (tools/make_golden.py::constcs_alpha_n_min) for 14 models - direct hits, the V_s/100 retry, the
    two-parameter L-BFGS-B stage and the fall-back to the defaults are all in there."""
    import logging
    from pttools_b200.models import ConstCSModel
    cases = json.load(open(os.path.join(GOLDEN, "ref_constcs_alpha_n_min.json")))
    assert len(cases) >= 14
    logging.disable(logging.CRITICAL)
    try:
        for c in cases:
            kw = {} if c["a_s"] is None else {"a_s": c["a_s"]}
            make = lambda: ConstCSModel(c["css2"], c["csb2"], a_b=c["a_b"], V_s=c["V_s"], alpha_n_min=c["alpha_n_min"], **kw)
            if "raises" in c:
                with pytest.raises(Exception) as info:
                    make()
                assert type(info.value).__name__ == c["raises"]
                continue
            m = make()
            got = [m.a_s, m.a_b, m.V_s, m.V_b, m.alpha_n_min, m.T_crit, m.w_crit]
            np.testing.assert_allclose(got, c["result"], rtol=1e-9, atol=1e-12, err_msg=str(c))
    finally:
        logging.disable(logging.NOTSET)


def test_constcs_critical_temp_rejects_what_the_reference_rejects():
    """model.py:476-557: no T_crit when p_s(T_min) >= p_b(T_min) or when p_s never exceeds p_b."""
    from pttools_b200.models import ConstCSModel
    with pytest.raises(ValueError):
        ConstCSModel(0.3, 0.28, a_s=0.5, a_b=1.0, V_s=1.0)       # a_s < a_b with mu_s < mu_b: p_s stays below p_b
    with pytest.raises(ValueError):
        ConstCSModel(0.3, 0.28, a_s=2.0, a_b=1.0, V_s=0.5, V_b=1.0)      # V_s < V_b
    m = ConstCSModel(0.3, 0.28, a_s=0.5, a_b=1.0, V_s=1.0, allow_invalid=True)
    assert m.T_crit > 0


def test_wn_on_a_grid_solves_each_distinct_alpha_n_once():
    """ConstCSModel.wn(alpha_n) is an iterative solve; on a (v_wall, alpha_n) grid every alpha_n repeats for every wall
    speed and is solved once.  The iteration stops when all its values have converged and the set of values is the same,
    so the result equals the plain solve of the full array bit for bit."""
    import logging
    import pttools_b200.models as M
    logging.disable(logging.CRITICAL)
    try:
        m = M.ConstCSModel(css2=0.3, csb2=0.28, a_s=5, a_b=1, V_s=1, alpha_n_min=0.05)
    finally:
        logging.disable(logging.NOTSET)
    an = np.tile(np.linspace(0.05, 0.5, 97), 13).reshape(13, 97)
    got = m.wn(an)
    plain_unique = M.np.unique
    M.np.unique = lambda a, return_inverse=False: (a, np.arange(a.size)) if return_inverse else a     # no duplicates found
    try:
        want = m.wn(an)
    finally:
        M.np.unique = plain_unique
    ok = np.isfinite(got)                                   # NaN below the model's own alpha_n_min
    assert got.shape == an.shape and ok.sum() > an.size // 2
    assert np.array_equal(got, want, equal_nan=True)
    np.testing.assert_allclose(m.alpha_n(got[ok]), an[ok], rtol=1e-9)
