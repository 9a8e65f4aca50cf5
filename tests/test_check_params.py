"""Constructor-time soft validation (smooth_nmf.py:145-237): bad values are PRINTED and reset, never raised.
Where the reference tree is present (the build container) our estimator is compared with the real one -- printed text
and resulting parameters -- for every branch; elsewhere the expectations recorded below are used."""
import contextlib
import io

import numpy as np
import pytest

from oracle import ref_import

CASES = {
    "lambda_type": dict(lambda_L="a"),
    "linesearch_type": dict(linesearch=1),
    "mu_type": dict(mu="x"),
    "eps_type": dict(epsilon_reg=None),
    "algo_type": dict(algo=3),
    "simplex_types": dict(simplex_H=1, simplex_W=0),
    "dicotomy_type": dict(dicotomy_tol="1e-3"),
    "gamma_type": dict(gamma=(1, 2)),
    "verbose_type": dict(verbose="yes"),
    "debug_type": dict(debug=1),
    "l2_type": dict(l2=1),
    "k_type": dict(n_components=2.5),
    "algo_unknown": dict(algo="newton"),
    "lambda_negative": dict(lambda_L=-1.0),
    "eps_nonpositive": dict(epsilon_reg=0.0),
    "mu_negative": dict(mu=np.array([0.1, -0.2])),
    "both_simplex": dict(simplex_H=True, simplex_W=True),
    "linesearch_l2": dict(linesearch=True, l2=True, algo="l2_surrogate", lambda_L=2.0),
    "linesearch_lambda0": dict(linesearch=True, lambda_L=0.0),
    "l2_wrong_algo": dict(l2=True, algo="log_surrogate"),
    "clean": dict(lambda_L=2.0, mu=0.05, simplex_H=True, simplex_W=False),
}
PARAMS = ("lambda_L", "linesearch", "mu", "epsilon_reg", "algo", "simplex_H", "simplex_W", "dicotomy_tol", "gamma",
          "verbose", "debug", "l2", "n_components")


def _build(cls, kw):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        est = cls(**kw)
    return est, buf.getvalue()


def _same(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.array_equal(np.asarray(a), np.asarray(b))
    return a == b and type(a) is type(b)


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("tag", sorted(CASES))
def test_check_params_matches_reference(tag):
    from espm_b200 import SmoothNMF
    ref = ref_import.load_reference()
    ours, out_ours = _build(SmoothNMF, CASES[tag])
    theirs, out_ref = _build(ref.estimators.SmoothNMF, CASES[tag])
    for name in PARAMS:
        assert _same(getattr(ours, name), getattr(theirs, name)), (tag, name, getattr(ours, name), getattr(theirs, name))
    assert out_ours == out_ref, tag


def test_check_params_expectations():
    """Reference-free pins of the same behaviour (values observed from the reference in the build container)."""
    from espm_b200 import SmoothNMF
    est, out = _build(SmoothNMF, CASES["both_simplex"])
    assert est.simplex_W is True and est.simplex_H is False and "applied to W and not to H" in out
    est, out = _build(SmoothNMF, CASES["algo_unknown"])
    assert est.algo == "log_surrogate" and out
    est, out = _build(SmoothNMF, CASES["linesearch_lambda0"])
    assert est.lambda_L == 1 and est.linesearch is True
    est, out = _build(SmoothNMF, CASES["l2_wrong_algo"])
    assert est.l2 is False
    est, out = _build(SmoothNMF, CASES["clean"])
    assert out == "" and est.lambda_L == 2.0 and est.simplex_H is True
