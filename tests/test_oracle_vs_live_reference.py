"""The oracle against the UNMODIFIED reference run live (only where the reference tree is present, i.e. the build
container): randomised small problems and flag combinations beyond the committed golden vectors.  This is the second
pin of oracle/smooth_nmf_oracle.py (the first being tests/golden/*.npz, which the same reference produced)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from conftest import rel_err
from oracle import ref_import
from oracle import smooth_nmf_oracle as orc

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from gen_golden import synth_problem  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present")

CONFIGS = [
    dict(simplex_H=True, simplex_W=False),
    dict(simplex_H=True, simplex_W=False, lambda_L=1.3, mu=0.04),
    dict(simplex_H=False, simplex_W=True, lambda_L=0.7),
    dict(simplex_H=False, simplex_W=False, mu=0.1),
    dict(simplex_H=True, simplex_W=False, normalize=True, mu=0.02),
    dict(simplex_H=True, simplex_W=False, lambda_L=0.9, algo="l2_surrogate"),
    dict(simplex_H=True, simplex_W=False, lambda_L=0.6, linesearch=True),
    dict(simplex_H=True, simplex_W=False, lambda_L=0.5, mu=0.03, algo="bmd", identity=True),
    dict(simplex_H=False, simplex_W=True, identity=True),
]


@pytest.mark.parametrize("seed", [11, 12, 13])
@pytest.mark.parametrize("ci", range(len(CONFIGS)))
def test_oracle_equals_live_reference(seed, ci):
    ref = ref_import.load_reference()
    kw = dict(CONFIGS[ci])
    identity = kw.pop("identity", False)
    rng = np.random.default_rng(1000 * seed + ci)
    nx, ny = int(rng.integers(4, 9)), int(rng.integers(4, 9))
    n, k, m = int(rng.integers(40, 90)), int(rng.integers(2, 5)), int(rng.integers(5, 9))
    pr = synth_problem(rng, n, nx, ny, k, m, counts=float(rng.uniform(10, 60)), identity_G=identity)
    X, G, W0, H0 = pr["X"], pr["G"], pr["W0"], pr["H0"]
    common = dict(tol=0, no_stop_criterion=True, max_iter=9, shape_2d=(nx, ny))
    est = ref.estimators.SmoothNMF(n_components=k, G=G, verbose=0, **common, **kw)
    with contextlib.redirect_stdout(io.StringIO()):
        est.fit_transform(X.copy(), W=W0.copy(), H=H0.copy())
    res = orc.fit(X, G, W0, H0, **common, **kw)
    assert res["n_iter"] == est.n_iter_
    assert rel_err(res["losses"], np.array(est.losses_)) < 1e-10
    assert rel_err(res["W"], est.W_) < 1e-9
    assert rel_err(res["H"], est.H_) < 1e-9
    assert rel_err(res["rel"], np.array(est.rel_)) < 1e-7


@pytest.mark.parametrize("case", ["both_missing", "only_H_given", "only_W_given", "both_given", "identity_both_missing"])
@pytest.mark.parametrize("simplex", ["H", "W", "none"])
def test_initialize_factors_equals_reference(case, simplex):
    """espm_b200.host.initialize_factors (row a11) against initialize_algorithms (updates.py:160-223), same random
    state: NNDSVD through scikit-learn, least squares for one missing factor, simplex rescale, clamp."""
    from espm_b200.host import initialize_factors
    ref = ref_import.load_reference()
    rng = np.random.default_rng(77)
    identity = case.startswith("identity")
    pr = synth_problem(rng, 60, 6, 7, 3, 7, identity_G=identity)
    X, G = pr["X"], pr["G"]
    X = np.maximum(X, 1e-14)
    W = pr["W0"].copy() if case in ("only_W_given", "both_given") else None
    H = pr["H0"].copy() if case in ("only_H_given", "both_given") else None
    sH, sW = simplex == "H", simplex == "W"
    Gr, Wr, Hr = ref.updates.initialize_algorithms(X, G, None if W is None else W.copy(), None if H is None else H.copy(),
                                                   3, None, 5, sH, sW)
    Go, Wo, Ho = initialize_factors(X, G, W, H, 3, None, 5, sH, sW, 1e-14)
    assert np.array_equal(Go, Gr)
    assert rel_err(Wo, Wr) < 1e-12 and rel_err(Ho, Hr) < 1e-12


def test_host_helpers_equal_reference():
    """rescaled_DH (utils.py:79-96, row a15), normalization_factor (base.py:16-18), remove_zeros_lines
    (base.py:519-528) and the truth-tracking measures find_min_angle / find_min_MSE (measures.py:125-153, 258-285,
    row a16) of espm_b200.host against the reference's functions."""
    from espm_b200 import host
    ref = ref_import.load_reference()
    rng = np.random.default_rng(5)
    D = rng.uniform(0.1, 1.0, size=(40, 3))
    H = rng.uniform(0.05, 1.0, size=(3, 90))
    Dr, Hr = ref.utils.rescaled_DH(D.copy(), H.copy())
    Do, Ho = host.rescaled_DH(D.copy(), H.copy())
    assert rel_err(Do, Dr) < 1e-12 and rel_err(Ho, Hr) < 1e-12
    X = rng.poisson(0.3, size=(40, 90)).astype(float)
    X[7, :] = 0
    X[:, 11] = 0
    assert host.normalization_factor(X, 3) == ref.estimators.base.normalization_factor(X, 3)
    est = ref.estimators.SmoothNMF(n_components=3, verbose=0)
    assert np.array_equal(host.remove_zeros_lines(X, 1e-14), est.remove_zeros_lines(X, 1e-14))
    with pytest.raises(ValueError):
        host.remove_zeros_lines(-np.ones((3, 3)), 1e-14)
    tv, av = rng.uniform(size=(3, 40)), rng.uniform(size=(3, 40))
    assert np.allclose(host.find_min_angle(tv, av), ref.measures.find_min_angle(tv, av, unique=True), rtol=1e-12)
    tm, am = rng.uniform(size=(3, 90)), rng.uniform(size=(3, 90))
    assert np.allclose(host.find_min_MSE(tm, am), ref.measures.find_min_MSE(tm, am, unique=True), rtol=1e-12)
