"""GPU parity of the alternative update rules: Frobenius (l2) and Bregman
(bmd) branches of the multiplicative steps, projected gradient, gradients, Lipschitz bounds, the line search on the
Laplacian surrogate -- against golden vectors produced by the unmodified reference (tests/golden/variants_small.npz,
oracle/gen_golden_variants.py) and against the oracle on seeded inputs."""
import numpy as np
import pytest

from conftest import rel_err
from test_oracle_vs_golden import VARIANT_FITS, variant_inputs

pytestmark = pytest.mark.gpu

STEP_TOL_F64 = 1e-10
STEP_TOL_F32 = 1e-5
TRAJ_TOL = 1e-4


@pytest.fixture(scope="module")
def ops():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from espm_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def orc():
    from oracle import smooth_nmf_oracle
    return smooth_nmf_oracle


def test_frobenius_steps_golden(ops, golden_variants):
    g = golden_variants
    X, G, W0, H0 = g["S__X"], g["S__G"], g["S__W0"], g["S__H0"]
    assert rel_err(ops.multiplicative_step_h(X, G, W0, H0, simplex_H=False, l2=True), g["h_l2"]) < STEP_TOL_F64
    assert rel_err(ops.multiplicative_step_h(X, G, W0, H0, simplex_H=True, l2=True), g["h_l2_simplex"]) < STEP_TOL_F64
    assert rel_err(ops.multiplicative_step_w(X, G, W0, H0, simplex_W=False, l2=True), g["w_l2"]) < STEP_TOL_F64
    # reference test_updates.py:457: with log_shift=0 the l2 W step is a fixed point at the truth-free toy problem
    ref = np.sum((G @ W0 @ H0 - X) ** 2)
    assert abs(ops.Frobenius_loss(X, G @ W0, H0) - ref) < 1e-10 * ref
    # known answer of the docstring (measures.py:366-372)
    Xd = np.array([[1, 1, -1], [2, 4, 5]], dtype=float)
    assert abs(ops.Frobenius_loss(Xd, np.array([[1.0], [1.0]]), np.array([[1.0, 2.0, 3.0]])) - 26) < 1e-12


def test_bregman_steps_golden(ops, golden_variants):
    g = golden_variants
    X, G, W0, H0 = g["S__X"], g["S__G"], g["S__W0"], g["S__H0"]
    sh = tuple(int(v) for v in g["S__shape"])
    Lg = ops.create_laplacian_matrix(*sh)
    assert rel_err(ops.multiplicative_step_h(X, G, W0, H0, simplex_H=False, use_bregman=True), g["h_bmd"]) < STEP_TOL_F64
    out = ops.multiplicative_step_h(X, G, W0, H0, simplex_H=True, mu=g["S__mu_vec"], lambda_L=1.5, L=Lg,
                                    use_bregman=True)
    assert rel_err(out, g["h_bmd_simplex_mu_lap"]) < STEP_TOL_F64
    out = ops.multiplicative_step_w(g["I__X"], np.eye(40), g["I__W0"], g["I__H0"], use_bregman=True)
    assert rel_err(out, g["w_bmd_identity"]) < STEP_TOL_F64
    out = ops.multiplicative_step_w(g["I__X"], g["I__Gsq"], g["I__W0sq"], g["I__H0"], use_bregman=True)
    assert rel_err(out, g["w_bmd_square"]) < STEP_TOL_F64
    with pytest.raises(ValueError):      # updates.py:42 cannot broadcast a non-square G against eye(n)
        ops.multiplicative_step_w(X, G, W0, H0, use_bregman=True)


def test_projected_gradient_golden(ops, golden_variants):
    g = golden_variants
    X, G, W0, H0 = g["S__X"], g["S__G"], g["S__W0"], g["S__H0"]
    sh = tuple(int(v) for v in g["S__shape"])
    Lg = ops.create_laplacian_matrix(*sh)
    mu = g["S__mu_vec"]
    assert rel_err(ops.gradH(X, G, W0, H0, mu=mu, lambda_L=0.7, L=Lg, epsilon_reg=0.5), g["gradH"]) < STEP_TOL_F64
    assert np.max(np.abs(ops.gradH(X, G, W0, H0, l2=True) - g["gradH_l2"])) < 1e-9 * np.max(np.abs(g["gradH_l2"]))
    assert rel_err(ops.gradW(X, G, W0, H0), g["gradW"]) < STEP_TOL_F64 * 10
    assert np.max(np.abs(ops.gradW(X, G, W0, H0, l2=True) - g["gradW_l2"])) < 1e-9 * np.max(np.abs(g["gradW_l2"]))
    out = ops.proj_grad_step_h(X, G, W0, H0, 60.0, simplex_H=True, mu=mu, lambda_L=0.7, L=Lg)
    assert rel_err(out, g["pg_h_simplex"]) < STEP_TOL_F64
    assert rel_err(ops.proj_grad_step_h(X, G, W0, H0, 60.0, simplex_H=False), g["pg_h_plain"]) < STEP_TOL_F64
    assert rel_err(ops.proj_grad_step_h(X, G, W0, H0, 400.0, simplex_H=True, l2=True), g["pg_h_l2"]) < 1e-9
    assert rel_err(ops.proj_grad_step_w(X, G, W0, H0, 3000.0, simplex_W=False), g["pg_w"]) < STEP_TOL_F64
    assert rel_err(ops.proj_grad_step_w(X, G, W0, H0, 2.0e4, simplex_W=False, l2=True), g["pg_w_l2"]) < 1e-9
    with pytest.raises(NotImplementedError):
        ops.proj_grad_step_w(X, G, W0, H0, 3000.0, simplex_W=True)
    lip_h = ops.estimate_Lipschitz_bound_h(1e-14, X, G, 3, lambda_L=0.7, mu=0.1, epsilon_reg=0.5)
    assert rel_err(lip_h, g["lip_h"]) < 1e-12
    assert rel_err(ops.estimate_Lipschitz_bound_w(1e-14, X, G, 3), g["lip_w"]) < 1e-12
    nu, its = ops.dichotomy_simplex_projected_gradient(g["pgd_a"], log_shift=1e-14, tol=1e-6, return_its=True)
    assert rel_err(nu, g["pgd_nu"]) < 1e-13
    f = np.sum(np.maximum(g["pgd_a"] + nu, 1e-14), axis=0) - 1
    assert np.max(np.abs(f)) <= 1e-6


@pytest.mark.parametrize("tag", sorted(t for t in VARIANT_FITS if not t.startswith("truth")))
@pytest.mark.parametrize("verbose", [0, 1])
def test_variant_fit_trajectory_golden(golden_variants, tag, verbose):
    from espm_b200 import SmoothNMF
    g = golden_variants
    X, G, W0, H0, sh = variant_inputs(g, tag)
    kw = dict(tol=0, no_stop_criterion=True, max_iter=10, shape_2d=sh, verbose=verbose)
    kw.update(VARIANT_FITS[tag])
    est = SmoothNMF(n_components=W0.shape[1], G=G, **kw)
    out = est.fit_transform(X, W=W0.copy(), H=H0.copy())
    assert est.n_iter_ == int(g[tag + "__n_iter"])
    assert rel_err(est.losses_, g[tag + "__losses"]) < TRAJ_TOL
    assert rel_err(est.losses_, g[tag + "__losses"]) < 1e-8       # what we actually get in fp64
    assert rel_err(est.W_, g[tag + "__W"]) < 1e-7
    assert rel_err(est.H_, g[tag + "__H"]) < 1e-7
    assert rel_err(out, g[tag + "__out"]) < 1e-7
    assert rel_err(est.reconstruction_err_, g[tag + "__rec"]) < 1e-8
    det = np.array(est.detailed_losses_, dtype=float)
    np.testing.assert_allclose(det, g[tag + "__detailed"], rtol=1e-7, atol=1e-16)     # incl. the gamma column


def test_variant_fits_fp32_vs_oracle(orc, golden_variants):
    """fp32 storage + arithmetic of the variants against the fp64 oracle on the fp32-rounded inputs."""
    from espm_b200 import SmoothNMF
    g = golden_variants
    for tag in ("l2", "bmd", "pg", "ls_log"):
        X, G, W0, H0, sh = variant_inputs(g, tag)
        X32, W32, H32 = X.astype(np.float32), W0.astype(np.float32), H0.astype(np.float32)
        G32 = None if G is None else G.astype(np.float32)
        kw = dict(tol=0, no_stop_criterion=True, max_iter=10, shape_2d=sh)
        kw.update(VARIANT_FITS[tag])
        ref = orc.fit(X32.astype(np.float64), None if G is None else G32.astype(np.float64), W32.astype(np.float64),
                      H32.astype(np.float64), **kw)
        est = SmoothNMF(n_components=W0.shape[1], G=G32, verbose=0, **kw)
        est.fit_transform(X32, W=W32.copy(), H=H32.copy())
        assert est.W_.dtype == np.float32
        assert rel_err(est.losses_, ref["losses"]) < TRAJ_TOL, tag
        if tag == "pg":   # additive update: small entries of H carry the ABSOLUTE fp32 error of H - grad / gamma
            assert np.max(np.abs(est.H_ - ref["H"])) < TRAJ_TOL * np.max(ref["H"]), tag
        else:
            assert rel_err(est.H_, ref["H"]) < 50 * TRAJ_TOL, tag


@pytest.mark.parametrize("tag", ["truth", "truth_free"])
def test_ground_truth_tracking(golden_variants, tag):
    """true_D / true_H: per-iteration angles, MSE and the loss on true_D @ true_H (base.py:301-347)."""
    from espm_b200 import SmoothNMF
    g = golden_variants
    X, G, W0, H0, sh = variant_inputs(g, tag)
    kw = dict(tol=0, no_stop_criterion=True, max_iter=10, shape_2d=sh, verbose=0)
    kw.update(VARIANT_FITS[tag])
    kw.pop("track")
    est = SmoothNMF(n_components=3, G=G, true_D=g["S__true_D"], true_H=g["S__true_H"], **kw)
    est.fit_transform(X, W=W0.copy(), H=H0.copy())
    assert rel_err(est.losses_, g[tag + "__losses"]) < 1e-9
    assert rel_err(est.true_losses_, g[tag + "__true_losses"]) < 1e-8
    assert rel_err(est.H_, g[tag + "__H"]) < 1e-8
    np.testing.assert_allclose(np.array(est.angles_), g[tag + "__angles"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(np.array(est.mse_), g[tag + "__mse"], rtol=1e-7, atol=1e-14)
    # get_losses() exposes them with the reference's column names (base.py:479-498)
    rec = est.get_losses()
    assert rec.dtype.names[-7:] == ("ang_p0", "ang_p1", "ang_p2", "mse_p0", "mse_p1", "mse_p2", "true_KL_loss")
    np.testing.assert_allclose(rec["ang_p1"], g[tag + "__angles"][:, 1], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(rec["mse_p2"], g[tag + "__mse"][:, 2], rtol=1e-7, atol=1e-14)
    assert rel_err(rec["true_KL_loss"], g[tag + "__true_losses"]) < 1e-8
    assert rel_err(rec["full_loss"], g[tag + "__losses"]) < 1e-9
    # a truth with another number of components is ignored with the reference's message (base.py:307-308)
    est2 = SmoothNMF(n_components=3, G=G, true_D=g["S__true_D"][:, :2], true_H=g["S__true_H"][:2], **kw)
    est2.fit_transform(X, W=W0.copy(), H=H0.copy())
    assert not hasattr(est2, "true_losses_")
    assert rel_err(est2.losses_, g[tag + "__losses"]) < 1e-9


def test_variant_guards():
    from espm_b200 import SmoothNMF
    X = np.random.default_rng(0).poisson(3.0, size=(20, 12)).astype(float)
    with pytest.raises(NotImplementedError):      # updates.py:365-366
        SmoothNMF(n_components=2, max_iter=2, verbose=0, algo="projected_gradient", simplex_W=True).fit_transform(X)
    with pytest.raises(ValueError):               # bmd needs a square G (updates.py:42)
        SmoothNMF(n_components=2, max_iter=2, verbose=0, algo="bmd", G=np.ones((20, 3)), simplex_W=False).fit_transform(X)
