"""Pin the oracle (oracle/smooth_nmf_oracle.py) against fixtures produced by the unmodified reference
(oracle/gen_golden.py) -- CPU only."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import smooth_nmf_oracle as orc

TOL = 1e-12


def test_laplacian_matches_reference_matrix(golden_steps):
    g = golden_steps
    nx, ny = int(g["nx"]), int(g["ny"])
    assert np.array_equal(orc.laplacian_dense(nx, ny), g["L_dense"])
    np.testing.assert_allclose(orc.laplacian_apply(g["H0"], (nx, ny)), g["HL"], rtol=0, atol=1e-13)
    assert rel_err(orc.trace_xtLx(g["H0"], (nx, ny)), g["trace_xtLx"]) < TOL


@pytest.mark.parametrize("nx,ny", [(2, 2), (2, 7), (3, 4), (4, 2)])
def test_laplacian_stencil_equals_dense(nx, ny):
    rng = np.random.default_rng(nx * 10 + ny)
    H = rng.uniform(size=(3, nx * ny))
    np.testing.assert_allclose(orc.laplacian_apply(H, (nx, ny)), H @ orc.laplacian_dense(nx, ny).astype(float),
                               rtol=0, atol=1e-13)
    # known answers of the reference's test_measures.py:212-221
    assert orc.trace_xtLx(np.ones((1, nx * ny)), (nx, ny)) == 0


def test_losses(golden_steps):
    g = golden_steps
    GW = g["G"] @ g["W0"]
    assert rel_err(orc.KLdiv_loss(g["X"], GW, g["H0"]), g["kl_loss"]) < TOL
    assert rel_err(orc.KLdiv_loss(g["X"], GW, g["H0"], average=True), g["kl_loss_avg"]) < TOL
    assert rel_err(orc.log_reg(g["H0"], g["mu_vec"], 1.0), g["log_reg"]) < TOL
    assert rel_err(orc.log_reg(g["H0"], 0.3, 0.5), g["log_reg_scalar"]) < TOL


def test_step_h_variants(golden_steps):
    g = golden_steps
    X, G, W, H = g["X"], g["G"], g["W0"], g["H0"]
    sh = (int(g["nx"]), int(g["ny"]))
    assert rel_err(orc.multiplicative_step_h(X, G, W, H, simplex_H=True), g["h_simplex"]) < TOL
    assert rel_err(orc.multiplicative_step_h(X, G, W, H, simplex_H=False), g["h_plain"]) < TOL
    assert rel_err(orc.multiplicative_step_h(X, G, W, H, simplex_H=False, log_shift=0), g["h_plain_ls0"]) < TOL
    out = orc.multiplicative_step_h(X, G, W, H, simplex_H=True, mu=g["mu_vec"], lambda_L=2.0, shape_2d=sh)
    assert rel_err(out, g["h_simplex_mu_lap"]) < TOL
    out = orc.multiplicative_step_h(X, G, W, H, simplex_H=False, mu=0.07, lambda_L=0.5, shape_2d=sh,
                                    epsilon_reg=0.5, sigmaL=6.0)
    assert rel_err(out, g["h_plain_mu_scalar_lap"]) < TOL
    out = orc.multiplicative_step_h(X, G, W, H, simplex_H=True, fixed_H=g["fixed_H"])
    assert rel_err(out, g["h_simplex_fixed"]) < TOL
    out = orc.multiplicative_step_h(X, G, W, H, simplex_H=True, dicotomy_tol=1e-8)
    assert rel_err(out, g["h_simplex_tol1e-8"]) < TOL
    out = orc.multiplicative_step_h(X, G, W, H, simplex_H=True, lambda_L=2.0, shape_2d=None)
    assert rel_err(out, g["h_simplex_lap_identity"]) < TOL


def test_step_hq_variants(golden_steps):
    g = golden_steps
    X, G, W, H = g["X"], g["G"], g["W0"], g["H0"]
    sh = (int(g["nx"]), int(g["ny"]))
    assert rel_err(orc.multiplicative_step_hq(X, G, W, H, simplex_H=True, lambda_L=1.5, shape_2d=sh),
                   g["hq_simplex_lap"]) < TOL
    assert rel_err(orc.multiplicative_step_hq(X, G, W, H, simplex_H=False), g["hq_plain"]) < TOL
    assert rel_err(orc.multiplicative_step_hq(X, G, W, H, simplex_H=True), g["hq_simplex"]) < TOL


def test_step_w_variants(golden_steps):
    g = golden_steps
    X, G, W, H1 = g["X"], g["G"], g["W0"], g["h_simplex"]
    assert rel_err(orc.multiplicative_step_w(X, G, W, H1, simplex_W=False), g["w_plain"]) < TOL
    assert rel_err(orc.multiplicative_step_w(X, G, W, H1, simplex_W=True), g["w_simplex"]) < TOL
    assert rel_err(orc.multiplicative_step_w(X, G, W, H1, simplex_W=False, fixed_W=g["fixed_W"]),
                   g["w_fixed"]) < TOL
    out = orc.multiplicative_step_w(X, G, W, H1, simplex_W=True, simplex_rows=g["simplex_rows"])
    assert rel_err(out, g["w_simplex_rows"]) < TOL


def test_identity_G_steps(golden_identity):
    g = golden_identity
    X, W, H = g["X"], g["W0"], g["H0"]
    Gid = np.eye(X.shape[0])
    h = orc.multiplicative_step_h(X, Gid, W, H, simplex_H=False)
    assert rel_err(h, g["h_plain"]) < TOL
    assert rel_err(orc.multiplicative_step_w(X, Gid, W, h, simplex_W=True), g["w_simplex"]) < TOL


def test_bisection_known_answers(golden_bisect):
    g = golden_bisect
    nu = orc.dichotomy_simplex(g["num"], g["den"], 1e-14, 1e-5)
    assert np.array_equal(nu, g["nu"])          # same bracket, same lock-step count -> bit-exact
    assert np.array_equal(orc.dichotomy_simplex(g["num"], g["den"], 1e-14, 1e-9), g["nu_tol1e-9"])
    assert np.array_equal(orc.dichotomy_simplex(g["num"], g["den"], 0, 1e-6), g["nu_ls0"])
    f = np.sum(np.maximum(g["num"] / (g["den"] + nu), 1e-14), axis=0) - 1
    assert np.max(np.abs(f)) <= 1e-5            # root property (reference test_updates.py:93-249)
    acc = orc.dichotomy_simplex_acc(3.0, g["acc_b"], g["acc_mc"], 1e-14, 1e-5)
    assert np.array_equal(acc, g["acc_nu"])
    with pytest.raises(ValueError):
        orc.dichotomy_simplex(np.ones((4, 2)), np.ones((4, 2)), log_shift=0.3)


FIT_CASES = {
    "c1": dict(simplex_H=True, simplex_W=False),
    "c2": dict(simplex_H=True, simplex_W=False, lambda_L=2.0, mu=0.05),
    "c2b": dict(simplex_H=True, simplex_W=False, lambda_L=1.0, mu=np.array([0.0, 0.1, 0.3]), shape_2d=None),
    "sw": dict(simplex_H=False, simplex_W=True, lambda_L=0.5),
    "none": dict(simplex_H=False, simplex_W=False),
    "norm": dict(simplex_H=True, simplex_W=False, normalize=True, mu=0.02),
    "stop": dict(simplex_H=True, simplex_W=False, tol=2e-3, max_iter=200, no_stop_criterion=False),
    "hq": dict(simplex_H=True, simplex_W=False, lambda_L=1.0, algo="l2_surrogate"),
}


@pytest.mark.parametrize("tag", sorted(FIT_CASES))
def test_fit_trajectories(golden_fits, tag):
    g = golden_fits
    kw = dict(tol=0, no_stop_criterion=True, max_iter=12, shape_2d=tuple(int(v) for v in g["A__shape"]))
    kw.update(FIT_CASES[tag])
    res = orc.fit(g["A__X"], g["A__G"], g["A__W0"], g["A__H0"], **kw)
    assert res["n_iter"] == int(g[tag + "__n_iter"])
    assert rel_err(res["losses"], g[tag + "__losses"]) < 1e-11
    assert rel_err(res["rel"], g[tag + "__rel"]) < 1e-9
    assert rel_err(res["detailed_losses"][:, :3], g[tag + "__detailed"][:, :3] + 0.0) < 1e-9 or \
        np.allclose(res["detailed_losses"][:, :3], g[tag + "__detailed"][:, :3], rtol=1e-9, atol=1e-18)
    assert rel_err(res["W"], g[tag + "__W"]) < 1e-10
    assert rel_err(res["H"], g[tag + "__H"]) < 1e-10
    assert rel_err(res["reconstruction_err"], g[tag + "__rec"]) < 1e-11


def test_fit_fixed_and_identity_and_model(golden_fits):
    g = golden_fits
    sh = tuple(int(v) for v in g["A__shape"])
    common = dict(tol=0, no_stop_criterion=True, max_iter=12)
    res = orc.fit(g["A__X"], g["A__G"], g["A__W0"], g["A__H0"], simplex_H=True, simplex_W=False,
                  shape_2d=sh, fixed_H=g["A__fixed_H"], fixed_W=g["A__fixed_W"], **common)
    assert rel_err(res["losses"], g["fixed__losses"]) < 1e-11
    assert rel_err(res["H"], g["fixed__H"]) < 1e-10
    # G=None
    res = orc.fit(g["I__X"], None, g["I__W0"], g["I__H0"], simplex_H=False, simplex_W=True,
                  shape_2d=(6, 7), **common)
    assert rel_err(res["losses"], g["c5__losses"]) < 1e-11
    assert rel_err(res["W"], g["c5__W"]) < 1e-10

    # synthetic PhysicalModel (same deterministic refresh rule as oracle/gen_golden.py)
    G0 = g["A__G"].copy()
    state = {"G": G0.copy()}
    rows = list(range(G0.shape[1] - 2))

    def g_update(W):
        s = np.mean(W[rows, :])
        newG = state["G"].copy()
        newG[:, -2] = G0[:, -2] * (1.0 + 0.3 * np.tanh(5 * s))
        newG[:, -1] = G0[:, -1] * (1.0 - 0.2 * np.tanh(3 * s))
        state["G"] = newG
        return newG

    res = orc.fit(g["A__X"], G0, g["A__W0"], g["A__H0"], simplex_H=False, simplex_W=True, shape_2d=sh,
                  lambda_L=0.3, g_update=g_update, simplex_rows=rows, **common)
    assert rel_err(res["losses"], g["pm__losses"]) < 1e-11
    assert rel_err(res["W"], g["pm__W"]) < 1e-10
    assert rel_err(res["G"], g["pm__G"]) < 1e-12


# ---------------------------------------------------------------------------------- alternative update rules
VARIANT_FITS = {
    "l2": dict(simplex_H=True, simplex_W=False, lambda_L=0.8, algo="l2_surrogate", l2=True),
    "l2w": dict(simplex_H=False, simplex_W=False, lambda_L=0.0, algo="l2_surrogate", l2=True),
    "bmd": dict(simplex_H=True, simplex_W=False, lambda_L=0.5, mu=0.03, algo="bmd"),
    "pg": dict(simplex_H=True, simplex_W=False, lambda_L=0.4, mu=0.02, algo="projected_gradient", gamma=[80.0, 4000.0]),
    "pgdef": dict(simplex_H=False, simplex_W=False, algo="projected_gradient"),
    "ls_log": dict(simplex_H=True, simplex_W=False, lambda_L=1.0, mu=0.02, linesearch=True),
    "ls_hq": dict(simplex_H=True, simplex_W=False, lambda_L=1.0, algo="l2_surrogate", linesearch=True),
    "ls_bmd": dict(simplex_H=True, simplex_W=False, lambda_L=0.5, algo="bmd", linesearch=True),
    "ls_pg": dict(simplex_H=True, simplex_W=False, lambda_L=0.4, algo="projected_gradient", gamma=[80.0, 4000.0],
                  linesearch=True),
    "truth": dict(simplex_H=True, simplex_W=False, lambda_L=0.3, track=True),
    "truth_free": dict(simplex_H=False, simplex_W=False, track=True),
}


def variant_inputs(g, tag):
    """(X, G, W0, H0, shape_2d) of a variants_small.npz fit: the bmd fits run on the identity-G problem."""
    if "bmd" in tag:
        return g["I__X"], None, g["I__W0"], g["I__H0"], (5, 6)
    return g["S__X"], g["S__G"], g["S__W0"], g["S__H0"], tuple(int(v) for v in g["S__shape"])


def test_variant_steps(golden_variants):
    g = golden_variants
    X, G, W0, H0 = g["S__X"], g["S__G"], g["S__W0"], g["S__H0"]
    sh = tuple(int(v) for v in g["S__shape"])
    mu = g["S__mu_vec"]
    assert rel_err(orc.multiplicative_step_h(X, G, W0, H0, simplex_H=False, l2=True), g["h_l2"]) < 1e-12
    assert rel_err(orc.multiplicative_step_h(X, G, W0, H0, simplex_H=True, l2=True), g["h_l2_simplex"]) < 1e-12
    assert rel_err(orc.multiplicative_step_w(X, G, W0, H0, simplex_W=False, l2=True), g["w_l2"]) < 1e-12
    assert rel_err(orc.multiplicative_step_h(X, G, W0, H0, simplex_H=False, use_bregman=True), g["h_bmd"]) < 1e-12
    out = orc.multiplicative_step_h(X, G, W0, H0, simplex_H=True, mu=mu, lambda_L=1.5, shape_2d=sh, use_bregman=True)
    assert rel_err(out, g["h_bmd_simplex_mu_lap"]) < 1e-12
    out = orc.multiplicative_step_w(g["I__X"], np.eye(40), g["I__W0"], g["I__H0"], use_bregman=True)
    assert rel_err(out, g["w_bmd_identity"]) < 1e-12
    out = orc.multiplicative_step_w(g["I__X"], g["I__Gsq"], g["I__W0sq"], g["I__H0"], use_bregman=True)
    assert rel_err(out, g["w_bmd_square"]) < 1e-12
    assert rel_err(orc.gradH(X, G, W0, H0, mu=mu, lambda_L=0.7, shape_2d=sh, epsilon_reg=0.5), g["gradH"]) < 1e-11
    assert rel_err(orc.gradH(X, G, W0, H0, l2=True), g["gradH_l2"]) < 1e-11
    assert rel_err(orc.gradW(X, G, W0, H0), g["gradW"]) < 1e-11
    assert rel_err(orc.gradW(X, G, W0, H0, l2=True), g["gradW_l2"]) < 1e-11
    out = orc.proj_grad_step_h(X, G, W0, H0, 60.0, simplex_H=True, mu=mu, lambda_L=0.7, shape_2d=sh)
    assert rel_err(out, g["pg_h_simplex"]) < 1e-11
    assert rel_err(orc.proj_grad_step_h(X, G, W0, H0, 60.0, simplex_H=False), g["pg_h_plain"]) < 1e-11
    assert rel_err(orc.proj_grad_step_h(X, G, W0, H0, 400.0, simplex_H=True, l2=True), g["pg_h_l2"]) < 1e-11
    assert rel_err(orc.proj_grad_step_w(X, G, W0, H0, 3000.0, simplex_W=False), g["pg_w"]) < 1e-11
    assert rel_err(orc.proj_grad_step_w(X, G, W0, H0, 2.0e4, simplex_W=False, l2=True), g["pg_w_l2"]) < 1e-11
    with pytest.raises(NotImplementedError):
        orc.proj_grad_step_w(X, G, W0, H0, 3000.0, simplex_W=True)
    assert rel_err(orc.estimate_Lipschitz_bound_h(1e-14, X, G, 3, lambda_L=0.7, mu=0.1, epsilon_reg=0.5), g["lip_h"]) < 1e-12
    assert rel_err(orc.estimate_Lipschitz_bound_w(1e-14, X, G, 3), g["lip_w"]) < 1e-12
    nu = orc.dichotomy_simplex_projected_gradient(g["pgd_a"].copy(), log_shift=1e-14, tol=1e-6)
    assert rel_err(nu, g["pgd_nu"]) < 1e-13
    assert rel_err(orc.diff_surrogate(H0, g["S__H1"], sh, sigmaL=8, algo="log_surrogate"), g["diff_log"]) < 1e-10
    assert rel_err(orc.diff_surrogate(H0, g["S__H1"], sh, sigmaL=0.3, algo="l2_surrogate"), g["diff_l2"]) < 1e-10


@pytest.mark.parametrize("tag", sorted(VARIANT_FITS))
def test_variant_fit_trajectories(golden_variants, tag):
    g = golden_variants
    X, G, W0, H0, sh = variant_inputs(g, tag)
    kw = dict(tol=0, no_stop_criterion=True, max_iter=10, shape_2d=sh)
    kw.update(VARIANT_FITS[tag])
    if kw.pop("track", False):
        kw.update(true_D=g["S__true_D"], true_H=g["S__true_H"])
    res = orc.fit(X, G, W0, H0, **kw)
    assert res["n_iter"] == int(g[tag + "__n_iter"])
    assert rel_err(res["losses"], g[tag + "__losses"]) < 1e-10
    assert rel_err(res["W"], g[tag + "__W"]) < 1e-9
    assert rel_err(res["H"], g[tag + "__H"]) < 1e-9
    assert rel_err(res["reconstruction_err"], g[tag + "__rec"]) < 1e-10
    np.testing.assert_allclose(res["detailed_losses"], g[tag + "__detailed"], rtol=1e-9, atol=1e-18)   # incl. gamma
    if tag.startswith("truth"):
        assert rel_err(res["true_losses"], g[tag + "__true_losses"]) < 1e-10
