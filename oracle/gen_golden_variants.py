"""TEST INFRASTRUCTURE ONLY -- golden vectors of the alternative update rules
from the UNMODIFIED reference:  ``python oracle/gen_golden_variants.py``  ->  tests/golden/variants_small.npz

Frobenius (l2=True) steps, Bregman (algo="bmd") steps, projected-gradient steps + gradients + Lipschitz bounds,
the line search on the Laplacian surrogate and ground-truth tracking, plus 10-iteration fits of each.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.gen_golden import synth_problem  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ref = load_reference()
    upd, dic, utl, sur = ref.updates, ref.dicotomy, ref.utils, ref.surrogates
    v = {}
    rng = np.random.default_rng(20240911)
    nx, ny, n, k, m = 6, 8, 96, 3, 7
    pr = synth_problem(rng, n, nx, ny, k, m, counts=30.0)
    X, G, W0, H0 = pr["X"], pr["G"], pr["W0"], pr["H0"]
    L = utl.create_laplacian_matrix(nx, ny)
    for key in ("X", "G", "W0", "H0"):
        v["S__" + key] = pr[key]
    v["S__shape"] = np.array([nx, ny])
    mu_vec = np.array([0.0, 0.04, 0.3])
    v["S__mu_vec"] = mu_vec
    # ---- Frobenius branches (updates.py:29-36, 109-118)
    v["h_l2"] = upd.multiplicative_step_h(X, G, W0, H0, simplex_H=False, l2=True)
    v["h_l2_simplex"] = upd.multiplicative_step_h(X, G, W0, H0, simplex_H=True, l2=True)
    v["w_l2"] = upd.multiplicative_step_w(X, G, W0, H0, simplex_W=False, l2=True)
    # ---- Bregman branches (updates.py:40-48, 120-125): the W branch needs a square G (np.allclose vs eye)
    v["h_bmd"] = upd.multiplicative_step_h(X, G, W0, H0, simplex_H=False, use_bregman=True)
    v["h_bmd_simplex_mu_lap"] = upd.multiplicative_step_h(X, G, W0, H0, simplex_H=True, mu=mu_vec, lambda_L=1.5, L=L,
                                                          use_bregman=True)
    prI = synth_problem(np.random.default_rng(3), 40, 5, 6, 3, 0, counts=25.0, identity_G=True)
    for key in ("X", "W0", "H0"):
        v["I__" + key] = prI[key]
    Gid = np.eye(40)
    v["w_bmd_identity"] = upd.multiplicative_step_w(prI["X"], Gid, prI["W0"], prI["H0"], use_bregman=True)
    Gsq = np.abs(np.random.default_rng(4).normal(size=(40, 40))) * 0.1 + np.eye(40)
    v["I__Gsq"] = Gsq
    W0sq = np.abs(np.linalg.lstsq(Gsq, prI["W0"], rcond=None)[0]) + 0.01
    v["I__W0sq"] = W0sq
    v["w_bmd_square"] = upd.multiplicative_step_w(prI["X"], Gsq, W0sq, prI["H0"], use_bregman=True)
    # ---- gradients, projected-gradient steps, Lipschitz bounds (updates.py:303-413, dicotomy.py:83-108)
    v["gradH"] = upd.gradH(X, G, W0, H0, mu=mu_vec, lambda_L=0.7, L=L, epsilon_reg=0.5)
    v["gradH_l2"] = upd.gradH(X, G, W0, H0, l2=True)
    v["gradW"] = upd.gradW(X, G, W0, H0)
    v["gradW_l2"] = upd.gradW(X, G, W0, H0, l2=True)
    v["pg_h_simplex"] = upd.proj_grad_step_h(X, G, W0, H0, 60.0, simplex_H=True, mu=mu_vec, lambda_L=0.7, L=L)
    v["pg_h_plain"] = upd.proj_grad_step_h(X, G, W0, H0, 60.0, simplex_H=False)
    v["pg_h_l2"] = upd.proj_grad_step_h(X, G, W0, H0, 400.0, simplex_H=True, l2=True)
    v["pg_w"] = upd.proj_grad_step_w(X, G, W0, H0, 3000.0, simplex_W=False)
    v["pg_w_l2"] = upd.proj_grad_step_w(X, G, W0, H0, 2.0e4, simplex_W=False, l2=True)
    v["lip_h"] = upd.estimate_Lipschitz_bound_h(1e-14, X, G, k, lambda_L=0.7, mu=0.1, epsilon_reg=0.5)
    v["lip_w"] = upd.estimate_Lipschitz_bound_w(1e-14, X, G, k)
    a = np.random.default_rng(9).normal(size=(4, 50)) * 0.3
    v["pgd_a"] = a
    v["pgd_nu"] = dic.dichotomy_simplex_projected_gradient(a.copy(), log_shift=1e-14, tol=1e-6)
    # ---- surrogates (surrogates.py)
    H1 = upd.multiplicative_step_h(X, G, W0, H0, simplex_H=True, lambda_L=1.0, L=L)
    v["S__H1"] = H1
    v["diff_log"] = sur.diff_surrogate(H0, H1, L=L, sigmaL=8, algo="log_surrogate")
    v["diff_l2"] = sur.diff_surrogate(H0, H1, L=L, sigmaL=0.3, algo="l2_surrogate")

    # ---- fits
    def run_fit(tag, pr_, shape_2d, **kw):
        est = ref.SmoothNMF(n_components=pr_["W0"].shape[1], shape_2d=shape_2d, verbose=0, **kw)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            out = est.fit_transform(pr_["X"], W=pr_["W0"].copy(), H=pr_["H0"].copy())
        d = dict(out=out, W=est.W_, H=est.H_, losses=np.array(est.losses_),
                 detailed=np.array(est.detailed_losses_, dtype=float), rel=np.array(est.rel_), n_iter=est.n_iter_,
                 rec=est.reconstruction_err_)
        for extra in ("true_losses_", "angles_", "mse_"):
            if hasattr(est, extra):
                d[extra.rstrip("_")] = np.array(getattr(est, extra), dtype=float)
        for k_, val in d.items():
            v["%s__%s" % (tag, k_)] = val

    common = dict(tol=0, no_stop_criterion=True, max_iter=10)
    run_fit("l2", pr, (nx, ny), G=G, simplex_H=True, simplex_W=False, lambda_L=0.8, algo="l2_surrogate", l2=True,
            **common)
    run_fit("l2w", pr, (nx, ny), G=G, simplex_H=False, simplex_W=False, lambda_L=0.0, algo="l2_surrogate", l2=True,
            **common)
    run_fit("bmd", prI, (5, 6), G=None, simplex_H=True, simplex_W=False, lambda_L=0.5, mu=0.03, algo="bmd", **common)
    run_fit("pg", pr, (nx, ny), G=G, simplex_H=True, simplex_W=False, lambda_L=0.4, mu=0.02,
            algo="projected_gradient", gamma=[80.0, 4000.0], **common)
    run_fit("pgdef", pr, (nx, ny), G=G, simplex_H=False, simplex_W=False, algo="projected_gradient", **common)
    run_fit("ls_log", pr, (nx, ny), G=G, simplex_H=True, simplex_W=False, lambda_L=1.0, mu=0.02, linesearch=True,
            **common)
    run_fit("ls_hq", pr, (nx, ny), G=G, simplex_H=True, simplex_W=False, lambda_L=1.0, algo="l2_surrogate",
            linesearch=True, **common)
    run_fit("ls_bmd", prI, (5, 6), G=None, simplex_H=True, simplex_W=False, lambda_L=0.5, algo="bmd",
            linesearch=True, **common)
    run_fit("ls_pg", pr, (nx, ny), G=G, simplex_H=True, simplex_W=False, lambda_L=0.4,
            algo="projected_gradient", gamma=[80.0, 4000.0], linesearch=True, **common)
    Dt = G @ pr["W_true"]
    v["S__true_D"], v["S__true_H"] = Dt, pr["H_true"]
    run_fit("truth", pr, (nx, ny), G=G, simplex_H=True, simplex_W=False, lambda_L=0.3, true_D=Dt,
            true_H=pr["H_true"], **common)
    run_fit("truth_free", pr, (nx, ny), G=G, simplex_H=False, simplex_W=False, true_D=Dt, true_H=pr["H_true"],
            **common)
    np.savez_compressed(os.path.join(OUT, "variants_small.npz"), **v)
    print("written", os.path.join(OUT, "variants_small.npz"), os.path.getsize(os.path.join(OUT, "variants_small.npz")))
    for key in sorted(v):
        if key.endswith("__losses"):
            print(key, v[key][[0, -1]])


if __name__ == "__main__":
    main()
