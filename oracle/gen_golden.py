"""TEST INFRASTRUCTURE ONLY -- regenerate ``tests/golden/*.npz`` from the UNMODIFIED reference.

Run in the build container (where ``/root/reference`` exists):

    python oracle/gen_golden.py

Every fixture stores the seeded inputs and the outputs the reference's own functions produced
(espm/estimators/updates.py, dicotomy.py, measures.py, utils.py, base.py, smooth_nmf.py).  The GPU box
has no reference tree; tests there compare against these files and against the oracle restatement.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_import import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def synth_problem(rng, n, nx, ny, k, m, counts=30.0, identity_G=False, dtype=np.float64):
    """Small EDXS-like problem: smooth non-negative G columns, Poisson counts X (many zeros)."""
    p = nx * ny
    x = np.linspace(0, 1, n)
    if identity_G:
        G = None
        D = np.abs(rng.normal(size=(n, k))) + 0.05
        D /= D.sum(0, keepdims=True)
        W_true = D
    else:
        G = np.zeros((n, m))
        for j in range(m - 2):
            c, s = rng.uniform(0.05, 0.95), rng.uniform(0.01, 0.04)
            G[:, j] = np.exp(-0.5 * ((x - c) / s) ** 2)
        G[:, m - 2] = np.exp(-3 * x) + 0.05
        G[:, m - 1] = (1 - x) * 0.5 + 0.05
        W_true = rng.uniform(0, 1, size=(m, k)) * (rng.uniform(size=(m, k)) > 0.4)
        W_true[-2:, :] = rng.uniform(0.05, 0.2, size=(2, k))
        W_true /= W_true.sum(0, keepdims=True)
        D = G @ W_true
    H_true = rng.uniform(size=(k, p)) ** 2
    H_true /= H_true.sum(0, keepdims=True)
    lam = D @ H_true
    lam = lam / lam.sum(0, keepdims=True) * counts
    X = rng.poisson(lam).astype(dtype)
    mW = n if identity_G else m
    W0 = rng.uniform(0.05, 1.0, size=(mW, k))
    H0 = rng.uniform(0.05, 1.0, size=(k, p))
    H0 /= H0.sum(0, keepdims=True)
    return dict(X=X, G=G, W0=W0, H0=H0, W_true=W_true, H_true=H_true)


def main():
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    upd, dic, mea, utl = ref.updates, ref.dicotomy, ref.measures, ref.utils

    # ------------------------------------------------------------------ single steps
    rng = np.random.default_rng(20240501)
    nx, ny, n, k, m = 7, 9, 150, 3, 8
    pr = synth_problem(rng, n, nx, ny, k, m)
    X, G, W0, H0 = pr["X"], pr["G"], pr["W0"], pr["H0"]
    L = utl.create_laplacian_matrix(nx, ny)
    mu_vec = np.array([0.0, 0.05, 0.2])
    fixed_H = -np.ones_like(H0)
    fixed_H[1, 5:17] = 0.25
    fixed_W = -np.ones_like(W0)
    fixed_W[2, :] = 0.125
    steps = dict(X=X, G=G, W0=W0, H0=H0, nx=nx, ny=ny, mu_vec=mu_vec, fixed_H=fixed_H, fixed_W=fixed_W)
    steps["h_simplex"] = upd.multiplicative_step_h(X, G, W0, H0, simplex_H=True)
    steps["h_plain"] = upd.multiplicative_step_h(X, G, W0, H0, simplex_H=False)
    steps["h_plain_ls0"] = upd.multiplicative_step_h(X, G, W0, H0, simplex_H=False, log_shift=0)
    steps["h_simplex_mu_lap"] = upd.multiplicative_step_h(
        X, G, W0, H0, simplex_H=True, mu=mu_vec, lambda_L=2.0, L=L, epsilon_reg=1)
    steps["h_plain_mu_scalar_lap"] = upd.multiplicative_step_h(
        X, G, W0, H0, simplex_H=False, mu=0.07, lambda_L=0.5, L=L, epsilon_reg=0.5, sigmaL=6.0)
    steps["h_simplex_fixed"] = upd.multiplicative_step_h(X, G, W0, H0, simplex_H=True, fixed_H=fixed_H)
    steps["h_simplex_tol1e-8"] = upd.multiplicative_step_h(X, G, W0, H0, simplex_H=True, dicotomy_tol=1e-8)
    # identity "Laplacian" (shape_2d=None in the estimator, base.py:289-291)
    from scipy.sparse import lil_matrix
    Lid = lil_matrix((nx * ny, nx * ny), dtype=np.float32)
    Lid.setdiag([1] * (nx * ny))
    steps["h_simplex_lap_identity"] = upd.multiplicative_step_h(
        X, G, W0, H0, simplex_H=True, lambda_L=2.0, L=Lid)
    H1 = steps["h_simplex"]
    steps["w_plain"] = upd.multiplicative_step_w(X, G, W0, H1, simplex_W=False)
    steps["w_simplex"] = upd.multiplicative_step_w(X, G, W0, H1, simplex_W=True)
    steps["w_fixed"] = upd.multiplicative_step_w(X, G, W0, H1, simplex_W=False, fixed_W=fixed_W)

    class _Rows:
        def NMF_simplex(self):
            return [0, 1, 3, 4, 5]

    steps["simplex_rows"] = np.array(_Rows().NMF_simplex())
    steps["w_simplex_rows"] = upd.multiplicative_step_w(X, G, W0, H1, simplex_W=True, physics_model=_Rows())
    steps["hq_simplex_lap"] = upd.multiplicative_step_hq(X, G, W0, H0, simplex_H=True, lambda_L=1.5, L=L)
    steps["hq_plain"] = upd.multiplicative_step_hq(X, G, W0, H0, simplex_H=False)
    steps["hq_simplex"] = upd.multiplicative_step_hq(X, G, W0, H0, simplex_H=True)
    # losses
    GW = G @ W0
    steps["kl_loss"] = mea.KLdiv_loss(X, GW, H0)
    steps["kl_loss_avg"] = mea.KLdiv_loss(X, GW, H0, average=True)
    steps["log_reg"] = mea.log_reg(H0, mu_vec, 1.0)
    steps["log_reg_scalar"] = mea.log_reg(H0, 0.3, 0.5)
    steps["trace_xtLx"] = mea.trace_xtLx(L, H0.T)
    steps["HL"] = H0 @ L
    steps["L_dense"] = np.asarray(L.todense())
    np.savez_compressed(os.path.join(OUT, "steps_small.npz"), **steps)

    # ------------------------------------------------------------------ identity-G steps (G=None)
    rng = np.random.default_rng(77)
    pr = synth_problem(rng, 96, 6, 5, 4, 0, identity_G=True)
    X, W0, H0 = pr["X"], pr["W0"], pr["H0"]
    Gid = np.diag(np.ones(X.shape[0]))
    ident = dict(X=X, W0=W0, H0=H0)
    ident["h_plain"] = upd.multiplicative_step_h(X, Gid, W0, H0, simplex_H=False)
    ident["w_simplex"] = upd.multiplicative_step_w(X, Gid, W0, ident["h_plain"], simplex_W=True)
    np.savez_compressed(os.path.join(OUT, "steps_identity.npz"), **ident)

    # ------------------------------------------------------------------ bisection known answers
    rng = np.random.default_rng(5)
    num = rng.uniform(size=(5, 64)) * (rng.uniform(size=(5, 64)) > 0.3)
    num[0, :] += 0.01
    den = rng.uniform(size=(5, 64))
    bis = dict(num=num, den=den)
    bis["nu"] = dic.dichotomy_simplex(num.copy(), den.copy(), log_shift=1e-14, tol=1e-5)
    bis["nu_tol1e-9"] = dic.dichotomy_simplex(num.copy(), den.copy(), log_shift=1e-14, tol=1e-9)
    bis["nu_ls0"] = dic.dichotomy_simplex(num.copy(), den.copy(), log_shift=0, tol=1e-6)
    a, b, mc = 3.0, rng.uniform(size=(5, 64)), rng.uniform(size=(5, 64))
    bis["acc_b"], bis["acc_mc"] = b, mc
    bis["acc_nu"] = dic.dichotomy_simplex_acc(a, b.copy(), mc.copy(), log_shift=1e-14, tol=1e-5)
    np.savez_compressed(os.path.join(OUT, "bisect_small.npz"), **bis)

    # ------------------------------------------------------------------ fit trajectories
    def run_fit(tag, pr, shape_2d, **kw):
        w0 = pr["W0"].copy()
        h0 = pr["H0"].copy()
        est = ref.SmoothNMF(n_components=w0.shape[1], shape_2d=shape_2d, verbose=0, **kw)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            out = est.fit_transform(pr["X"], W=w0, H=h0)
        d = {tag + "__" + k_: v for k_, v in dict(
            out=out, W=est.W_, H=est.H_, G=est.G_, losses=np.array(est.losses_),
            detailed=np.array(est.detailed_losses_, dtype=float), rel=np.array(est.rel_),
            n_iter=est.n_iter_, rec=est.reconstruction_err_, components=est.components_,
            stdout=np.array(buf.getvalue())).items()}
        return d

    fits = {}
    rng = np.random.default_rng(91)
    nx, ny, n, k, m = 8, 10, 128, 3, 9
    pr = synth_problem(rng, n, nx, ny, k, m, counts=25.0)
    for key in ("X", "G", "W0", "H0"):
        fits["A__" + key] = pr[key]
    fits["A__shape"] = np.array([nx, ny])
    common = dict(tol=0, no_stop_criterion=True, max_iter=12)
    # C1-like: simplex_H, no regularisation
    fits.update(run_fit("c1", pr, (nx, ny), G=pr["G"], simplex_H=True, simplex_W=False, **common))
    # C2-like: + Laplacian + log-reg
    fits.update(run_fit("c2", pr, (nx, ny), G=pr["G"], simplex_H=True, simplex_W=False,
                        lambda_L=2.0, mu=0.05, **common))
    # vector mu, identity-L (shape_2d None with lambda_L>0)
    fits.update(run_fit("c2b", pr, None, G=pr["G"], simplex_H=True, simplex_W=False,
                        lambda_L=1.0, mu=np.array([0.0, 0.1, 0.3]), **common))
    # simplex_W (default flags) with ndarray G
    fits.update(run_fit("sw", pr, (nx, ny), G=pr["G"], simplex_H=False, simplex_W=True,
                        lambda_L=0.5, **common))
    # neither simplex -> rescaled_DH at the end
    fits.update(run_fit("none", pr, (nx, ny), G=pr["G"], simplex_H=False, simplex_W=False, **common))
    # normalize
    fits.update(run_fit("norm", pr, (nx, ny), G=pr["G"], simplex_H=True, simplex_W=False,
                        normalize=True, mu=0.02, **common))
    # stop criteria active
    fits.update(run_fit("stop", pr, (nx, ny), G=pr["G"], simplex_H=True, simplex_W=False,
                        tol=2e-3, max_iter=200))
    # l2_surrogate (quadratic surrogate H step)
    fits.update(run_fit("hq", pr, (nx, ny), G=pr["G"], simplex_H=True, simplex_W=False,
                        lambda_L=1.0, algo="l2_surrogate", **common))
    # hyperspy-compatible layout
    prT = dict(pr)
    prT["X"] = np.ascontiguousarray(pr["X"].T)
    fits.update(run_fit("hspy", prT, (nx, ny), G=pr["G"], simplex_H=True, simplex_W=False,
                        hspy_comp=True, lambda_L=1.0, **common))
    # fixed entries
    fixed_H = -np.ones_like(pr["H0"])
    fixed_H[0, :11] = 0.5
    fixed_W = -np.ones_like(pr["W0"])
    fixed_W[1, :] = 0.2
    fits["A__fixed_H"], fits["A__fixed_W"] = fixed_H, fixed_W
    fits.update(run_fit("fixed", pr, (nx, ny), G=pr["G"], simplex_H=True, simplex_W=False,
                        fixed_H=fixed_H, fixed_W=fixed_W, **common))

    # synthetic PhysicalModel: deterministic refresh of the last two columns of G every 3 its
    PM = ref.models_base.PhysicalModel

    class FakeModel(PM):
        def __init__(self, G):
            self.G = G.copy()
            self.G0 = G.copy()

        def generate_g_matr(self, *a, **k_):
            pass

        def generate_phases(self, *a, **k_):
            pass

        def NMF_initialize_W(self, D):
            return np.abs(np.linalg.lstsq(self.G, D, rcond=None)[0])

        def NMF_simplex(self):
            return list(range(self.G.shape[1] - 2))

        def NMF_update(self, W=None):
            if W is None:
                return self.G
            s = np.mean(W[self.NMF_simplex(), :])
            newG = self.G.copy()
            newG[:, -2] = self.G0[:, -2] * (1.0 + 0.3 * np.tanh(5 * s))
            newG[:, -1] = self.G0[:, -1] * (1.0 - 0.2 * np.tanh(3 * s))
            self.G = newG
            return self.G

    fits.update(run_fit("pm", pr, (nx, ny), G=FakeModel(pr["G"]), simplex_H=False, simplex_W=True,
                        lambda_L=0.3, **common))

    # G=None (identity), simplex_W, C5-like
    rng = np.random.default_rng(95)
    prI = synth_problem(rng, 80, 6, 7, 4, 0, counts=40.0, identity_G=True)
    for key in ("X", "W0", "H0"):
        fits["I__" + key] = prI[key]
    fits.update(run_fit("c5", prI, (6, 7), G=None, simplex_H=False, simplex_W=True, **common))
    # float32 end-to-end (X fp32, G None): reference computes in fp32
    prI32 = {k_: (v.astype(np.float32) if isinstance(v, np.ndarray) else v) for k_, v in prI.items()}
    fits.update(run_fit("c5f32", prI32, (6, 7), G=None, simplex_H=False, simplex_W=True, **common))
    np.savez_compressed(os.path.join(OUT, "fits_small.npz"), **fits)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  ", f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
