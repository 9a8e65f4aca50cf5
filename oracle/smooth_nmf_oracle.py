"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of espm's SmoothNMF fit loop.

This file is the *oracle* for the CUDA path in ``espm_b200``.  It must never be imported by the
product package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it.  It restates the algorithm of adriente/espm v1.1.3 (citations are
``file:line`` relative to the reference root) with plain NumPy:

* the Laplacian is a 5-point Neumann stencil instead of a scipy.sparse matrix (utils.py:39-76),
* the bisection bracket is vectorised instead of a Python loop over pixels (dicotomy.py:29-35),
* everything else follows the reference's operation order (same n x p temporaries, same BLAS
  contractions), so that it is also a fair CPU timing baseline of the reference's algorithm.

Parity pin: ``tests/test_oracle_vs_golden.py`` checks every function here against
``tests/golden/*.npz`` which were produced by importing the unmodified reference
(``oracle/gen_golden.py``), and -- when ``/root/reference`` is present -- against the live reference.
"""
import numpy as np

# espm/conf.py:55-59
LOG_SHIFT = 1e-14
DICOTOMY_TOL = 1e-5
SIGMA_L = 8
MAXIT_DICHOTOMY = 100


# --------------------------------------------------------------------------------------------
# Laplacian (utils.py:39-76; identity when shape_2d is None, base.py:289-291)
# --------------------------------------------------------------------------------------------
def laplacian_apply(H, shape_2d):
    """Return H @ L for the graph Laplacian of the nx x ny pixel grid (pixel index = i*ny + j).

    L = D - A with A the 4-neighbour adjacency and D the number of in-bounds neighbours
    (utils.py:61-75).  ``shape_2d is None`` means L = identity (base.py:289-291).
    """
    if shape_2d is None:
        return H.copy()
    nx, ny = shape_2d
    k = H.shape[0]
    Hi = H.reshape(k, nx, ny)
    out = np.zeros_like(Hi)
    # vertical neighbours (i-1, i+1)
    out[:, 1:, :] += Hi[:, 1:, :] - Hi[:, :-1, :]
    out[:, :-1, :] += Hi[:, :-1, :] - Hi[:, 1:, :]
    # horizontal neighbours (j-1, j+1)
    out[:, :, 1:] += Hi[:, :, 1:] - Hi[:, :, :-1]
    out[:, :, :-1] += Hi[:, :, :-1] - Hi[:, :, 1:]
    return out.reshape(k, nx * ny)


def laplacian_apply_scaled32(H, shape_2d, lam):
    """H @ (lam * L) the way the reference's gradH forms it (updates.py:343, ``lambda_L * L @ H.T``): the
    sparse L holds float32 entries (utils.py:57), so ``lambda_L * L`` is ROUNDED TO FLOAT32 before it meets
    the float64 H: entries fl32(deg * fl32(lam)) on the diagonal and -fl32(lam) off it."""
    lam32 = np.float32(lam)
    if shape_2d is None:
        return float(lam32) * H
    nx, ny = shape_2d
    k = H.shape[0]
    Hi = H.reshape(k, nx, ny)
    deg = np.zeros((nx, ny), dtype=np.float32)
    nb = np.zeros_like(Hi)
    deg[1:, :] += 1
    nb[:, 1:, :] += Hi[:, :-1, :]
    deg[:-1, :] += 1
    nb[:, :-1, :] += Hi[:, 1:, :]
    deg[:, 1:] += 1
    nb[:, :, 1:] += Hi[:, :, :-1]
    deg[:, :-1] += 1
    nb[:, :, :-1] += Hi[:, :, 1:]
    cdeg = (deg * lam32).astype(np.float64)            # float32 product, then exact widening
    return (cdeg[None] * Hi - float(lam32) * nb).reshape(k, nx * ny)


def laplacian_dense(nx, ny=None):
    """Dense restatement of create_laplacian_matrix (utils.py:39-76), for small tests."""
    if ny is None:
        ny = nx
    assert nx > 1 and ny > 1
    p = nx * ny
    L = np.zeros((p, p), dtype=np.float32)
    for i in range(nx):
        for j in range(ny):
            a = i * ny + j
            for di, dj in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                ii, jj = i + di, j + dj
                if 0 <= ii < nx and 0 <= jj < ny:
                    L[a, a] += 1
                    L[a, ii * ny + jj] = -1
    return L


def trace_xtLx(H, shape_2d):
    """sum(H * (H L)) (measures.py:560-577 called with x = H.T, smooth_nmf.py:466)."""
    return np.sum(H * laplacian_apply(H, shape_2d))


# --------------------------------------------------------------------------------------------
# Lock-step bisection (dicotomy.py)
# --------------------------------------------------------------------------------------------
def dicotomy(a, b, func, maxit, tol, return_its=False, force_its=None):
    """dicotomy.py:111-173.  Vectorised bisection with a GLOBAL stop test (line 152).

    ``force_its`` (not in the reference; test infrastructure) runs exactly that many updates instead of applying the
    stop test: the fp32 parity tests use it to evaluate "what the reference computes when its lock-step count is c"
    when the fp32 device data and the fp64 oracle disagree on a knife-edge stop decision."""
    func_max = func(a)
    func_min = func(b)
    assert np.sum(func_min >= 0) == 0
    assert np.sum(func_max <= 0) == 0
    assert np.sum(np.isnan(func_max)) == 0
    assert np.sum(np.isnan(func_min)) == 0
    it = 0
    new = (a + b) / 2
    func_new = func(new)
    while (np.max(np.abs(func_new)) > tol) if force_its is None else (it < force_its):
        it += 1
        func_a = func(a)
        minus_bool = func_a * func_new <= 0
        plus_bool = np.logical_not(minus_bool)
        b[minus_bool] = new[minus_bool]
        a[plus_bool] = new[plus_bool]
        new = (a + b) / 2
        func_new = func(new)
        if it >= maxit and force_its is None:
            break
    if return_its:
        return new, it
    return new


def dichotomy_simplex(num, denum, log_shift=LOG_SHIFT, tol=DICOTOMY_TOL, maxit=MAXIT_DICHOTOMY,
                      return_its=False, force_its=None):
    """dicotomy.py:4-55: root of sum_i max(num_i/(x+den_i), log_shift) - 1 per column."""
    assert (num >= 0).all()
    assert (denum >= 0).all()
    assert (np.sum(num, axis=0) > 0).all()
    if log_shift > 0:
        if denum.shape[0] * log_shift >= 1:
            raise ValueError("No solution exists!")
    # dicotomy.py:29-43: a = max over rows with num>0 of (num/2 - denum)
    cand = np.where(num > 0, num / 2 - denum, -np.inf)
    a = np.max(cand, axis=0)
    # dicotomy.py:49
    b = len(num) * np.max(num, axis=0) / 0.5 - np.min(denum, axis=0)

    def func(x):
        return np.sum(np.maximum(num / (x + denum), log_shift), axis=0) - 1

    return dicotomy(a, b, func, maxit, tol, return_its=return_its, force_its=force_its)


def dichotomy_simplex_acc(a, b, minus_c, log_shift=LOG_SHIFT, tol=DICOTOMY_TOL,
                          maxit=MAXIT_DICHOTOMY, return_its=False):
    """dicotomy.py:57-81 (quadratic-surrogate H step)."""
    assert a >= 0
    assert (minus_c >= 0).all()
    if log_shift > 0:
        if b.shape[0] * log_shift >= 1:
            raise ValueError("No solution exists!")
    n_p = len(b)
    nu_max = n_p * np.max(b ** 2 / a + 2 * a + 2 * (b + minus_c), axis=0) * 1.5 + 1e-3
    nu_min = -(2 * a + np.sum(b, axis=0)) / n_p * 1.1 - 1e-3

    def func(x):
        return 2 * a - np.sum(
            np.maximum(np.sqrt((b + x) ** 2 + 4 * a * minus_c) - x - b, log_shift * 2 * a), axis=0)

    return dicotomy(nu_max, nu_min, func, maxit, tol, return_its=return_its)


# --------------------------------------------------------------------------------------------
# Update steps (updates.py)
# --------------------------------------------------------------------------------------------
def _mu_column(mu):
    if np.isscalar(mu):
        return mu
    mu = np.asarray(mu)
    if mu.ndim == 1:
        mu = mu[:, None]
    return mu


def multiplicative_step_h(X, G, W, H, simplex_H=False, mu=0, log_shift=LOG_SHIFT, epsilon_reg=1,
                          safe=True, dicotomy_tol=DICOTOMY_TOL, lambda_L=0, shape_2d=None,
                          sigmaL=SIGMA_L, fixed_H=None, return_its=False, l2=False, use_bregman=False,
                          force_its=None):
    """updates.py:83-156 (``L`` is replaced by ``shape_2d``: stencil Laplacian).  ``force_its``: see ``dicotomy``."""
    if lambda_L != 0:
        HL = laplacian_apply(H, shape_2d)                      # updates.py:96
    if safe:
        assert np.sum(H < -log_shift / 2) == 0                  # updates.py:101-105
        assert np.sum(W < -log_shift / 2) == 0
        assert np.sum(G < -log_shift / 2) == 0
        H = np.maximum(H, log_shift)
        W = np.maximum(W, log_shift)
    GW = G @ W                                                 # updates.py:107
    if l2:                                                     # updates.py:109-118
        assert lambda_L == 0
        assert (mu == 0) if np.isscalar(mu) else (np.asarray(mu) == 0).all()
        num = GW.T @ X
        denum = (GW.T @ GW) @ H
    else:
        GWH = GW @ H                                           # updates.py:121/127
        if use_bregman:                                        # updates.py:120-125
            sigmaR = np.sum(X, axis=0, keepdims=True)
            num = sigmaR / H
            gradg = -GW.T @ (X / GWH) + np.sum(GW, axis=0, keepdims=True).T
            denum = gradg + sigmaR / H
        else:
            with np.errstate(divide="ignore", invalid="ignore"):
                num = GW.T @ (X / GWH)                         # updates.py:128
            if np.any(np.isnan(num)):                          # updates.py:129-131
                GWH = np.maximum(GWH, log_shift)
                num = GW.T @ (X / GWH)
            denum = np.sum(GW, axis=0, keepdims=True).T        # updates.py:132
        if not (np.isscalar(mu) and mu == 0):                  # updates.py:134-137
            denum = denum + _mu_column(mu) / (H + epsilon_reg)
        if lambda_L != 0:                                      # updates.py:138-141
            maxH = np.max(H, axis=1, keepdims=True)
            num = num + lambda_L * sigmaL * maxH
            denum = denum + lambda_L * sigmaL * maxH + lambda_L * HL
    num = H * num                                              # updates.py:142
    its = 0
    if simplex_H:                                              # updates.py:143-146
        nu, its = dichotomy_simplex(num, denum, log_shift=log_shift, tol=dicotomy_tol,
                                    return_its=True, force_its=force_its)
    else:
        nu = 0
    if safe:
        assert np.sum(denum < 0) == 0
        assert np.sum(num < 0) == 0
    new_H = np.maximum(num / (denum + nu), log_shift)          # updates.py:152
    if fixed_H is not None:                                    # updates.py:154-155
        new_H[fixed_H >= 0] = fixed_H[fixed_H >= 0]
    if return_its:
        return new_H, its
    return new_H


def multiplicative_step_hq(X, G, W, H, simplex_H=True, log_shift=LOG_SHIFT, safe=True,
                           dicotomy_tol=DICOTOMY_TOL, lambda_L=0, shape_2d=None, sigmaL=SIGMA_L,
                           fixed_H=None):
    """updates.py:263-301 (algo="l2_surrogate")."""
    if safe:
        assert np.sum(H < -log_shift / 2) == 0
        assert np.sum(W < -log_shift / 2) == 0
        assert np.sum(G < -log_shift / 2) == 0
    GW = G @ W
    GWH = GW @ H
    minus_c = H * (GW.T @ (X / (GWH + log_shift)))             # updates.py:280
    b = np.sum(GW, axis=0, keepdims=True).T                    # updates.py:282
    if lambda_L != 0:
        b = b + lambda_L * laplacian_apply(H, shape_2d) - lambda_L * sigmaL * H   # :284
        a = lambda_L * sigmaL
        if simplex_H:
            nu = dichotomy_simplex_acc(a, b, minus_c, log_shift=log_shift, tol=dicotomy_tol)
            b = b + nu
        new_H = (-b + np.sqrt(b ** 2 + 4 * a * minus_c)) / (2 * a)                # :289
    else:
        if simplex_H:
            nu = dichotomy_simplex(minus_c, b, log_shift=log_shift, tol=dicotomy_tol)
            b = b + nu
        new_H = minus_c / b
    new_H = np.maximum(new_H, log_shift)
    if fixed_H is not None:
        new_H[fixed_H >= 0] = fixed_H[fixed_H >= 0]
    return new_H


def multiplicative_step_w(X, G, W, H, simplex_W=False, log_shift=LOG_SHIFT, safe=True,
                          fixed_W=None, simplex_rows=None, dicotomy_tol=DICOTOMY_TOL, l2=False,
                          use_bregman=False):
    """updates.py:6-78.  ``simplex_rows`` stands for physics_model.NMF_simplex()."""
    if safe:
        assert np.sum(H < -log_shift / 2) == 0                  # updates.py:22-27
        assert np.sum(W < -log_shift / 2) == 0
        assert np.sum(G < -log_shift / 2) == 0
        H = np.maximum(H, log_shift)
        W = np.maximum(W, log_shift)
    if l2:                                                     # updates.py:29-36
        GGWHH = (G.T @ G) @ W @ (H @ H.T)
        GXH = G.T @ (X @ H.T)
        new_W = np.maximum(W / GGWHH * GXH, log_shift)
        if fixed_W is not None:
            new_W[fixed_W >= 0] = fixed_W[fixed_W >= 0]
        return new_W
    if use_bregman:                                            # updates.py:40-48
        GWH = (G @ W) @ H
        if np.allclose(G, np.eye(G.shape[0])) if G.shape[0] == G.shape[1] else False:
            sigmaR = np.sum(X, axis=1, keepdims=True)
        else:
            sigmaR = np.sum(X)
        num = sigmaR * W
        gradg = -G.T @ (X / GWH) @ H.T + np.sum(G, axis=0, keepdims=True).T @ np.sum(H, axis=1, keepdims=True).T
        denum = gradg * W + sigmaR
        new_W = np.maximum(num / denum, log_shift)
        if fixed_W is not None:
            new_W[fixed_W >= 0] = fixed_W[fixed_W >= 0]
        return new_W
    GW = G @ W                                                 # updates.py:38-39
    GWH = GW @ H
    with np.errstate(divide="ignore", invalid="ignore"):
        op1 = X / GWH                                          # updates.py:53
    if np.any(np.isnan(op1)):                                  # updates.py:54-56
        GWH = np.maximum(GWH, log_shift)
        op1 = X / GWH
    mult1 = G.T @ op1                                          # updates.py:58
    num = W * (mult1 @ H.T)                                    # updates.py:59
    denum = np.sum(G, axis=0, keepdims=True).T @ np.sum(H, axis=1, keepdims=True).T   # :60
    if simplex_W:                                              # updates.py:61-68
        # NB the reference uses the module constant dicotomy_tol here (updates.py:64,67)
        if simplex_rows is not None:
            idx = np.asarray(simplex_rows)
            nu = dichotomy_simplex(num[idx, :], denum[idx, :], log_shift=log_shift, tol=dicotomy_tol)
            denum[idx, :] = denum[idx, :] + nu
        else:
            nu = dichotomy_simplex(num, denum, log_shift=log_shift, tol=dicotomy_tol)
            denum = denum + nu
    new_W = num / denum                                        # updates.py:70
    new_W = np.maximum(new_W, log_shift)                       # updates.py:72
    if fixed_W is not None:                                    # updates.py:75-76
        new_W[fixed_W >= 0] = fixed_W[fixed_W >= 0]
    return new_W


def dichotomy_simplex_projected_gradient(a, log_shift=LOG_SHIFT, tol=DICOTOMY_TOL, maxit=MAXIT_DICHOTOMY,
                                         return_its=False):
    """dicotomy.py:83-108: root of sum_i max(a_i + x, log_shift) - 1 per column."""
    if log_shift > 0:
        if a.shape[0] * log_shift >= 1:
            raise ValueError("No solution exists!")
    nu_min = -np.max(a, axis=0)
    nu_max = 1 / a.shape[0] - np.min(a, axis=0)

    def func(x):
        return np.sum(np.maximum(a + x, log_shift), axis=0) - 1

    return dicotomy(nu_max, nu_min, func, maxit, tol, return_its=return_its)


def gradW(X, G, W, H, log_shift=LOG_SHIFT, safe=False, l2=False):
    """updates.py:303-314."""
    if safe:
        H = np.maximum(H, log_shift)
        W = np.maximum(W, log_shift)
    if l2:
        return 2 * G.T @ ((G @ W) @ H - X) @ H.T
    DH = (G @ W) @ H
    return G.T @ (-(X / DH) @ H.T + np.sum(H, axis=1, keepdims=True).T)


def gradH(X, G, W, H, mu=0, lambda_L=0, shape_2d=None, epsilon_reg=1, log_shift=LOG_SHIFT, safe=False,
          l2=False):
    """updates.py:316-345 (``L`` replaced by ``shape_2d``; L is symmetric so (lambda L H^T)^T = lambda H L)."""
    if lambda_L != 0:
        HL0 = laplacian_apply(H, shape_2d)                     # noqa: F841  (updates.py:320, unused there too)
    if safe:
        H = np.maximum(H, log_shift)
        W = np.maximum(W, log_shift)
    D = G @ W
    if l2:
        grad = D.T @ (D @ H - X)
    else:
        grad = -D.T @ (X / (D @ H)) + np.sum(D, axis=0, keepdims=True).T
    if not (np.isscalar(mu) and mu == 0):
        grad = grad + _mu_column(mu) / (H + epsilon_reg)
    if lambda_L != 0:
        grad = grad + laplacian_apply_scaled32(H, shape_2d, lambda_L)   # updates.py:343 with the (clamped) H
    return grad


def proj_grad_step_w(X, G, W, H, gamma, simplex_W=True, log_shift=LOG_SHIFT, safe=True, l2=False, fixed_W=None):
    """updates.py:347-367."""
    if safe:
        H = np.maximum(H, log_shift)
        W = np.maximum(W, log_shift)
    grad = gradW(X, G, W, H, log_shift=log_shift, safe=safe, l2=l2)
    new_W = np.maximum(W - 1 / gamma * grad, log_shift)
    if fixed_W is not None:
        new_W[fixed_W >= 0] = fixed_W[fixed_W >= 0]
    if simplex_W:
        raise NotImplementedError("Simplex constraint not implemented for W using the projected gradient method")
    return new_W


def proj_grad_step_h(X, G, W, H, gamma, simplex_H=True, mu=0, log_shift=LOG_SHIFT, epsilon_reg=1, safe=True,
                     dicotomy_tol=DICOTOMY_TOL, lambda_L=0, shape_2d=None, l2=False, fixed_H=None,
                     return_its=False):
    """updates.py:369-391."""
    if safe:
        H = np.maximum(H, log_shift)
        W = np.maximum(W, log_shift)
    grad = gradH(X, G, W, H, log_shift=log_shift, safe=safe, mu=mu, epsilon_reg=epsilon_reg, lambda_L=lambda_L,
                 shape_2d=shape_2d, l2=l2)
    new_H = H - 1 / gamma * grad
    its = 0
    if simplex_H:
        nu, its = dichotomy_simplex_projected_gradient(new_H, log_shift=log_shift, tol=dicotomy_tol, return_its=True)
    else:
        nu = 0
    new_H = np.maximum(new_H + nu, log_shift)
    if fixed_H is not None:
        new_H[fixed_H >= 0] = fixed_H[fixed_H >= 0]
    if return_its:
        return new_H, its
    return new_H


def estimate_Lipschitz_bound_w(log_shift, X, G, k):
    """updates.py:393-401."""
    if G is None:
        G = np.eye(X.shape[0])
    Wlim = np.ones([G.shape[1], k]) * log_shift
    Hlim = np.ones([k, X.shape[1]]) * log_shift
    DH = (G @ Wlim) @ Hlim
    return np.max((np.sum(Hlim, axis=0, keepdims=True) * X / (DH ** 2)) @ Hlim.T)


def estimate_Lipschitz_bound_h(log_shift, X, G, k, lambda_L=0, mu=0, epsilon_reg=1):
    """updates.py:403-413."""
    if G is None:
        G = np.eye(X.shape[0])
    Wlim = np.ones([G.shape[1], k]) * log_shift
    Hlim = np.ones([k, X.shape[1]]) * log_shift
    D = G @ Wlim
    DH = D @ Hlim
    return np.max(D.T @ (np.sum(D, axis=1, keepdims=True) * X / (DH ** 2))) + 2 * lambda_L + mu * epsilon_reg


# --------------------------------------------------------------------------------------------
# Laplacian surrogates for the line search (surrogates.py)
# --------------------------------------------------------------------------------------------
def smooth_l2_surrogate(Ht, shape_2d, H=None, sigmaL=SIGMA_L, lambda_L=1):
    """surrogates.py:5-53."""
    HtTL = laplacian_apply(Ht, shape_2d)
    t1 = np.sum(HtTL * Ht)
    if H is None:
        t2, t3 = t1, 0
    else:
        t2 = np.sum(HtTL * H)
        t3 = np.sum((Ht - H) ** 2)
    return lambda_L / 2 * (2 * t2 - t1 + sigmaL * t3)


def smooth_dgkl_surrogate(Ht, shape_2d, H=None, sigmaL=SIGMA_L, lambda_L=1):
    """surrogates.py:66-114."""
    HtTL = laplacian_apply(Ht, shape_2d)
    t1 = np.sum(HtTL * Ht)
    if H is None:
        t2, t3 = t1, 0
    else:
        t2 = np.sum(HtTL * H)
        maxH = np.max(H, axis=1)
        t3 = np.sum(maxH * np.sum(Ht * np.log(Ht / H) - Ht + H, axis=1))
    return lambda_L / 2 * (2 * t2 - t1 + sigmaL * t3)


def diff_surrogate(Ht, H, shape_2d, sigmaL=SIGMA_L, lambda_L=1, algo="log_surrogate"):
    """surrogates.py:116-149."""
    b_inf = trace_xtLx(H, shape_2d) * lambda_L / 2
    if algo in ("log_surrogate", "bmd"):
        b_supp = smooth_dgkl_surrogate(Ht, shape_2d, H=H, sigmaL=sigmaL, lambda_L=lambda_L)
    elif algo == "l2_surrogate":
        b_supp = smooth_l2_surrogate(Ht, shape_2d, H=H, sigmaL=sigmaL, lambda_L=lambda_L)
    else:
        raise ValueError("Unknown algorithm")
    return b_supp - b_inf


def quadratic_surrogate(x, xt, f_xt, gradf_xt, sigma):
    """surrogates.py:153-171."""
    return f_xt + np.sum((x - xt) * gradf_xt) + sigma * np.sum((x - xt) ** 2)


# --------------------------------------------------------------------------------------------
# Losses (measures.py)
# --------------------------------------------------------------------------------------------
def Frobenius_loss(X, W, H, average=False):
    """measures.py:350-385."""
    DH = W @ H
    if average:
        return np.mean((DH - X) ** 2)
    return np.sum((DH - X) ** 2)



def KLdiv_loss(X, GW, H, log_shift=LOG_SHIFT, average=False):
    """measures.py:456-504."""
    GW = np.maximum(GW, log_shift)
    H = np.maximum(H, log_shift)
    X = np.maximum(X, log_shift)
    Y = GW @ H
    if average:
        return np.mean(Y) - np.mean(X * np.log(Y))
    return np.sum(Y) - np.sum(X * np.log(Y))


def log_reg(H, mu, epsilon=1, average=False):
    """measures.py:524-548."""
    mu = _mu_column(mu)
    if average:
        return np.mean(mu * np.log(H + epsilon))
    return np.sum(mu * np.log(H + epsilon))


def const_KL(X, log_shift=LOG_SHIFT):
    """base.py:200-201."""
    return np.sum(X * np.log(np.maximum(X, log_shift))) - np.sum(X)


def full_loss(X, G, W, H, mu=0, epsilon_reg=1, lambda_L=0, shape_2d=None, log_shift=LOG_SHIFT,
              const=None, gamma=SIGMA_L, l2=False, average=True):
    """base.py:167-207 + smooth_nmf.py:457-475.

    Returns (loss, [kl, log_reg, lapl, gamma]) like ``detailed_loss_``.
    """
    n, p = X.shape
    numel = n * p if average else 1
    if l2:                                                      # base.py:197-198
        kl = 0.5 * Frobenius_loss(X, G @ W, H, average=False) / numel
    else:
        if const is None:
            const = const_KL(X, log_shift)
        kl = (KLdiv_loss(X, G @ W, H, log_shift, average=False) + const) / numel
    reg = log_reg(H, mu, epsilon_reg, average=False) / numel
    lap = 0.5 * lambda_L * trace_xtLx(H, shape_2d) / numel
    return kl + reg + lap, [kl, reg, lap, gamma]


# --------------------------------------------------------------------------------------------
# Prologue / epilogue helpers (base.py, utils.py)
# --------------------------------------------------------------------------------------------
def remove_zeros_lines(X, epsilon):
    """base.py:519-528."""
    if np.all(X >= 0):
        new_X = X.copy()
        sum_cols = X.sum(axis=0)
        sum_rows = X.sum(axis=1)
        new_X[:, np.where(sum_cols == 0)] = epsilon
        new_X[np.where(sum_rows == 0), :] = epsilon
        return new_X
    raise ValueError("Negative values in data")


def normalization_factor(X, nc):
    """base.py:16-18."""
    return nc / (np.mean(X) * X.shape[0])


def rescaled_DH(D, H):
    """utils.py:79-96."""
    from scipy.optimize import nnls
    _, p = H.shape
    o = np.ones((p,))
    s = np.linalg.lstsq(H.T, o, rcond=None)[0]
    if (s <= 0).any():
        s = np.maximum(nnls(H.T, o)[0], 1e-10)
    return D @ np.diag(1 / s), np.diag(s) @ H


def rel_change(new, old, tol):
    """base.py:323-324."""
    return np.max(np.abs(new - old) / (new + tol * np.mean(new)))


# --------------------------------------------------------------------------------------------
# The fit loop (base.py:209-420, smooth_nmf.py:284-455)
# --------------------------------------------------------------------------------------------
def fit(X, G, W0, H0, n_components=None, lambda_L=0.0, mu=0, epsilon_reg=1, simplex_H=False,
        simplex_W=True, tol=1e-4, max_iter=200, shape_2d=None, normalize=False,
        log_shift=LOG_SHIFT, dicotomy_tol=DICOTOMY_TOL, no_stop_criterion=False, fixed_H=None,
        fixed_W=None, algo="log_surrogate", debug=False, g_update=None, simplex_rows=None,
        gamma=None, l2=False, linesearch=False, true_D=None, true_H=None):
    """Run the reference's fit loop with user-supplied W0, H0 (NNDSVD is bypassed,
    updates.py:177-223 only clamps to log_shift in that case).

    ``g_update(W) -> G`` stands for PhysicalModel.NMF_update (called every 3rd iteration,
    base.py:388-390) and ``simplex_rows`` for PhysicalModel.NMF_simplex().
    Returns a dict with W, H, G, losses, detailed_losses, rel, n_iter, reconstruction_err, reason
    (+ gammas, true_losses when tracked).
    """
    X = remove_zeros_lines(np.asarray(X), log_shift)            # base.py:262
    n, p = X.shape
    k = W0.shape[1]
    norm_factor = None
    if normalize:                                               # base.py:264-267
        norm_factor = normalization_factor(X, k if n_components is None else n_components)
        X = norm_factor * X
    if G is None:                                               # updates.py:163-166
        G = np.diag(np.ones(n).astype(X.dtype))
    W = np.maximum(W0, log_shift)                               # updates.py:220-221
    H = np.maximum(H0, log_shift)
    if algo != "l2_surrogate":                                  # smooth_nmf.py:233-237
        l2 = False
    const = const_KL(X, log_shift)
    state = {"gamma": None}

    def gamma_report():
        g = state["gamma"]
        return g[0] if isinstance(g, list) else g

    def loss(G_, W_, H_, average=True, Xe=None, cst=None):
        return full_loss(X if Xe is None else Xe, G_, W_, H_, mu, epsilon_reg, lambda_L, shape_2d, log_shift,
                         const if cst is None else cst, gamma_report(), l2=l2, average=average)

    track = true_D is not None and true_H is not None and true_D.shape[1] == k and true_H.shape[0] == k
    if track:
        true_DH = true_D @ true_H
        # base.py:200-201 mixes the tracked X with log(self.X_) -- but const_KL_ is already set by eval_init
        true_losses = []
    eval_before = np.inf
    eval_init, _ = loss(G, W, H)                                # base.py:295 (gamma_ still None there)
    losses, detailed, rels, gammas = [], [], [], []
    n_iter = 0
    reason = None
    while True:                                                 # base.py:314
        old_W, old_H = W.copy(), H.copy()
        # ---------------- smooth_nmf.py:284-455 ----------------
        if n_iter == 0:                                         # smooth_nmf.py:290-306
            if gamma is None:
                if algo in ("l2_surrogate", "log_surrogate", "bmd"):
                    state["gamma"] = SIGMA_L
                else:
                    state["gamma"] = [estimate_Lipschitz_bound_h(log_shift, X, G, k, lambda_L=lambda_L, mu=mu,
                                                                 epsilon_reg=epsilon_reg),
                                      estimate_Lipschitz_bound_w(log_shift, X, G, k)]
            else:
                state["gamma"] = list(gamma) if isinstance(gamma, list) else gamma
        g_ = state["gamma"]
        Hold = H
        if algo == "l2_surrogate":
            H = multiplicative_step_hq(X, G, W, H, simplex_H=simplex_H, log_shift=log_shift,
                                       safe=debug, dicotomy_tol=dicotomy_tol, lambda_L=lambda_L,
                                       shape_2d=shape_2d, sigmaL=g_, fixed_H=fixed_H)
        elif algo in ("log_surrogate", "bmd"):
            H = multiplicative_step_h(X, G, W, H, simplex_H=simplex_H, mu=mu, log_shift=log_shift,
                                      epsilon_reg=epsilon_reg, safe=debug, dicotomy_tol=dicotomy_tol,
                                      lambda_L=lambda_L, shape_2d=shape_2d, sigmaL=g_,
                                      fixed_H=fixed_H, l2=l2, use_bregman=(algo == "bmd"))
        elif algo == "projected_gradient":
            H = proj_grad_step_h(X, G, W, H, g_[0], simplex_H=simplex_H, mu=mu, log_shift=log_shift,
                                 epsilon_reg=epsilon_reg, safe=debug, dicotomy_tol=dicotomy_tol,
                                 lambda_L=lambda_L, shape_2d=shape_2d, l2=l2, fixed_H=fixed_H)
        else:
            raise ValueError("Unknown algorithm")
        if linesearch:                                          # smooth_nmf.py:376-401
            if algo in ("l2_surrogate", "log_surrogate", "bmd"):
                d = diff_surrogate(Hold, H, shape_2d, sigmaL=g_, algo=algo)
                state["gamma"] = g_ / 1.05 if d > 0 else g_ * 1.5
            else:
                gradf_xt = gradH(X, G, W, Hold, mu=mu, lambda_L=lambda_L, shape_2d=shape_2d,
                                 epsilon_reg=epsilon_reg, log_shift=log_shift, safe=debug)
                f_xt, _ = loss(G, W, Hold, average=False)
                f_x, _ = loss(G, W, H, average=False)
                d = quadratic_surrogate(H, Hold, f_xt, gradf_xt, g_[0]) - f_x
                g_[0] = g_[0] / 1.05 if d > 0 else g_[0] * 1.5
        if algo in ("l2_surrogate", "log_surrogate", "bmd"):    # smooth_nmf.py:403-426
            W = multiplicative_step_w(X, G, W, H, simplex_W=simplex_W, log_shift=log_shift, safe=debug,
                                      fixed_W=fixed_W, simplex_rows=simplex_rows, l2=l2,
                                      use_bregman=(algo == "bmd"))
        else:                                                   # smooth_nmf.py:427-447
            Wold = W
            # NB the reference does not pass fixed_W here (smooth_nmf.py:430-437)
            W = proj_grad_step_w(X, G, W, H, g_[1], simplex_W=simplex_W, log_shift=log_shift, safe=debug)
            if linesearch:
                gradf_xt = gradW(X, G, Wold, H, log_shift=log_shift, safe=debug)
                f_xt, _ = loss(G, Wold, H, average=False)
                f_x, _ = loss(G, W, H, average=False)
                d = quadratic_surrogate(W, Wold, f_xt, gradf_xt, g_[1]) - f_x
                g_[1] = g_[1] / 1.05 if d > 0 else g_[1] * 1.5
        # ---------------- base.py:320-351 ----------------
        eval_after, det = loss(G, W, H)                         # base.py:320
        n_iter += 1
        rel_W = rel_change(W, old_W, tol)                       # base.py:323-324
        rel_H = rel_change(H, old_H, tol)
        if track:                                               # base.py:335-347
            Ht_ = H if (simplex_H or simplex_W) else rescaled_DH(W, H)[1]
            true_losses.append(loss(G, W, Ht_, Xe=true_DH)[0])
        losses.append(eval_after)
        detailed.append(det)
        gammas.append(gamma_report())
        rels.append([rel_W, rel_H])
        if n_iter >= max_iter:                                  # base.py:354-378
            reason = "max_iter"
            break
        if not no_stop_criterion:
            if max(rel_H, rel_W) < tol:
                reason = "rel"
                break
            elif abs((eval_before - eval_after) / eval_init) < tol:
                reason = "loss"
                break
            elif np.isnan(eval_after):
                reason = "nan"
                break
            elif (eval_before - eval_after) < 0:
                reason = "increase"
                break
        if g_update is not None and n_iter % 3 == 0:           # base.py:388-392
            G = g_update(W)
            eval_before, _ = loss(G, W, H)
        else:
            eval_before = eval_after
    if not simplex_H and not simplex_W:                         # base.py:399-400
        W, H = rescaled_DH(W, H)
    rec, _ = loss(G, W, H)                                      # base.py:407
    if normalize:                                               # base.py:409-410
        W = W / norm_factor
    out = dict(W=W, H=H, G=G, losses=np.array(losses), detailed_losses=np.array(detailed, dtype=float),
               rel=np.array(rels), n_iter=n_iter, reconstruction_err=rec, reason=reason,
               eval_init=eval_init, norm_factor=norm_factor, gammas=np.array(gammas, dtype=float))
    if track:
        out["true_losses"] = np.array(true_losses)
    return out
