"""Operator-level API: the functions of the reference that sit on the fit-loop path, with the same
names and argument meaning, executed by the CUDA kernels.

    multiplicative_step_h / multiplicative_step_w   espm/estimators/updates.py:83 / :6
    multiplicative_step_hq                           espm/estimators/updates.py:263
    dichotomy_simplex                                espm/estimators/dicotomy.py:4
    KLdiv_loss / log_reg / trace_xtLx                espm/measures.py:456 / :524 / :560
    create_laplacian_matrix                          espm/utils.py:39 (matrix-like object that also carries the
                                                     image shape: the kernels apply a stencil)

Inputs and outputs are NumPy arrays on the host (like the reference); each call uploads, runs the
kernels once and downloads -- these entry points exist for parity tests and small problems, the
estimator keeps everything resident instead.  No CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _lib as _L
from .conf import dicotomy_tol as _TOL
from .conf import log_shift as _LS
from .conf import maxit_dichotomy as _MAXIT
from .conf import sigmaL as _SIGMA
from .engine import FitEngine


class GridLaplacian:
    """The matrix ``create_laplacian_matrix`` returns in the reference (utils.py:39-76), or the identity the fit
    uses when ``shape_2d`` is None (base.py:289-291).

    The kernels apply it as a 5-point stencil and only ever need ``shape_2d``; everything a caller can do with
    the reference's scipy.sparse matrix (``L @ x``, ``x @ L``, ``.dot``, ``.toarray()``, ``.shape``, indexing,
    ...) works as well: the float32 sparse matrix is assembled on first use and every other attribute is
    delegated to it."""

    __array_priority__ = 10.1        # like scipy.sparse: ndarray @ L defers to __rmatmul__
    __array_ufunc__ = None
    ndim = 2

    def __init__(self, nx, ny=None, identity=None):
        if identity is not None:
            self.shape_2d = None
            self._p = int(identity)
        else:
            self.shape_2d = (int(nx), int(ny))
            self._p = int(nx) * int(ny)
        self._mat = None

    @property
    def shape(self):
        return (self._p, self._p)

    @property
    def dtype(self):
        return np.dtype(np.float32)

    def tosparse(self):
        """scipy.sparse matrix with the reference's entries: diagonal = number of in-bounds 4-neighbours,
        -1 for each neighbour, pixel index = i * ny + j (float32, utils.py:57)."""
        if self._mat is None:
            import scipy.sparse as sp
            if self.shape_2d is None:
                self._mat = sp.identity(self._p, dtype=np.float32, format="csr")
            else:
                def path(n):    # 1-D Neumann Laplacian of a path with n nodes
                    d = np.full(n, 2.0, dtype=np.float32)
                    d[0] = d[-1] = 1.0
                    o = np.full(n - 1, -1.0, dtype=np.float32)
                    return sp.diags([o, d, o], [-1, 0, 1], format="csr", dtype=np.float32)
                nx, ny = self.shape_2d
                eye = lambda n: sp.identity(n, dtype=np.float32, format="csr")   # noqa: E731
                self._mat = (sp.kron(eye(nx), path(ny)) + sp.kron(path(nx), eye(ny))).tocsr().astype(np.float32)
        return self._mat

    def __matmul__(self, other):
        return self.tosparse() @ other

    def __rmatmul__(self, other):
        return other @ self.tosparse()

    def dot(self, other):
        return self.tosparse().dot(other)

    def __mul__(self, other):
        return self.tosparse() * other

    def __rmul__(self, other):
        return other * self.tosparse()

    def __getitem__(self, idx):
        return self.tosparse()[idx]

    def __getattr__(self, name):
        if name.startswith("__") or name in ("_mat", "_p", "shape_2d"):
            raise AttributeError(name)
        return getattr(self.tosparse(), name)


def create_laplacian_matrix(nx, ny=None):
    """utils.py:39-76.  Returns an object that behaves like the reference's sparse matrix and carries the image
    shape for the kernels (which apply the 5-point Neumann Laplacian as a stencil)."""
    if ny is None:
        ny = nx
    assert nx > 1
    assert ny > 1
    return GridLaplacian(nx, ny)


def check_shape_2d(shape_2d, p):
    if shape_2d is None:
        return None
    nx, ny = int(shape_2d[0]), int(shape_2d[1])
    if nx * ny != p:
        raise ValueError("shape_2d %s does not match the number of pixels %d" % (shape_2d, p))
    return nx, ny


def _shape_from_L(Lm, p):
    """Accept what the reference accepts for ``L``: our GridLaplacian, a (nx, ny) tuple, None, or the
    reference's own matrix (scipy.sparse / dense) -- identity or grid Laplacian -- whose shape is inferred."""
    if Lm is None:
        return None, False
    if isinstance(Lm, GridLaplacian):
        if Lm.shape != (p, p):
            raise ValueError("Laplacian has shape %s, expected %s" % (Lm.shape, (p, p)))
        return Lm.shape_2d, True
    if isinstance(Lm, tuple):
        return (int(Lm[0]), int(Lm[1])), True
    try:
        import scipy.sparse as sp
        M = sp.csr_matrix(Lm)
    except Exception as exc:  # pragma: no cover
        raise TypeError("unsupported Laplacian object %r" % type(Lm)) from exc
    if M.shape != (p, p):
        raise ValueError("Laplacian has shape %s, expected %s" % (M.shape, (p, p)))
    if M.nnz == p and np.all(M.diagonal() == 1):
        return None, True                       # identity (base.py:289-291)
    row0 = M.getrow(0).indices
    ny = int(row0.max())
    if ny < 1 or p % ny != 0:
        raise ValueError("cannot infer the image shape from the Laplacian matrix")
    nx = p // ny
    if p <= 4096:
        from scipy.sparse import lil_matrix
        ref = lil_matrix((p, p), dtype=np.float64)
        for i in range(nx):
            for j in range(ny):
                a = i * ny + j
                for di, dj in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                    ii, jj = i + di, j + dj
                    if 0 <= ii < nx and 0 <= jj < ny:
                        ref[a, a] += 1
                        ref[a, ii * ny + jj] = -1
        if abs(ref.tocsr() - M.astype(np.float64)).sum() != 0:
            raise ValueError("L is neither the identity nor the grid Laplacian of utils.py:39-76")
    return (nx, ny), True


def _engine(X, G, W, H, **kw):
    X = np.asarray(X)
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    return FitEngine(X, None if G is None else np.asarray(G), np.asarray(W), np.asarray(H), **kw)


def _is_identity(G):
    return G is not None and G.shape[0] == G.shape[1] and np.array_equal(G, np.eye(G.shape[0], dtype=G.dtype))


def _raise_flags(rec):
    flags = int(rec[_L.S_DEV_FLAGS])
    if flags & _L.DEV_NONFINITE:
        raise FloatingPointError("espm_b200: non-finite ratio sums (zero row in G W?)")
    if flags & (_L.DEV_BRACKET | _L.DEV_NEGATIVE):
        raise AssertionError("espm_b200: bisection preconditions violated (dicotomy.py:17-19,141-144)")


def multiplicative_step_h(X, G, W, H, simplex_H=False, mu=0, log_shift=_LS, epsilon_reg=1, safe=True,
                          dicotomy_tol=_TOL, lambda_L=0, L=None, l2=False, sigmaL=_SIGMA, fixed_H=None,
                          use_bregman=False, return_its=False):
    """updates.py:83-156: the KL, Frobenius (``l2``) and Bregman (``use_bregman``) branches."""
    if l2:                                                            # updates.py:110-114
        assert lambda_L == 0
        assert (mu == 0) if np.isscalar(mu) else (np.asarray(mu) == 0).all()
    p = np.shape(H)[1]
    shape_2d = None
    if lambda_L != 0:
        if L is None:
            raise ValueError("Please provide the laplacian")          # updates.py:94-95
        shape_2d, _ = _shape_from_L(L, p)
    if simplex_H and log_shift > 0 and np.shape(H)[0] * log_shift >= 1:
        raise ValueError("No solution exists!")                       # dicotomy.py:22-23
    Ge = None if (G is None or _is_identity(np.asarray(G))) else G
    eng = _engine(X, Ge, W, H, shape_2d=shape_2d, lambda_L=lambda_L, mu=mu, epsilon_reg=epsilon_reg,
                  log_shift=log_shift, dicotomy_tol=dicotomy_tol, sigma=sigmaL, simplex_H=simplex_H,
                  simplex_W=False, fixed_H=fixed_H, max_records=8, clamp_init=bool(safe),   # updates.py:104-105
                  algo="bmd" if (use_bregman and not l2) else "log_surrogate", l2_h=bool(l2))
    Hn, rec = eng.step_h_only()
    if int(rec[_L.S_DEV_FLAGS]) & _L.DEV_NONFINITE:      # updates.py:129-131: NaN -> GWH = max(GWH, log_shift)
        eng.enable_clamp()
        Hn, rec = eng.step_h_only()
    _raise_flags(rec)
    if return_its:
        return Hn, int(rec[_L.S_BISECT_ITS_H])
    return Hn


def multiplicative_step_hq(X, G, W, H, simplex_H=True, log_shift=_LS, safe=True, dicotomy_tol=_TOL, lambda_L=0,
                           L=None, sigmaL=_SIGMA, fixed_H=None, return_its=False):
    """updates.py:263-301: the quadratic-surrogate H step of ``algo="l2_surrogate"``."""
    p = np.shape(H)[1]
    shape_2d = None
    if lambda_L != 0:
        if L is None:
            raise ValueError("Please provide the laplacian")          # updates.py:267-269
        shape_2d, _ = _shape_from_L(L, p)
    if simplex_H and log_shift > 0 and np.shape(H)[0] * log_shift >= 1:
        raise ValueError("No solution exists!")                       # dicotomy.py:71-72
    Ge = None if (G is None or _is_identity(np.asarray(G))) else G
    # updates.py:263-301 never clamps W, H (only asserts when safe): upload them as given
    eng = _engine(X, Ge, W, H, shape_2d=shape_2d, lambda_L=lambda_L, log_shift=log_shift,
                  dicotomy_tol=dicotomy_tol, sigma=sigmaL, simplex_H=simplex_H, simplex_W=False, fixed_H=fixed_H,
                  max_records=8, algo="l2_surrogate", clamp_init=False)
    Hn, rec = eng.step_h_only()
    _raise_flags(rec)
    if return_its:
        return Hn, int(rec[_L.S_BISECT_ITS_H])
    return Hn


def multiplicative_step_w(X, G, W, H, simplex_W=False, log_shift=_LS, safe=True, l2=False, fixed_W=None,
                          physics_model=None, use_bregman=False):
    """updates.py:6-78: the KL, Frobenius (``l2``) and Bregman (``use_bregman``) branches."""
    bmd = bool(use_bregman) and not l2
    if bmd:
        Gm = np.asarray(G)
        if np.allclose(Gm, np.eye(Gm.shape[0])):     # updates.py:42 (raises for a non-square G, like the reference)
            G = None
    if l2 or bmd:
        simplex_W = False                             # those branches never project (updates.py:29-48)
    rows = None
    if simplex_W and physics_model is not None:
        rows = physics_model.NMF_simplex()
    nrows = len(rows) if rows is not None else np.shape(W)[0]
    if simplex_W and log_shift > 0 and nrows * log_shift >= 1:
        raise ValueError("No solution exists!")
    Ge = None if (G is None or _is_identity(np.asarray(G))) else G
    eng = _engine(X, Ge, W, H, log_shift=log_shift, simplex_H=False, simplex_W=simplex_W, simplex_rows=rows,
                  fixed_W=fixed_W, max_records=8, clamp_init=bool(safe),                        # updates.py:26-27
                  algo="bmd" if bmd else "log_surrogate", l2=bool(l2))
    Wn, rec = eng.step_w_only()
    if int(rec[_L.S_DEV_FLAGS]) & _L.DEV_NONFINITE:      # updates.py:54-56
        eng.enable_clamp()
        Wn, rec = eng.step_w_only()
    _raise_flags(rec)
    return Wn


def gradH(X, G, W, H, mu=0, lambda_L=0, L=None, epsilon_reg=1, log_shift=_LS, safe=False, l2=False):
    """updates.py:316-345: gradient of the (KL or Frobenius) data term + regularisers with respect to H."""
    p = np.shape(H)[1]
    shape_2d = None
    if lambda_L != 0:
        if L is None:
            raise ValueError("Please provide the laplacian")
        shape_2d, _ = _shape_from_L(L, p)
    Ge = None if (G is None or _is_identity(np.asarray(G))) else G
    eng = _engine(X, Ge, W, H, shape_2d=shape_2d, lambda_L=lambda_L, mu=mu, epsilon_reg=epsilon_reg,
                  log_shift=log_shift, simplex_H=False, simplex_W=False, max_records=8, clamp_init=bool(safe),
                  algo="projected_gradient", l2_h=bool(l2), gamma_pg=(1.0, 1.0))
    eng.evaluate(0)
    return eng.den[:eng.k, :eng.p_loc].cpu().numpy()


def gradW(X, G, W, H, log_shift=_LS, safe=False, l2=False):
    """updates.py:303-314: gradient of the data term with respect to W."""
    Ge = None if (G is None or _is_identity(np.asarray(G))) else G
    eng = _engine(X, Ge, W, H, log_shift=log_shift, simplex_H=False, simplex_W=False, max_records=8,
                  clamp_init=bool(safe), algo="projected_gradient", l2=bool(l2), gamma_pg=(1.0, 1.0))
    eng.step_w_only()
    return eng.w_den.cpu().numpy()


def proj_grad_step_h(X, G, W, H, gamma, simplex_H=True, mu=0, log_shift=_LS, epsilon_reg=1, safe=True,
                     dicotomy_tol=_TOL, lambda_L=0, L=None, l2=False, fixed_H=None, return_its=False):
    """updates.py:369-391: gradient step in H, then projection on the simplex (dicotomy.py:83-108) / the orthant."""
    p = np.shape(H)[1]
    shape_2d = None
    if lambda_L != 0:
        if L is None:
            raise ValueError("Please provide the laplacian")
        shape_2d, _ = _shape_from_L(L, p)
    if simplex_H and log_shift > 0 and np.shape(H)[0] * log_shift >= 1:
        raise ValueError("No solution exists!")                       # dicotomy.py:94-96
    Ge = None if (G is None or _is_identity(np.asarray(G))) else G
    eng = _engine(X, Ge, W, H, shape_2d=shape_2d, lambda_L=lambda_L, mu=mu, epsilon_reg=epsilon_reg,
                  log_shift=log_shift, dicotomy_tol=dicotomy_tol, simplex_H=simplex_H, simplex_W=False,
                  fixed_H=fixed_H, max_records=8, clamp_init=bool(safe), algo="projected_gradient", l2_h=bool(l2),
                  gamma_pg=(float(gamma), 1.0))
    Hn, rec = eng.step_h_only()
    if int(rec[_L.S_DEV_FLAGS]) & _L.DEV_BRACKET:
        raise AssertionError("espm_b200: bisection preconditions violated (dicotomy.py:141-144)")
    if return_its:
        return Hn, int(rec[_L.S_BISECT_ITS_H])
    return Hn


def proj_grad_step_w(X, G, W, H, gamma, simplex_W=True, log_shift=_LS, safe=True, l2=False, fixed_W=None):
    """updates.py:347-367: gradient step in W and projection on the orthant."""
    Ge = None if (G is None or _is_identity(np.asarray(G))) else G
    eng = _engine(X, Ge, W, H, log_shift=log_shift, simplex_H=False, simplex_W=False, fixed_W=fixed_W, max_records=8,
                  clamp_init=bool(safe), algo="projected_gradient", l2=bool(l2), gamma_pg=(1.0, float(gamma)))
    Wn, _ = eng.step_w_only()
    if simplex_W:                                                     # updates.py:365-366 (after the step, like there)
        raise NotImplementedError("Simplex constraint not implemented for W using the projected gradient method")
    return Wn


def estimate_Lipschitz_bound_w(log_shift, X, G, k):
    """updates.py:393-401.  With W = H = log_shift everywhere the bound is max_c rowsum(X)_c / (k ls^2 rowsum(G)_c^2);
    the sums of X come from the device."""
    X = np.asarray(X)
    eng = _engine(X, None, np.ones((X.shape[0], k), dtype=X.dtype if X.dtype in (np.float32, np.float64) else float),
                  np.ones((k, X.shape[1])), simplex_H=False, simplex_W=False, max_records=8)
    _, rs = eng.compute_x_sums()
    rs = rs[:X.shape[0]].cpu().numpy().astype(np.float64)
    rsG = np.ones(X.shape[0]) if G is None else np.asarray(G, dtype=np.float64).sum(1)
    return float(np.max(rs / (k * log_shift ** 2 * rsG ** 2)))


def estimate_Lipschitz_bound_h(log_shift, X, G, k, lambda_L=0, mu=0, epsilon_reg=1):
    """updates.py:403-413: max_j colsum(X)_j / (k ls^2) + 2 lambda_L + mu eps (G cancels)."""
    X = np.asarray(X)
    eng = _engine(X, None, np.ones((X.shape[0], k), dtype=X.dtype if X.dtype in (np.float32, np.float64) else float),
                  np.ones((k, X.shape[1])), simplex_H=False, simplex_W=False, max_records=8)
    cs, _ = eng.compute_x_sums()
    return float(cs[:X.shape[1]].max().item()) / (k * log_shift ** 2) + 2 * lambda_L + mu * epsilon_reg


def dichotomy_simplex_projected_gradient(a, log_shift=_LS, tol=_TOL, maxit=_MAXIT, return_its=False):
    """dicotomy.py:83-108: per-column root of sum_i max(a_i + x, log_shift) - 1."""
    lib = _L.load()
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    k, p = a.shape
    if log_shift > 0 and k * log_shift >= 1:
        raise ValueError("No solution exists!")
    dev = torch.device("cuda", torch.cuda.current_device())
    d_a = torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    nu = torch.empty(p, dtype=torch.float64, device=dev)
    mask = torch.zeros(4, dtype=torch.int32, device=dev)
    flags = torch.zeros(4, dtype=torch.int32, device=dev)
    its = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _L.check(lib.espm_dichotomy_simplex_pg(_L.F64, k, p, d_a.data_ptr(), float(log_shift), float(tol), int(maxit),
                                          nu.data_ptr(), mask.data_ptr(), flags.data_ptr(), its.data_ptr(), stream))
    if int(flags[0].item()) & _L.DEV_BRACKET:
        raise AssertionError("dichotomy_simplex_projected_gradient: preconditions violated (dicotomy.py:141-144)")
    out = nu.cpu().numpy()
    if return_its:
        return out, int(its.item())
    return out


def dichotomy_simplex(num, denum, log_shift=_LS, tol=_TOL, maxit=_MAXIT, return_its=False):
    """dicotomy.py:4-55: per-column root of sum_i max(num_i/(x+den_i), log_shift) - 1 with the
    reference's bracket and its GLOBAL (lock-step) stop test."""
    lib = _L.load()
    num = np.asarray(num)
    denum = np.asarray(denum)
    if num.ndim == 1:
        num, denum = num[:, None], denum[:, None]
    k, p = num.shape
    if log_shift > 0 and k * log_shift >= 1:
        raise ValueError("No solution exists!")
    dt = np.result_type(num.dtype, denum.dtype, np.float32)
    code = _L.F64 if dt == np.float64 else _L.F32
    tdt = torch.float64 if code == _L.F64 else torch.float32
    dev = torch.device("cuda", torch.cuda.current_device())
    d_num = torch.as_tensor(np.ascontiguousarray(num), dtype=tdt).to(dev)
    d_den = torch.as_tensor(np.ascontiguousarray(denum), dtype=tdt).to(dev)
    nu = torch.empty(p, dtype=tdt, device=dev)
    mask = torch.zeros(4, dtype=torch.int32, device=dev)
    flags = torch.zeros(4, dtype=torch.int32, device=dev)
    its = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _L.check(lib.espm_dichotomy_simplex(code, k, p, d_num.data_ptr(), d_den.data_ptr(), float(log_shift), float(tol),
                                       int(maxit), nu.data_ptr(), mask.data_ptr(), flags.data_ptr(),
                                       its.data_ptr(), stream))
    f = int(flags[0].item())
    if f & (_L.DEV_BRACKET | _L.DEV_NEGATIVE):
        raise AssertionError("dichotomy_simplex: preconditions violated (dicotomy.py:17-19,141-144)")
    out = nu.cpu().numpy()
    n_it = int(its.item())
    if n_it >= maxit:
        print("Dicotomy stopped for maximum number of iterations")
    if return_its:
        return out, n_it
    return out


def dichotomy_simplex_acc(a, b, minus_c, log_shift=_LS, tol=_TOL, maxit=_MAXIT, return_its=False):
    """dicotomy.py:57-81: the bisection of the quadratic-surrogate H step, with the reference's bracket
    and lock-step stop test."""
    lib = _L.load()
    assert a >= 0                                                       # dicotomy.py:67-68
    b = np.asarray(b, dtype=np.float64)
    minus_c = np.asarray(minus_c, dtype=np.float64)
    assert (minus_c >= 0).all()
    if b.ndim == 1:
        b, minus_c = b[:, None], minus_c[:, None]
    k, p = b.shape
    if log_shift > 0 and k * log_shift >= 1:
        raise ValueError("No solution exists!")
    dev = torch.device("cuda", torch.cuda.current_device())
    d_b = torch.as_tensor(np.ascontiguousarray(b)).to(dev)
    d_c = torch.as_tensor(np.ascontiguousarray(minus_c)).to(dev)
    nu = torch.empty(p, dtype=torch.float64, device=dev)
    mask = torch.zeros(4, dtype=torch.int32, device=dev)
    flags = torch.zeros(4, dtype=torch.int32, device=dev)
    its = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _L.check(lib.espm_dichotomy_simplex_acc(_L.F64, k, p, float(a), d_b.data_ptr(), d_c.data_ptr(), float(log_shift),
                                           float(tol), int(maxit), nu.data_ptr(), mask.data_ptr(), flags.data_ptr(),
                                           its.data_ptr(), stream))
    if int(flags[0].item()) & (_L.DEV_BRACKET | _L.DEV_NEGATIVE):
        raise AssertionError("dichotomy_simplex_acc: preconditions violated (dicotomy.py:67-68,141-144)")
    out = nu.cpu().numpy()
    n_it = int(its.item())
    if n_it >= maxit:
        print("Dicotomy stopped for maximum number of iterations")
    if return_its:
        return out, n_it
    return out


def full_loss(X, G, W, H, mu=0, epsilon_reg=1, lambda_L=0, shape_2d=None, log_shift=_LS, const=0.0,
              average=True, l2=False, clamp=True):
    """KL (or 0.5 Frobenius, ``l2``) + log-reg + Laplacian loss of (W, H) (base.py:167-207, smooth_nmf.py:457-475).
    Returns (loss, [kl, log_reg, lapl]).  ``clamp=False`` uploads W and H as given: G W and H are then clamped to
    ``log_shift`` inside the KL term only (measures.py:493-495), exactly like ``SmoothNMF.loss``; the default clamps
    the factors themselves first, which is what the fit loop evaluates (its iterates are >= log_shift)."""
    eng = _engine(X, G, W, H, shape_2d=shape_2d, lambda_L=lambda_L, mu=mu, epsilon_reg=epsilon_reg,
                  log_shift=log_shift, simplex_H=False, simplex_W=False, max_records=8, l2=bool(l2),
                  clamp_init=clamp and not l2)
    if not clamp and not l2:
        eng.set_flag(_L.FLAG_LOSS_DUAL)
    eng.evaluate(0)
    rec = eng.read_records(0, 1)[0]
    numel = X.shape[0] * np.shape(H)[1] if average else 1
    kl, reg, lap = eng.loss_parts(rec, const, numel)
    if l2:
        kl = 0.5 * rec[_L.S_XLOGY] / numel
    return kl + reg + lap, [kl, reg, lap]


def Frobenius_loss(X, W, H, average=False):
    """measures.py:350-385 with ``W`` playing the role of G W (n x k): sum (W H - X)^2."""
    X = np.asarray(X)
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    eng = _engine(X, None, W, H, log_shift=0.0, simplex_H=False, simplex_W=False, max_records=8, l2=True,
                  clamp_init=False)
    eng.evaluate(0)
    val = eng.read_records(0, 1)[0][_L.S_XLOGY]
    return val / X.size if average else val


def KLdiv_loss(X, W, H, log_shift=_LS, average=False):
    """measures.py:456-504 with ``W`` playing the role of G W (n x k)."""
    X = np.asarray(X)
    # the engine clamps W, H to log_shift on upload, which is exactly measures.py:493-494
    eng = _engine(X, None, W, H, log_shift=log_shift, simplex_H=False, simplex_W=False, max_records=8)
    eng.evaluate(0)
    rec = eng.read_records(0, 1)[0]
    val = rec[_L.S_SUMY] - rec[_L.S_XLOGY]
    return val / X.size if average else val


def log_reg(H, mu, epsilon=1, average=False):
    """measures.py:524-548, evaluated by the h_finish kernel."""
    H = np.asarray(H)
    k, p = H.shape
    X = np.ones((1, p), dtype=H.dtype if H.dtype in (np.float32, np.float64) else np.float64)
    eng = _engine(X, None, np.ones((1, k), dtype=X.dtype), H, mu=mu, epsilon_reg=epsilon, log_shift=0.0,
                  simplex_H=False, simplex_W=False, max_records=8, clamp_init=False)
    eng.evaluate(0)
    val = eng.read_records(0, 1)[0][_L.S_LOGREG]
    return val / H.size if average else val


def trace_xtLx(Lm, x, average=False):
    """measures.py:560-577 for the grid (or identity) Laplacian; ``x`` is H.T (p x k) as in
    smooth_nmf.py:466."""
    x = np.asarray(x)
    H = np.ascontiguousarray(x.T)
    k, p = H.shape
    shape_2d, _ = _shape_from_L(Lm, p)
    X = np.ones((1, p), dtype=H.dtype if H.dtype in (np.float32, np.float64) else np.float64)
    eng = _engine(X, None, np.ones((1, k), dtype=X.dtype), H, lambda_L=1.0, shape_2d=shape_2d, log_shift=0.0,
                  simplex_H=False, simplex_W=False, max_records=8, clamp_init=False)
    eng.evaluate(0)
    val = eng.read_records(0, 1)[0][_L.S_LAPL]
    return val / x.size if average else val
