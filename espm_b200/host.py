"""Host-side pieces of the fit that run once per fit on k x p / m x k data (NumPy).

They are part of the reference's prologue / epilogue: the
zero-line repair and normalisation of X, the initial factors, and the final column rescale.  None of
them is on the per-iteration path.
"""
import numpy as np


def is_physical_model(G):
    """Duck-typed ``isinstance(G, espm.models.base.PhysicalModel)`` (base.py:269): the three callbacks
    the loop uses are NMF_update / NMF_simplex / NMF_initialize_W (models/base.py:217-264)."""
    return (G is not None and not isinstance(G, np.ndarray)
            and all(hasattr(G, a) for a in ("NMF_update", "NMF_simplex", "NMF_initialize_W")))


def remove_zeros_lines(X, epsilon):
    """All-zero rows / columns of X are set to epsilon; negatives are rejected (base.py:519-528)."""
    if not np.all(X >= 0):
        raise ValueError("Negative values in data")
    out = X.copy()
    zero_cols = np.flatnonzero(X.sum(axis=0) == 0)
    zero_rows = np.flatnonzero(X.sum(axis=1) == 0)
    if zero_cols.size:
        out[:, zero_cols] = epsilon
    if zero_rows.size:
        out[zero_rows, :] = epsilon
    return out


def normalization_factor(X, nc):
    """base.py:16-18."""
    return nc / (np.mean(X) * X.shape[0])


def rescaled_DH(D, H):
    """Column rescale so that the columns of H sum to ~1 (utils.py:79-96)."""
    from scipy.optimize import nnls
    ones = np.ones((H.shape[1],))
    s = np.linalg.lstsq(H.T, ones, rcond=None)[0]
    if (s <= 0).any():
        s = np.maximum(nnls(H.T, ones)[0], 1e-10)
    return D @ np.diag(1 / s), np.diag(s) @ H


def initialize_factors(X, G, W, H, n_components, init, random_state, simplex_H, simplex_W, log_shift,
                       physics_model=None, nmf_init=None):
    """Initial (G, W, H) following updates.py:160-223.

    Missing factors come from scikit-learn's NMF initialisation plus least squares; user-supplied
    ones are only clamped to ``log_shift`` (updates.py:220-221).  Returns ``G`` as a dense array even
    when it is the identity (the ``G_`` attribute of the reference, updates.py:166).
    """
    identity = G is None
    if identity:
        G_dense = np.diag(np.ones(X.shape[0]).astype(X.dtype))
    else:
        G_dense = np.asarray(G)
    if W is None:
        if H is None:
            if nmf_init is not None:     # NNDSVD with the SVD on the device (init_device.py); X may be None then
                D, H = nmf_init(n_components, init, random_state)
            else:
                from sklearn.decomposition._nmf import _initialize_nmf
                D, H = _initialize_nmf(X, n_components=n_components, init=init, random_state=random_state)
            if simplex_H:
                H = np.nan_to_num(H, nan=1.0 / H.shape[0])
                scale = np.sum(H, axis=0, keepdims=True)
                H = H / scale
                D = D * np.mean(scale)
        else:
            D = np.abs(np.linalg.lstsq(H.T, X.T, rcond=None)[0].T)
        if identity:
            W = D
        elif physics_model is not None:
            W = physics_model.NMF_initialize_W(D)
            if simplex_W:
                idx = physics_model.NMF_simplex()
                W = np.nan_to_num(W, nan=1.0 / W.shape[0])
                W[idx, :] = W[idx, :] / np.sum(W[idx, :], axis=0, keepdims=True)
        else:
            W = np.abs(np.linalg.lstsq(G_dense, D, rcond=None)[0])
            if simplex_W:
                W = np.nan_to_num(W, nan=1.0 / W.shape[0])
                W = W / np.sum(W, axis=0, keepdims=True)
    elif H is None:
        D = G_dense @ W
        H = np.abs(np.linalg.lstsq(D, X, rcond=None)[0])
        if simplex_H:
            H = H / np.sum(H, axis=0, keepdims=True)
    W = np.maximum(W, log_shift)
    H = np.maximum(H, log_shift)
    return G_dense, W, H


# ---- ground-truth tracking (base.py:335-347; measures.py:10-50, 125-153, 166-207, 256-285, 579-627) -------------
def _unique_min(matrix):
    """measures.py:166-207: assignment with distinct rows that minimises the sum (brute force over permutations)."""
    from itertools import permutations
    n = matrix.shape[0]
    perms = list(permutations(range(n), n))
    sums = []
    for perm in perms:
        acc = 0
        for i in range(n):
            acc += matrix[perm[i], i]
        sums.append(acc)
    best = perms[sums.index(min(sums))]
    return [matrix[best[i], i] for i in range(n)], best


def find_min_angle(true_vectors, algo_vectors):
    """measures.py:125-153 with unique=True: spectral angles (degrees) of the best one-to-one match."""
    v1 = true_vectors / np.sqrt(np.sum(true_vectors ** 2, axis=1, keepdims=True))
    v2 = algo_vectors / np.sqrt(np.sum(algo_vectors ** 2, axis=1, keepdims=True))
    ang = np.arccos(np.clip(v1 @ v2.T, -1.0, 1.0)) * 180 / np.pi
    return _unique_min(ang)[0]


def find_min_MSE(true_maps, algo_maps):
    """measures.py:256-285 with unique=True on the squared distances of measures.py:579-627."""
    xx = (true_maps * true_maps).sum(axis=1)
    yy = (algo_maps * algo_maps).sum(axis=1)
    xy = np.dot(true_maps, algo_maps.T)
    d = abs(np.kron(np.ones((algo_maps.shape[0], 1)), xx).T + np.kron(np.ones((true_maps.shape[0], 1)), yy) - 2 * xy)
    return _unique_min(d / true_maps.shape[1])[0]
