"""Device-side NNDSVD initialisation (espm_b200/init_device.py) against scikit-learn's
``_initialize_nmf`` -- the function ``initialize_algorithms`` calls (espm/estimators/updates.py:3, 179) -- on the same
inputs and random state, and a default ``fit_transform(X)`` (no W, no H) with the initialisation on the device vs on
the host."""
import numpy as np
import pytest

gpu = pytest.mark.gpu


def problem(dtype, seed=5, n=96, nx=20, ny=24, k=3, m=7):
    rng = np.random.default_rng(seed)
    p = nx * ny
    x = np.linspace(0, 1, n)
    G = np.stack([np.exp(-0.5 * ((x - c) / 0.04) ** 2) for c in np.linspace(0.1, 0.9, m - 2)]
                 + [np.exp(-3 * x) + 0.05, (1 - x) * 0.5 + 0.05], axis=1)
    Wt = rng.uniform(size=(m, k)) ** 2
    Ht = rng.uniform(size=(k, p)) ** 3
    Ht /= Ht.sum(0, keepdims=True)
    lam = G @ Wt @ Ht
    X = rng.poisson(lam / lam.sum(0, keepdims=True) * 60.0).astype(dtype)
    return X, G.astype(dtype), (nx, ny), k


def engine_for(X, G, k):
    from espm_b200.engine import FitEngine
    n, p = X.shape
    m = n if G is None else G.shape[1]
    return FitEngine(X, G, np.ones((m, k), dtype=X.dtype), np.ones((k, p), dtype=X.dtype), max_records=16,
                     ingest=dict(eps=1e-14, normalize=None))


@gpu
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-9), (np.float32, 2e-3)])
@pytest.mark.parametrize("init", [None, "nndsvd", "nndsvdar", "random"])
def test_initialize_nmf_device_matches_sklearn(dtype, tol, init):
    from sklearn.decomposition._nmf import _initialize_nmf
    from espm_b200.host import remove_zeros_lines
    from espm_b200.init_device import initialize_nmf_device
    X, G, _, k = problem(dtype)
    eng = engine_for(X, G, k)
    Wd, Hd = initialize_nmf_device(eng, k, init, random_state=3)
    Ws, Hs = _initialize_nmf(remove_zeros_lines(X, 1e-14), k, init=init, random_state=3)
    assert Wd.shape == Ws.shape and Hd.shape == Hs.shape and Wd.dtype == Ws.dtype
    scale_w, scale_h = np.abs(Ws).max(), np.abs(Hs).max()
    assert np.max(np.abs(Wd - Ws)) <= tol * scale_w
    assert np.max(np.abs(Hd - Hs)) <= tol * scale_h
    # the sign split must have taken the same branches: identical zero / filled patterns for plain nndsvd
    if init == "nndsvd":
        assert np.array_equal(Wd == 0, Ws == 0) and np.array_equal(Hd == 0, Hs == 0)


@gpu
@pytest.mark.parametrize("identity", [False, True])
def test_default_fit_device_init_equals_host_init(identity):
    import espm_b200
    from espm_b200 import SmoothNMF
    X, G, shape_2d, k = problem(np.float64)
    kw = dict(n_components=k, G=None if identity else G, shape_2d=shape_2d, lambda_L=1.0, mu=0.02,
              simplex_H=not identity, simplex_W=identity, tol=0, no_stop_criterion=True, max_iter=8, verbose=0,
              random_state=11)
    out = {}
    for flag in (True, False):
        espm_b200.config.device_init = flag
        try:
            est = SmoothNMF(**kw)
            est.fit_transform(X)
        finally:
            espm_b200.config.device_init = True
        out[flag] = (np.array(est.losses_), est.W_, est.H_)
    l_dev, W_dev, H_dev = out[True]
    l_host, W_host, H_host = out[False]
    assert np.max(np.abs(l_dev - l_host) / np.abs(l_host)) < 1e-7
    assert np.max(np.abs(W_dev - W_host)) <= 1e-6 * np.abs(W_host).max()
    assert np.max(np.abs(H_dev - H_host)) <= 1e-6 * np.abs(H_host).max()


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 1e-4)])
def test_init_host_logic_with_cpu_tensors(dtype, tol):
    """The host side of init_device.py (random stream, LU row permutation, svd_flip, NNDSVD split, tile-major
    indexing) with the tensor algebra on CPU tensors: no CUDA involved, no engine -- a stand-in object carries Xt."""
    import types
    import torch
    from sklearn.decomposition._nmf import _initialize_nmf
    from espm_b200.host import remove_zeros_lines
    from espm_b200.init_device import initialize_nmf_device
    X, _, _, k = problem(dtype)
    Xp = remove_zeros_lines(X, 1e-14)
    n, p = Xp.shape
    n_pad, nt = (n + 31) // 32 * 32, (p + 127) // 128
    Xt = np.zeros((nt, n_pad, 128), dtype=dtype)
    for t in range(nt):
        w = min(128, p - t * 128)
        Xt[t, :n, :w] = Xp[:, t * 128:t * 128 + w]
    from espm_b200 import _lib as L
    code = L.F64 if dtype == np.float64 else L.F32
    eng = types.SimpleNamespace(n=n, p=p, p_loc=p, j0=0, shard=None, Xt=torch.from_numpy(Xt.reshape(-1)), x_code=code,
                                c_code=code,
                                st=types.SimpleNamespace(n_pad=n_pad, n_tiles=nt))
    for init in (None, "nndsvd", "nndsvdar", "random"):
        Wd, Hd = initialize_nmf_device(eng, k, init, random_state=3)
        Ws, Hs = _initialize_nmf(Xp, k, init=init, random_state=3)
        assert Wd.dtype == Ws.dtype and Hd.dtype == Hs.dtype
        assert np.max(np.abs(Wd - Ws)) <= tol * np.abs(Ws).max()
        assert np.max(np.abs(Hd - Hs)) <= tol * np.abs(Hs).max()
