"""Compact storage of count data: X held as uint8 / uint16 on the device when it contains
only integers in range (EDXS spectrum images are Poisson counts, datasets/base.py:68), selected by a device pre-scan.
The arithmetic is the fp32 path's, entry for entry (the conversion is exact), so the results must equal the dense
fp32 storage's; against the fp64 oracle they meet the fp32 tolerances of the north star."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture
def storage():
    import espm_b200
    keep = espm_b200.config.x_storage

    def set_mode(mode):
        espm_b200.config.x_storage = mode
    yield set_mode
    espm_b200.config.x_storage = keep


def _problem(seed, n, nx, ny, k, m, counts=25.0, big=False):
    rng = np.random.default_rng(seed)
    p = nx * ny
    x = np.linspace(0, 1, n)
    G = np.zeros((n, m))
    for j in range(m - 2):
        c, s = rng.uniform(0.05, 0.95), rng.uniform(0.01, 0.04)
        G[:, j] = np.exp(-0.5 * ((x - c) / s) ** 2)
    G[:, m - 2] = np.exp(-3 * x) + 0.05
    G[:, m - 1] = (1 - x) * 0.5 + 0.05
    Wt = rng.uniform(size=(m, k))
    Ht = rng.uniform(size=(k, p)) ** 2
    Ht /= Ht.sum(0, keepdims=True)
    lam = G @ Wt @ Ht
    X = rng.poisson(lam / lam.sum(0, keepdims=True) * counts * n / 50).astype(np.float32)
    X += 1.0 * (X.sum(1, keepdims=True) == 0)          # no all-zero channel (that would need remove_zeros_lines)
    if big:
        X[::7, ::5] *= 300.0                            # entries above 255: uint16
    W0 = rng.uniform(0.05, 1.0, size=(m, k)).astype(np.float32)
    H0 = rng.uniform(0.05, 1.0, size=(k, p))
    H0 = (H0 / H0.sum(0, keepdims=True)).astype(np.float32)
    return X, G.astype(np.float32), W0, H0


KW = dict(simplex_H=True, simplex_W=False, lambda_L=1.5, mu=0.03, tol=0, no_stop_criterion=True, max_iter=10, verbose=0)


@pytest.mark.parametrize("big,expect", [(False, "uint8"), (True, "uint16")])
@pytest.mark.parametrize("layout", ["np", "hspy"])
def test_compact_fit_equals_dense_fit(storage, big, expect, layout):
    from espm_b200 import SmoothNMF
    from oracle import smooth_nmf_oracle as orc
    nx, ny, n, k, m = 21, 30, 300, 3, 7
    X, G, W0, H0 = _problem(3, n, nx, ny, k, m, big=big)
    Xin = np.ascontiguousarray(X.T) if layout == "hspy" else X
    out = {}
    for mode in ("dense", "auto"):
        storage(mode)
        est = SmoothNMF(n_components=k, G=G, shape_2d=(nx, ny), hspy_comp=(layout == "hspy"), **KW)
        est.fit_transform(Xin, W=W0.copy(), H=H0.copy())
        assert est.x_storage_ == ("dense" if mode == "dense" else expect)
        out[mode] = est
    d, c = out["dense"], out["auto"]
    # same arithmetic on the same values: agreement far below the fp32 tolerance (bit-identical in practice)
    assert rel_err(c.losses_, d.losses_) < 1e-7
    assert rel_err(c.W_, d.W_) < 1e-6 and rel_err(c.H_, d.H_) < 1e-6
    assert c.const_KL_ == pytest.approx(d.const_KL_, rel=1e-12)
    # and the fp32 tolerances against the fp64 oracle on the same inputs
    kw = {a: b for a, b in KW.items() if a != "verbose"}
    ref = orc.fit(X.astype(np.float64), G.astype(np.float64), W0.astype(np.float64), H0.astype(np.float64),
                  shape_2d=(nx, ny), **kw)
    assert rel_err(c.losses_, ref["losses"]) < 1e-4
    assert rel_err(c.W_, ref["W"]) < 1e-3


def test_compact_single_steps_vs_oracle(storage):
    """Operator-level API on count data: the H and W steps and the loss through the uint8 kernels."""
    from espm_b200 import ops
    from oracle import smooth_nmf_oracle as orc
    storage("uint8")
    nx, ny, n, k, m = 16, 16, 517, 5, 9
    X, G, W0, H0 = _problem(11, n, nx, ny, k, m)
    a64 = [a.astype(np.float64) for a in (X, G, W0, H0)]
    Lg = ops.create_laplacian_matrix(nx, ny)
    ref_plain = orc.multiplicative_step_h(*a64, simplex_H=False, mu=0.05, lambda_L=2.0, shape_2d=(nx, ny))
    h_plain = ops.multiplicative_step_h(X, G, W0, H0, simplex_H=False, mu=0.05, lambda_L=2.0, L=Lg)
    assert rel_err(h_plain, ref_plain) < 1e-5
    ref_w = orc.multiplicative_step_w(a64[0], a64[1], a64[2], ref_plain, simplex_W=False)
    w = ops.multiplicative_step_w(X, G, W0, ref_plain.astype(np.float32), simplex_W=False)
    assert rel_err(w, ref_w) < 1e-5
    val, _ = ops.full_loss(X, G, W0, H0, mu=0.05, lambda_L=2.0, shape_2d=(nx, ny), const=orc.const_KL(a64[0]))
    ref_val, _ = orc.full_loss(*a64, mu=0.05, lambda_L=2.0, shape_2d=(nx, ny))
    assert rel_err(val, ref_val) < 1e-5
    h, its = ops.multiplicative_step_h(X, G, W0, H0, simplex_H=True, mu=0.05, lambda_L=2.0, L=Lg, return_its=True)
    ref_h, its_ref = orc.multiplicative_step_h(*a64, simplex_H=True, mu=0.05, lambda_L=2.0, shape_2d=(nx, ny),
                                               return_its=True)
    assert its == its_ref
    rel = np.abs(h - ref_h) / ref_h
    assert np.quantile(rel, 0.99) < 1e-5 and rel.max() < 2.5e-5


def test_compact_free_nmf_identity_G(storage):
    """G=None, simplex_W (BASELINE config C5) on count data."""
    from espm_b200 import SmoothNMF
    nx, ny, n, k = 12, 20, 96, 4
    X, _, _, H0 = _problem(5, n, nx, ny, k, 6)
    W0 = np.random.default_rng(1).uniform(0.05, 1, size=(n, k)).astype(np.float32)
    res = {}
    for mode in ("dense", "auto"):
        storage(mode)
        est = SmoothNMF(n_components=k, G=None, shape_2d=(nx, ny), simplex_H=False, simplex_W=True, tol=0,
                        no_stop_criterion=True, max_iter=8, verbose=0)
        est.fit_transform(X, W=W0.copy(), H=H0.copy())
        res[mode] = est
    assert res["auto"].x_storage_ == "uint8"
    assert rel_err(res["auto"].losses_, res["dense"].losses_) < 1e-7
    assert rel_err(res["auto"].W_, res["dense"].W_) < 1e-6


@pytest.mark.parametrize("why", ["fraction", "zero_channel", "zero_pixel", "normalize", "fp64", "too_large", "negative"])
def test_compact_storage_is_refused_when_it_would_change_the_data(storage, why):
    """The pre-scan keeps dense storage whenever uint8 / uint16 could not hold the processed X exactly."""
    from espm_b200 import SmoothNMF
    storage("auto")
    nx, ny, n, k, m = 9, 14, 120, 3, 6
    X, G, W0, H0 = _problem(9, n, nx, ny, k, m)
    kw = dict(KW, max_iter=3)
    if why == "fraction":
        X[5, 7] = 2.5
    elif why == "zero_channel":
        X[11, :] = 0.0
    elif why == "zero_pixel":
        X[:, 40] = 0.0
    elif why == "normalize":
        kw["normalize"] = True
    elif why == "fp64":
        X, G = X.astype(np.float64), G.astype(np.float64)
    elif why == "too_large":
        X[3, 3] = 70000.0
    elif why == "negative":
        X[3, 3] = -1.0
    est = SmoothNMF(n_components=k, G=G, shape_2d=(nx, ny), **kw)
    if why == "negative":
        with pytest.raises(ValueError, match="Negative values in data"):      # base.py:528
            est.fit_transform(X, W=W0.copy(), H=H0.copy())
        return
    est.fit_transform(X, W=W0.copy(), H=H0.copy())
    assert est.x_storage_ == "dense"
    assert np.all(np.isfinite(est.losses_))


def test_device_initialisation_on_compact_storage(storage):
    """NNDSVD on the device (init_device.py) reads the uint8 Xt through a dense copy: same factors as with dense
    storage."""
    from espm_b200 import SmoothNMF
    nx, ny, n, k, m = 40, 40, 64, 3, 6
    X, G, _, _ = _problem(21, n, nx, ny, k, m)
    res = {}
    for mode in ("dense", "auto"):
        storage(mode)
        est = SmoothNMF(n_components=k, G=G, shape_2d=(nx, ny), simplex_H=True, simplex_W=False, tol=0,
                        no_stop_criterion=True, max_iter=4, verbose=0, random_state=0)
        est.fit_transform(X)
        res[mode] = est
    assert res["auto"].x_storage_ == "uint8"
    assert rel_err(res["auto"].losses_, res["dense"].losses_) < 1e-5
