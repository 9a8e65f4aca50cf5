"""Speculative replay of the lock-step bisection (simplex_H): espm_h_finish applies the count of the previous H update
right after its trace, espm_h_apply only confirms it.  The count is a hint, never an input: with the
speculation switched off (espm_b200.config.speculate_h = False -> ESPM_FLAG_NO_HSPEC) every bit of the fit must be the
same -- losses, W, H, rel_W / rel_H and the recorded counts -- for every update rule that bisects
(dicotomy.py:4-55, 57-81, 83-108), in both loop variants, and also while the count is still changing (first iterations)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture
def spec_switch():
    import espm_b200
    keep = espm_b200.config.speculate_h

    def set_mode(on):
        espm_b200.config.speculate_h = on
    yield set_mode
    espm_b200.config.speculate_h = keep


def _problem(seed, n, nx, ny, k, m, dtype):
    rng = np.random.default_rng(seed)
    p = nx * ny
    x = np.linspace(0, 1, n)
    G = np.zeros((n, m))
    for j in range(m - 1):
        c, s = rng.uniform(0.05, 0.95), rng.uniform(0.01, 0.05)
        G[:, j] = np.exp(-0.5 * ((x - c) / s) ** 2)
    G[:, m - 1] = np.exp(-3 * x) + 0.05
    Wt = rng.uniform(size=(m, k))
    Ht = rng.uniform(size=(k, p)) ** 2
    Ht /= Ht.sum(0, keepdims=True)
    lam = G @ Wt @ Ht
    X = rng.poisson(lam / lam.mean() * 0.4).astype(dtype)
    X += 1.0 * (X.sum(1, keepdims=True) == 0)
    W0 = rng.uniform(0.05, 1.0, size=(m, k)).astype(dtype)
    H0 = rng.uniform(0.05, 1.0, size=(k, p))
    H0 = (H0 / H0.sum(0, keepdims=True)).astype(dtype)
    return X, G.astype(dtype), W0, H0


CASES = {
    "kl": dict(lambda_L=1.5, mu=0.03),
    "kl_plain": dict(),
    "hq": dict(lambda_L=1.0, algo="l2_surrogate"),
    "pg": dict(algo="projected_gradient", lambda_L=0.5),
    "fixed_h": dict(lambda_L=0.7, fixed="H"),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("checked", [False, True])
def test_speculative_replay_changes_nothing(spec_switch, case, dtype, checked):
    from espm_b200 import SmoothNMF
    nx, ny, n, k, m = 23, 19, 200, 3, 6
    X, G, W0, H0 = _problem(11, n, nx, ny, k, m, dtype)
    kw = dict(CASES[case])
    if kw.pop("fixed", None) == "H":
        fH = -np.ones_like(H0)
        fH[0, ::7] = 0.25
        kw["fixed_H"] = fH
    kw.update(simplex_H=True, simplex_W=False, shape_2d=(nx, ny), max_iter=14)
    if checked:
        kw.update(tol=1e-9, verbose=1)            # the reference's loop with stop tests (one iteration ahead + rollback)
    else:
        kw.update(tol=0, no_stop_criterion=True, verbose=0)
    out = {}
    for on in (False, True):
        spec_switch(on)
        est = SmoothNMF(n_components=k, G=G, **kw)
        est.fit_transform(X, W=W0.copy(), H=H0.copy())
        out[on] = est
    a, b = out[False], out[True]
    assert a.n_iter_ == b.n_iter_
    assert np.array_equal(np.asarray(a.losses_), np.asarray(b.losses_))
    assert np.array_equal(np.asarray(a.rel_), np.asarray(b.rel_))
    assert np.array_equal(a.W_, b.W_) and np.array_equal(a.H_, b.H_)


def test_speculation_hits_in_steady_state(spec_switch):
    """The point of the exercise: after the first H updates espm_h_apply finds the count already applied
    (dev_flags[7] == count + 1) and has nothing to redo."""
    from espm_b200.engine import FitEngine
    from espm_b200 import _lib as L
    nx, ny, n, k, m = 23, 19, 200, 3, 6
    X, G, W0, H0 = _problem(5, n, nx, ny, k, m, np.float32)
    spec_switch(True)
    eng = FitEngine(X, G, W0, H0, shape_2d=(nx, ny), max_records=40, simplex_H=True, simplex_W=False, lambda_L=1.0,
                    mu=0.02, tol=0.0)
    eng.evaluate(0)
    hits = []
    for it in range(1, 25):
        applied = int(eng.dev_flags[7].item())
        eng.advance(it)
        count = int(eng.dev_flags[6].item())            # it* + 1 of the h_apply that just ran
        hits.append(applied == count)
        eng.evaluate(it)
    rec = eng.read_records(0, 25)
    assert all(int(r[L.S_DEV_FLAGS]) == 0 for r in rec)
    assert not hits[0]                                   # nothing to speculate with on the first update
    assert sum(hits[5:]) >= 10, hits                     # the count settles: most later updates are hits
    eng.close()
