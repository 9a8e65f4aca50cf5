"""Synthetic STEM-EDXS spectrum images of the BASELINE.json shapes (no network, no hyperspy/exspy):
an EDXS G (tabulated x-ray lines at 200 keV + two bremsstrahlung columns), sphere phase maps and Poisson counts,
following the recipe of SURVEY.md section 8d (espm/datasets/base.py:13-68, models/edxs.py:163-254,
weights/generate_weights.py:182-223).  Host-side NumPy for small cases, torch for device generation."""
import numpy as np


_TABLES = None


def _tables():
    """x-ray line energies / cross sections at 200 keV and the SDD efficiency curve of the reference's tables
    (espm/tables/200keV_xrays.json, SDD_efficiency.txt), extracted by scripts/gen_xray_table.py."""
    global _TABLES
    if _TABLES is None:
        import json
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "edxs_tables.json")
        with open(path) as fh:
            _TABLES = json.load(fh)
    return _TABLES


def edxs_G(n, n_elements, seed, e_offset=0.2, e_scale=0.01, E0=200.0):
    """G (n x (n_elements+2)) following SURVEY.md section 8d: energy axis x = linspace(offset, offset + n scale, n)
    (models/base.py:215); per element the column sum_lines cs N(x; E_line, (0.01 E_line + 0.065) / 2.3548) D(E_line)
    with the line energies and emission cross sections of espm/tables/200keV_xrays.json and the detector efficiency D
    of SDD_efficiency.txt (models/edxs.py:56-110, EDXS_function.py:20-46; absorption == 1); plus the two
    bremsstrahlung columns lifshin_b0 / lifshin_b1 (EDXS_function.py:159-172) times D(x); the element columns are
    normalised by their mean sum, the two continuum columns by their own sums (models/edxs.py:246-252).
    ``seed`` is unused (the element list is fixed: the first n_elements of the table's 25)."""
    tb = _tables()
    x = np.linspace(e_offset, e_offset + n * e_scale, n)
    ee = np.asarray(tb["sdd_efficiency"]["energy_keV"])
    ev = np.asarray(tb["sdd_efficiency"]["efficiency"])

    def det(e):
        return np.interp(e, ee, ev)
    Z = tb["elements"][:n_elements]
    if len(Z) < n_elements:
        raise ValueError("the table holds %d elements, %d requested" % (len(tb["elements"]), n_elements))
    G = np.zeros((n, n_elements + 2))
    for e, z in enumerate(Z):
        for _, E, cs in tb["lines"][str(z)]:
            if not (x[0] <= E <= x[-1]):
                continue
            w = (0.01 * E + 0.065) / 2.3548
            G[:, e] += cs * np.exp(-0.5 * ((x - E) / w) ** 2) / (w * np.sqrt(2 * np.pi)) * det(E)
    G[:, -2] = (E0 - x) / (E0 * x) * (1.0 - (E0 - x) / E0) * det(x)
    G[:, -1] = (E0 - x) ** 2 / (E0 * E0 * x) * det(x)
    norms = G.sum(axis=0, keepdims=True)
    norms[0, :-2] = np.mean(norms[0, :-2])
    return G / norms


def true_W(m, k, seed):
    rng = np.random.default_rng(seed + 1)
    W = np.zeros((m, k))
    for c in range(k):
        idx = rng.choice(m - 2, size=int(rng.integers(3, 7)), replace=False)
        W[idx, c] = rng.uniform(size=idx.size)
        W[:m - 2, c] /= W[:m - 2, c].sum()
        W[m - 2, c] = 1e-5 * rng.uniform(0.5, 1.0)
        W[m - 1, c] = 1e-3 * rng.uniform(0.5, 1.0)
    return W


def sphere_maps(nx, ny, k, seed):
    """H_true (k x p): k-1 spheres of radius ~nx/4 with concentration 1/(k-1); phase 0 is the complement."""
    rng = np.random.default_rng(seed + 2)
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    H = np.zeros((k, nx, ny))
    r = max(min(nx, ny) / 4.0, 1.0)
    for c in range(1, k):
        ci, cj = rng.uniform(0, nx), rng.uniform(0, ny)
        d2 = (ii - ci) ** 2 + (jj - cj) ** 2
        H[c] = np.clip(1.0 - d2 / r ** 2, 0.0, None) ** 0.5 / max(k - 1, 1)
    H[0] = 1.0 - H[1:].sum(axis=0)
    return H.reshape(k, nx * ny)


def init_factors(m, k, p, seed, dtype=np.float64):
    rng = np.random.default_rng(seed + 3)
    W0 = rng.uniform(size=(m, k)) + 1e-3
    H0 = rng.uniform(size=(k, p)) + 1e-3
    H0 /= H0.sum(axis=0, keepdims=True)
    return W0.astype(dtype), H0.astype(dtype)


def make_problem(nx, ny, n, k, n_elements, seed=91, counts=500.0, identity_G=False):
    """Returns dict(G, W_true, H_true, dens, D) -- everything but the (large) X."""
    G = edxs_G(n, n_elements, seed, e_scale=0.01 if n <= 2048 else 0.005)
    m = G.shape[1]
    Wt = true_W(m, k, seed)
    D = G @ Wt
    D = D / D.sum(axis=0, keepdims=True)
    Ht = sphere_maps(nx, ny, k, seed)
    rng = np.random.default_rng(seed + 4)
    dens = rng.uniform(0.6, 1.0, size=k)
    return dict(G=None if identity_G else G, G_full=G, W_true=Wt, H_true=Ht, D=D, dens=dens, counts=counts,
                m=(n if identity_G else m))


def poisson_X_numpy(prob, j0, j1, seed, dtype=np.float32):
    """Host generation of X[:, j0:j1] ~ Poisson(N * (D * dens) @ H_true)."""
    rng = np.random.default_rng([seed, j0])
    lam = prob["counts"] * (prob["D"] * prob["dens"][None, :]) @ prob["H_true"][:, j0:j1]
    return rng.poisson(lam).astype(dtype)


def poisson_X_torch(prob, j0, j1, seed, device, dtype, chunk=32768):
    """Device generation of the same distribution (different random stream), (n, j1-j0) tensor.

    The random stream is seeded per GLOBAL chunk of `chunk` pixels, and a chunk is always drawn whole, so the
    pixels [j0, j1) come out identical however the image is partitioned over ranks: 1-, 2-, 4- and 8-GPU runs of
    bench.py stream the same X and their losses / W can be compared with each other."""
    import torch
    Dd = torch.as_tensor(prob["counts"] * prob["D"] * prob["dens"][None, :], dtype=torch.float32, device=device)
    n = Dd.shape[0]
    p = prob["H_true"].shape[1]
    X = torch.empty(n, j1 - j0, dtype=dtype, device=device)
    gen = torch.Generator(device=device)
    for c in range(j0 // chunk, (j1 + chunk - 1) // chunk):
        a, b = c * chunk, min((c + 1) * chunk, p)
        gen.manual_seed(seed * 1000003 + c)
        Ht = torch.as_tensor(prob["H_true"][:, a:b], dtype=torch.float32, device=device)
        draw = torch.poisson(Dd @ Ht, generator=gen)
        lo, hi = max(a, j0), min(b, j1)
        X[:, lo - j0:hi - j0] = draw[:, lo - a:hi - a].to(dtype)
        del draw
    return X
