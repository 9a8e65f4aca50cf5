import os, sys, io, contextlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import smooth_nmf_oracle as orc
import espm_b200
from espm_b200 import SmoothNMF
from test_gpu_parity import _problem
rng = np.random.default_rng(6)
n, nx, ny, k, m = 120, 9, 8, 3, 5
X, G, W0, H0 = _problem(rng, n, nx, ny, k, m)
G[40, :] = 0.0
G[40, 1] = 0.7
fixed_W = -np.ones_like(W0)
fixed_W[1, :] = 0.0
kw = dict(simplex_H=True, simplex_W=False, shape_2d=(nx, ny), tol=0, no_stop_criterion=True, max_iter=6, fixed_W=fixed_W)
with np.errstate(all="ignore"):
    ref = orc.fit(X, G, W0, H0, **kw)
print("ref     ", ref["losses"])
for spec in (True, False):
    for native in (True, False):
        for verbose in (0, 1):
            espm_b200.config.speculate, espm_b200.config.native_loop = spec, native
            est = SmoothNMF(n_components=k, G=G, verbose=verbose, **kw)
            with contextlib.redirect_stdout(io.StringIO()):
                est.fit_transform(X, W=W0.copy(), H=H0.copy())
            print("spec=%d native=%d verbose=%d" % (spec, native, verbose), np.array(est.losses_))
