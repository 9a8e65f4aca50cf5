"""Diagnostic: wall time of repeated SmoothNMF.fit_transform(host X) calls and allocator state (bench.py's e2e leg)."""
import contextlib
import gc
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from espm_b200 import SmoothNMF, synth

nx = ny = 512
n, k, K = 2048, 4, 50
prob = synth.make_problem(nx, ny, n, k, 25, seed=93)
dev = torch.device("cuda", 0)
X = synth.poisson_X_torch(prob, 0, nx * ny, 93, dev, torch.float32)
Xh = torch.empty((n, nx * ny), dtype=torch.float32, pin_memory=True)
Xh.copy_(X)
del X
torch.cuda.synchronize()
torch.cuda.empty_cache()
W0, H0 = synth.init_factors(prob["G_full"].shape[1], k, nx * ny, 93, dtype=np.float32)
G = prob["G_full"].astype(np.float32)
kw = dict(simplex_H=True, simplex_W=False, lambda_L=2.0, mu=0.05)


def fit(max_iter, keep):
    est = SmoothNMF(n_components=k, G=G, shape_2d=(nx, ny), tol=0.0, no_stop_criterion=True, max_iter=max_iter,
                    verbose=0, **kw)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        est.fit_transform(Xh.numpy(), W=W0.copy(), H=H0.copy())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("max_iter %3d wall %.3f s   reserved %.2f GB allocated %.2f GB  gc objects %d" % (
        max_iter, dt, torch.cuda.memory_reserved() / 1e9, torch.cuda.memory_allocated() / 1e9, len(gc.get_objects())))
    return est if keep else None


for mi, keep in [(5, False), (50, True), (50, False), (5, False), (50, False), (50, True)]:
    e = fit(mi, keep)
print("gc.collect ->", gc.collect())
fit(50, False)
