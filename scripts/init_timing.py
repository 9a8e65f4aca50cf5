"""Wall time of a default SmoothNMF.fit_transform(X) (no W, no H: NNDSVD initialisation) at C3 size with the
randomized SVD on the device vs on the host (espm_b200.config.device_init)."""
import contextlib
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import espm_b200
from espm_b200 import SmoothNMF, synth

nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n, k, K = 2048, 4, 50
prob = synth.make_problem(nx, ny, n, k, 25, seed=93)
dev = torch.device("cuda", 0)
X = synth.poisson_X_torch(prob, 0, nx * ny, 93, dev, torch.float32)
Xh = torch.empty((n, nx * ny), dtype=torch.float32, pin_memory=True)
Xh.copy_(X)
del X
torch.cuda.synchronize()
G = prob["G_full"].astype(np.float32)
kw = dict(simplex_H=True, simplex_W=False, lambda_L=2.0, mu=0.05)
res = {}
for flag, reps in ((True, 3), (False, 1)):
    espm_b200.config.device_init = flag
    for rep in range(reps):
        est = SmoothNMF(n_components=k, G=G, shape_2d=(nx, ny), tol=0.0, no_stop_criterion=True, max_iter=K, verbose=0,
                        random_state=7, **kw)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            est.fit_transform(Xh.numpy())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("device_init=%s rep %d: fit_transform(X) %d iterations in %.3f s, final loss %.8f" % (
            flag, rep, K, dt, est.losses_[-1]), flush=True)
        res[flag] = np.array(est.losses_)
print("max rel diff of the loss histories:", np.max(np.abs(res[True] - res[False]) / np.abs(res[False])))
