"""Where does the wall time of SmoothNMF.fit_transform(host X) go?  (cProfile with CUDA_LAUNCH_BLOCKING=1)"""
import cProfile
import contextlib
import io
import os
import pstats
import sys
import time

os.environ.setdefault("CUDA_LAUNCH_BLOCKING", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

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
W0, H0 = synth.init_factors(prob["G_full"].shape[1], k, nx * ny, 93, dtype=np.float32)
G = prob["G_full"].astype(np.float32)
kw = dict(simplex_H=True, simplex_W=False, lambda_L=2.0, mu=0.05)
for rep in range(2):
    est = SmoothNMF(n_components=k, G=G, shape_2d=(nx, ny), tol=0.0, no_stop_criterion=True, max_iter=K, verbose=0, **kw)
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        pr.enable()
        est.fit_transform(Xh.numpy(), W=W0.copy(), H=H0.copy())
        pr.disable()
    torch.cuda.synchronize()
    print("rep %d wall %.3f s" % (rep, time.perf_counter() - t0))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue())
