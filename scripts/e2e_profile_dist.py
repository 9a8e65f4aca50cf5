"""cProfile of a pixel-sharded SmoothNMF.fit_transform (torchrun, rank 0 prints): where does the per-fit set-up go?"""
import cProfile
import contextlib
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import espm_b200
from espm_b200 import SmoothNMF, synth

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
nx = ny = 512
n, k, K = 2048, 4, 50
prob = synth.make_problem(nx, ny, n, k, 25, seed=93)
X = synth.poisson_X_torch(prob, 0, nx * ny, 93, dev, torch.float32)
Xh = torch.empty((n, nx * ny), dtype=torch.float32, pin_memory=True)
Xh.copy_(X)
del X
torch.cuda.synchronize()
W0, H0 = synth.init_factors(prob["G_full"].shape[1], k, nx * ny, 93, dtype=np.float32)
G = prob["G_full"].astype(np.float32)
espm_b200.config.distributed = True
kw = dict(simplex_H=True, simplex_W=False, lambda_L=2.0, mu=0.05)
for rep in range(3):
    est = SmoothNMF(n_components=k, G=G, shape_2d=(nx, ny), tol=0.0, no_stop_criterion=True, max_iter=K, verbose=0, **kw)
    pr = cProfile.Profile()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        pr.enable()
        est.fit_transform(Xh.numpy(), W=W0.copy(), H=H0.copy())
        pr.disable()
    torch.cuda.synchronize()
    if rank == 0:
        print("rep %d wall %.3f s" % (rep, time.perf_counter() - t0))
if rank == 0:
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(40)
    print(s.getvalue())
dist.destroy_process_group()
