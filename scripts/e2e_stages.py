"""Stage-by-stage wall time of SmoothNMF.fit_transform(X in pinned host memory) at C3 (20 iterations): which part of the
end-to-end figure is the PCIe copy, which is set-up, which is the loop.  Every stage is bracketed by a device
synchronisation, so the sum is an upper bound of the un-instrumented fit (printed beside it)."""
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
from espm_b200 import engine as E

nx = ny = 512
n, k, K = 2048, 4, 20
prob = synth.make_problem(nx, ny, n, k, 25, seed=93)
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
_print = print


def print(*a, **k_):                      # noqa: A001 -- rank 0 speaks
    if rank == 0:
        _print(*a, **k_)


X = synth.poisson_X_torch(prob, 0, nx * ny, 93, dev, torch.float32)
Xh = torch.empty((n, nx * ny), dtype=torch.float32, pin_memory=True)
Xh.copy_(X)
del X
torch.cuda.synchronize()
W0, H0 = synth.init_factors(prob["G_full"].shape[1], k, nx * ny, 93, dtype=np.float32)
G = prob["G_full"].astype(np.float32)
kw = dict(simplex_H=True, simplex_W=False, lambda_L=2.0, mu=0.05)
espm_b200.config.x_storage = os.environ.get("X_STORAGE", "dense")

# raw PCIe rate
for _ in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = Xh.to(dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print("plain H2D of X (%.2f GB): %.1f ms = %.1f GB/s" % (Xh.nbytes / 1e9, dt * 1e3, Xh.nbytes / dt / 1e9))
del d

stages = {}


def timed(cls, name):
    fn = getattr(cls, name)

    def wrap(self, *a, **kw_):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(self, *a, **kw_)
        torch.cuda.synchronize()
        stages[name] = stages.get(name, 0.0) + time.perf_counter() - t0
        return out
    setattr(cls, name, wrap)
    return fn


def fit():
    est = SmoothNMF(n_components=k, G=G, shape_2d=(nx, ny), tol=0.0, no_stop_criterion=True, max_iter=K, verbose=0, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        est.fit_transform(Xh.numpy(), W=W0.copy(), H=H0.copy())
    torch.cuda.synchronize()
    return time.perf_counter() - t0


fit()
plain = sorted(fit() for _ in range(5))
print("un-instrumented fits: %s ms" % ", ".join("%.1f" % (t * 1e3) for t in plain))
orig = {}
if world > 1:
    from espm_b200 import dist as D
    for nm in ("setup_peer", "setup_inbox", "close"):
        timed(D.PeerShard, nm)
    mk0 = D.make_shard

    def mk(*a, **kw_):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = mk0(*a, **kw_)
        stages["make_shard"] = stages.get("make_shard", 0.0) + time.perf_counter() - t0
        return out
    D.make_shard = mk
for nm in ("_stage_x", "_choose_storage", "_retile_x", "set_G", "_init_WH", "run_iterations", "evaluate", "read_records",
           "get_W", "get_H"):
    orig[nm] = timed(E.FitEngine, nm)
init0 = E.FitEngine.__init__


def init(self, *a, **kw_):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    init0(self, *a, **kw_)
    torch.cuda.synchronize()
    stages["__init__ (all)"] = stages.get("__init__ (all)", 0.0) + time.perf_counter() - t0


E.FitEngine.__init__ = init
R = 3
tot = sum(fit() for _ in range(R))
print("instrumented fit: %.1f ms" % (tot / R * 1e3))
for nm, t in sorted(stages.items(), key=lambda kv: -kv[1]):
    print("  %-18s %7.2f ms" % (nm, t / R * 1e3))
sub = sum(stages.get(nm, 0.0) for nm in ("_stage_x", "_choose_storage", "_retile_x", "set_G", "_init_WH", "setup_peer",
                                         "setup_inbox"))
print("  %-18s %7.2f ms" % ("__init__ (rest)", (stages["__init__ (all)"] - sub) / R * 1e3))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
