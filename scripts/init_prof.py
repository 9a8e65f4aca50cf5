import sys, time, contextlib, io
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from espm_b200 import synth
from espm_b200.engine import FitEngine
from espm_b200 import init_device as I
nx=ny=512; n,k=2048,4
prob = synth.make_problem(nx, ny, n, k, 25, seed=93)
dev = torch.device("cuda", 0)
X = synth.poisson_X_torch(prob, 0, nx*ny, 93, dev, torch.float32)
G = prob["G_full"].astype(np.float32)
eng = FitEngine(X, G, np.ones((G.shape[1],k),np.float32), np.ones((k,nx*ny),np.float32), max_records=16, x_local=True, shape_2d=(nx,ny), ingest=dict(eps=1e-14, normalize=None))
DX = I._DeviceX(eng)
def T(f, *a, reps=3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): out=f(*a)
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps, out
Qn = torch.randn(n, 14, device=dev)
for lib in ("default","cusolver","magma"):
    torch.backends.cuda.preferred_linalg_library(lib)
    t1,Y = T(DX.xt_times, Qn)
    t2,PL = T(I._lu_permute_l_device, Y)
    t3,Z = T(DX.x_times, PL)
    t4,_ = T(lambda y: torch.linalg.qr(y, mode="reduced"), Y)
    print(lib, "xt_times %.4f lu %.4f x_times %.4f qr %.4f"%(t1,t2,t3,t4), flush=True)
t,_=T(lambda: I.initialize_nmf_device(eng, k, None, 7), reps=2); print("full init %.3f s"%t)
