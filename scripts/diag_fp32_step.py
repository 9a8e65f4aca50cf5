"""Where does the fp32 H step lose accuracy at C3 size?  Compares num / den / H' of the device (fp32) with an fp64
evaluation on a pixel subset and prints error quantiles (diagnostic for tests/test_gpu_fullsize.py)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from espm_b200 import synth, _lib as L
from espm_b200.engine import FitEngine
from oracle import smooth_nmf_oracle as orc

nx = ny = 512
n, k, n_el, seed = 2048, 4, 25, 93
lam, mu, eps, LS, SIGMA = 2.0, 0.05, 1.0, 1e-14, 8.0
prob = synth.make_problem(nx, ny, n, k, n_el, seed=seed)
dev = torch.device("cuda", 0)
X = synth.poisson_X_torch(prob, 0, nx * ny, seed, dev, torch.float32)
G = prob["G_full"].astype(np.float32)
W0, H0 = synth.init_factors(G.shape[1], k, nx * ny, seed, dtype=np.float32)
p = nx * ny
eng = FitEngine(X, G, W0, H0, shape_2d=(nx, ny), lambda_L=lam, mu=mu, epsilon_reg=eps, simplex_H=True,
                simplex_W=False, tol=0.0, max_records=16, x_local=True)
eng.evaluate(0)
torch.cuda.synchronize()
numd = eng.num[:k, :p].double().cpu().numpy()
dend = eng.den[:k, :p].double().cpu().numpy()
eng.advance(1)
eng.evaluate(1)
recs = eng.read_records(0, 2)
H1 = eng.get_H().astype(np.float64)
its = int(recs[0][L.S_BISECT_ITS_H])
G64, W64, H64 = G.astype(np.float64), np.maximum(W0.astype(np.float64), LS), np.maximum(H0.astype(np.float64), LS)
rng = np.random.default_rng(7)
J = np.sort(rng.choice(p, size=4096, replace=False))
XJ = X[:, torch.as_tensor(J, device=dev)].double().cpu().numpy()
GW = G64 @ W64
HL = orc.laplacian_apply(H64, (nx, ny))[:, J]
HJ = H64[:, J]
ratio = GW.T @ (XJ / (GW @ HJ))
den = np.sum(GW, axis=0, keepdims=True).T + mu / (HJ + eps)
maxH = np.max(H64, axis=1, keepdims=True)
num = HJ * (ratio + lam * SIGMA * maxH)
den = den + lam * SIGMA * maxH + lam * HL


def q(e):
    e = np.abs(e).ravel()
    return "max %.2e  p99.9 %.2e  p99 %.2e  median %.2e" % (e.max(), np.quantile(e, 0.999), np.quantile(e, 0.99), np.median(e))


print("its", its)
print("num rel err:", q((numd[:, J] - num) / num))
print("den rel err:", q((dend[:, J] - den) / den))
# ratio part alone: numraw = sum_c GW x / y
nraw = eng.numraw[0, :k, :p].double().cpu().numpy() if eng.st.h_nsplit == 1 else None
if nraw is not None:
    print("ratio-sum rel err:", q((nraw[:, J] - ratio) / ratio))


def replay(num, den, its):
    kk = num.shape[0]
    a = np.max(np.where(num > 0, num / 2 - den, -np.inf), axis=0)
    b = kk * np.max(num, axis=0) / 0.5 - np.min(den, axis=0)
    f = lambda x: np.sum(np.maximum(num / (x + den), LS), axis=0) - 1
    new = (a + b) / 2
    fn = f(new)
    for it in range(its):
        minus = f(a) * fn <= 0
        b[minus] = new[minus]
        a[~minus] = new[~minus]
        new = (a + b) / 2
        fn = f(new)
    return new


nu = replay(num.copy(), den.copy(), its)
ref = np.maximum(num / (den + nu), LS)
err = np.abs(H1[:, J] - ref) / ref
print("H' rel err:", q(err))
nu_dev = replay(numd[:, J].copy(), dend[:, J].copy(), its)
ref_dev = np.maximum(numd[:, J] / (dend[:, J] + nu_dev), LS)
print("H' vs fp64 update of the DEVICE's num/den:", q((H1[:, J] - ref_dev) / ref_dev))
print("nu rel diff (device num/den vs fp64 num/den):", q((nu_dev - nu) / (np.abs(nu) + den.min(0))))
w = np.unravel_index(np.argmax(err), err.shape)
print("worst entry: phase %d pixel %d  H'=%.6g ref=%.6g  H=%.4g num=%.6g/%.6g den=%.6g/%.6g nu=%.6g/%.6g" % (
    w[0], J[w[1]], H1[w[0], J[w[1]]], ref[w], HJ[w], numd[w[0], J[w[1]]], num[w], dend[w[0], J[w[1]]], den[w],
    nu_dev[w[1]], nu[w[1]]))
print("sum of column of worst:", H1[:, J[w[1]]].sum(), ref[:, w[1]].sum())
