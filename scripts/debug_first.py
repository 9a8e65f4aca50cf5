"""Scratch: first contact with the GPU -- run each kernel stage on a small problem and compare."""
import sys, os, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(120, exit=True)
import numpy as np, torch
from espm_b200 import ops
from espm_b200.engine import FitEngine
from espm_b200 import _lib as L
from oracle import smooth_nmf_oracle as orc

rng = np.random.default_rng(0)
def prob(n, nx, ny, k, m):
    p = nx*ny
    x = np.linspace(0, 1, n)
    G = np.zeros((n, m))
    for j in range(m-2):
        G[:, j] = np.exp(-0.5*((x-rng.uniform(0.05,0.95))/0.03)**2)
    G[:, m-2] = np.exp(-3*x)+0.05; G[:, m-1] = (1-x)*0.5+0.05
    Wt = rng.uniform(size=(m,k)); Ht = rng.uniform(size=(k,p))**2; Ht /= Ht.sum(0, keepdims=True)
    lam = G@Wt@Ht
    X = rng.poisson(lam/lam.sum(0,keepdims=True)*20).astype(np.float64)
    W0 = rng.uniform(0.05,1,size=(m,k)); H0 = rng.uniform(0.05,1,size=(k,p)); H0/=H0.sum(0,keepdims=True)
    return X,G,W0,H0

def rel(a,b): return float(np.max(np.abs(a-b)/np.maximum(np.abs(b),1e-300)))

for (n,nx,ny,k,m) in [(64,5,7,3,5),(300,16,16,4,6),(1980,80,80,3,11)]:
    X,G,W0,H0 = prob(n,nx,ny,k,m)
    for dt in (np.float64, np.float32):
        Xd,Gd,Wd,Hd = [a.astype(dt) for a in (X,G,W0,H0)]
        eng = FitEngine(Xd,Gd,Wd,Hd, shape_2d=(nx,ny), lambda_L=2.0, mu=0.05, simplex_H=True, simplex_W=False, max_records=8)
        st = eng.st
        print("plan", dict(n_pad=st.n_pad,kp=st.kp,tiles=st.n_tiles,h_grid=st.h_grid,nsplit=st.h_nsplit,w_nb=st.w_nb,w_nr=st.w_nr,hd=st.h_depth,wd=st.w_depth,hs=st.h_smem,ws=st.w_smem,rows=st.w_sacc_rows))
        torch.cuda.synchronize()
        # check retile
        Xt = eng.Xt.view(st.n_tiles, st.n_pad, 128).cpu().numpy()
        Xr = np.zeros((st.n_tiles*128, st.n_pad)); Xr[:X.shape[1], :n] = Xd.T
        print(" retile ok", np.array_equal(Xt.transpose(0,2,1).reshape(-1, st.n_pad), Xr))
        GWd = eng.GWbuf[eng.iw[0]].cpu().numpy()[:n,:k]
        print(" GW err", rel(GWd, G@np.maximum(W0,1e-14)))
        eng.evaluate(0); torch.cuda.synchronize()
        GW = G@W0
        numraw_ref = GW.T@(X/(GW@H0))
        nr = eng.numraw.cpu().numpy().sum(0)[:k,:X.shape[1]]
        print(" numraw err", rel(nr, numraw_ref))
        rec = eng.read_records(0,1)[0]
        print(" xlogy", rec[L.S_XLOGY], np.sum(np.maximum(X,1e-14)*np.log(np.maximum(GW,1e-14)@H0)), "sumy", rec[L.S_SUMY], np.sum(GW@H0))
        print(" logreg", rec[L.S_LOGREG], orc.log_reg(H0,0.05,1), "lap", rec[L.S_LAPL], orc.trace_xtLx(H0,(nx,ny)), "its", rec[L.S_BISECT_ITS_H], "flags", rec[L.S_DEV_FLAGS])
        eng._call(eng.lib.espm_h_apply); torch.cuda.synchronize()
        Hn = eng.Hbuf[eng.ih[2]][:, eng.halo:eng.halo+eng.p_loc].cpu().numpy()
        Href, its = orc.multiplicative_step_h(X,G,W0,H0,simplex_H=True,mu=0.05,lambda_L=2.0,shape_2d=(nx,ny),return_its=True)
        print(" H err", rel(Hn, Href), "its ref", its)
        # W step
        eng2 = FitEngine(Xd,Gd,Wd,Href.astype(dt), simplex_H=False, simplex_W=False, max_records=8)
        Wn, rec = eng2.step_w_only()
        Wref = orc.multiplicative_step_w(X,G,W0,Href,simplex_W=False)
        print(" W err", rel(Wn, Wref))
print("DONE")
