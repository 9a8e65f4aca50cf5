"""The table-driven log2 of the fp64 H pass (espm_b200/csrc/common.cuh::log2_tab): the sum max(X, ls) * log(Y) of
KLdiv_loss (measures.py:497-503) is formed with it in fp64 mode.

CPU part: a NumPy restatement of the routine (same table, same polynomial, long double standing in for the FMA)
against a long-double log2 -- pins the constants.  GPU part: the device routine through the C ABI
(espm_log2_table) against the same long-double reference, including the special values that take the library path.
"""
import ctypes

import numpy as np
import pytest

LD = np.longdouble
COEF = [float.fromhex(h) for h in ("0x1.71547652b82ffp+0", "-0x1.715476529026dp-1", "0x1.ec709dc2ea15bp-2",
                                   "-0x1.7155b049ac044p-2", "0x1.277837d2b64aap-2")]
TOL_ULP = 2.0    # |error| <= TOL_ULP ulps of the result + TOL_ABS (polynomial truncation, matters only near y = 1)
TOL_ABS = 1e-16


def sample(seed=0, n=400000):
    rng = np.random.default_rng(seed)
    y = np.exp(rng.uniform(np.log(1e-300), np.log(1e300), n))
    y = np.concatenate([y, np.exp(rng.uniform(np.log(1e-14), np.log(1e4), n)), 1 + rng.uniform(-1e-2, 1e-2, n // 4),
                        1 + rng.uniform(-1e-9, 1e-9, 1000),
                        np.array([1.0, 2.0, 0.5, 1 - 2.0 ** -53, 1 + 2.0 ** -52, 1e-14, 2.2250738585072014e-308])])
    return y


def log2_tab_numpy(y):
    i = np.arange(128)
    inv = 1.0 / (1.0 + (i + 0.5) / 128.0)
    l2 = (-np.log2(inv.astype(LD))).astype(np.float64)
    bits = y.view(np.int64)
    e = ((bits >> 52) & 0x7ff) - 1023
    idx = (bits >> 45) & 127
    m = ((bits & ((1 << 52) - 1)) | (1023 << 52)).view(np.float64)
    r = (m.astype(LD) * inv[idx].astype(LD) - 1).astype(np.float64)
    p = np.full_like(r, COEF[4])
    for c in COEF[3::-1]:
        p = p * r + c
    return p * r + (l2[idx] + e)


def ulp_err(res, y):
    ref = np.log2(y.astype(LD))
    bound = TOL_ULP * np.spacing(np.abs(ref.astype(np.float64))) + TOL_ABS
    return np.max(np.abs(res.astype(LD) - ref).astype(np.float64) / bound)


def test_log2_table_restatement():
    y = sample()
    assert ulp_err(log2_tab_numpy(y), y) <= 1.0


@pytest.mark.gpu
def test_log2_table_device():
    import torch
    from espm_b200 import _lib as L
    lib = L.load()
    y = np.concatenate([sample(1), np.array([5e-324, 1e-310, 0.0, np.inf])])
    d = torch.from_numpy(y).cuda()
    out = torch.empty_like(d)
    L.check(lib.espm_log2_table(ctypes.c_void_p(d.data_ptr()), d.numel(), ctypes.c_void_p(out.data_ptr()),
                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    res = out.cpu().numpy()
    assert res[-2] == -np.inf and res[-1] == np.inf
    assert ulp_err(res[:-2], y[:-2]) <= 1.5      # the device fills its table with the CUDA log2 (<= 1 ulp)
    # the sum the H pass forms: relative error of sum x log2 y far below the 1e-10 step gate
    rng = np.random.default_rng(2)
    yy = y[400000:800000]
    x = rng.poisson(0.3, yy.size).astype(np.float64)
    s_dev = float(np.sum(np.maximum(x, 1e-14) * res[400000:800000]))
    s_ref = float(np.sum(np.maximum(x, 1e-14).astype(LD) * np.log2(yy.astype(LD))))
    assert abs(s_dev - s_ref) <= 1e-14 * abs(s_ref)
