"""CPU check of the root-anchored classification used by h_finish (csrc/small.cuh, KlAnchorEval): for random
(num, den) columns, every midpoint of the lock-step bisection is classified by the anchor's bounds and compared with
the direct evaluation of f; reports how many midpoints the bounds decide and asserts that none is decided wrongly."""
import numpy as np

LS, TOL = 1e-14, 1e-5


def f_eval(num, den, x):
    return np.sum(np.maximum(num / (x + den), LS)) - 1.0


def anchor(num, den, a0, b0):
    if not (np.isfinite(a0) and np.isfinite(b0) and a0 < b0) or np.any(num < 0):
        return None
    x = a0
    conv = False
    for it in range(24):
        t = x + den
        r = 1.0 / t
        q = num * r
        un = q > LS
        S = np.sum(np.where(un, q, LS))
        D = np.sum(np.where(un, q * r, 0.0))
        F2 = np.sum(np.where(un, q * r * r, 0.0))
        tm = np.min(np.where(un, t, 1e300))
        if not (S > 0 and D > 0 and tm > 0):
            return None
        if abs(S - 1.0) <= 4e-15:
            conv = True
            break
        xn = x + S * (S - 1.0) / D
        conv = abs(xn - x) <= 4.5e-16 * abs(x)
        x = xn
        if conv:
            break
    if not conv or not (a0 < x < b0):
        return None
    span = max(abs(a0), abs(b0)) + np.max(np.abs(den))
    mx = 1e-12 / D + (8 + 2 * len(num)) * 2.3e-16 * span
    an = dict(nu=x, mx=mx, its=it, eLlo=0.0, eRlo=0.0, eLhi=np.inf, eRhi=np.inf)
    F2 = 2 * F2
    if 1e-15 * span <= 1e-10 * tm:
        hi, lo = TOL * (1 + 1e-9), TOL * (1 - 1e-9)
        an["eLhi"] = max(hi / D, mx)
        an["eLlo"] = min(0.01 * tm, lo / (D + 0.52 * F2 * lo / D))
        an["eRlo"] = lo / D
        q = F2 * hi / D / D
        chord_e = hi * (b0 - x) / 0.4999
        an["eRhi"] = max(min(chord_e, hi / D / (1 - q)) if q < 0.25 else chord_e, mx)
    return an


def classify(an, x):
    """(le0, bad) or None when the bounds do not decide."""
    d = x - an["nu"]
    e = abs(d)
    if e <= an["mx"]:
        return None
    right = d > 0
    if e > (an["eRhi"] if right else an["eLhi"]):
        return (right, True)
    if e < (an["eRlo"] if right else an["eLlo"]):
        return (right, False)
    return (right, None)


def run(rng, k, kind):
    num = rng.uniform(0, 1, k) ** rng.integers(1, 6)
    den = rng.uniform(0, 1, k) ** rng.integers(1, 4) * 10 ** rng.uniform(-3, 2)
    if kind == 1:
        num[rng.integers(k)] = 0.0
    if kind == 2:
        num *= 10 ** rng.uniform(-8, 4)
    if kind == 3:
        num[:] = 1e-14 * rng.uniform(0.5, 2, k)          # an "all-zero pixel": root a few ulps from -den
        den[:] = rng.uniform(0.5, 2.0) + rng.uniform(0, 1e-3, k)
    if kind == 4:
        den[rng.integers(k)] = 0.0
    a = np.max(np.where(num > 0, num / 2 - den, -np.inf))
    b = k * np.max(num) / 0.5 - np.min(den)
    an = anchor(num, den, a, b)
    stats = dict(mid=0, decided=0, signonly=0, newton=0 if an is None else an["its"], anchored=an is not None)
    if an is None:
        return stats
    for j in range(100):
        new = (a + b) / 2
        fe = f_eval(num, den, new)
        stats["mid"] += 1
        c = classify(an, new)
        if c is not None:
            assert c[0] == (fe <= 0), ("sign", kind, num, den, new, fe, an)
            if c[1] is None:
                stats["signonly"] += 1
            else:
                assert c[1] == (abs(fe) > TOL), ("size", kind, num, den, new, fe, an)
                stats["decided"] += 1
        if new == a or new == b:
            break
        if fe <= 0:
            b = new
        else:
            a = new
    return stats


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    tot = dict(mid=0, decided=0, signonly=0, newton=0, anchored=0, n=0)
    with np.errstate(all="ignore"):
        for i in range(40000):
            s = run(rng, int(rng.integers(2, 9)), i % 5)
            for key in ("mid", "decided", "signonly", "newton", "anchored"):
                tot[key] += s[key]
            tot["n"] += 1
    print("columns %d  anchored %.4f  midpoints %d  decided by bounds %.4f  sign only %.5f  mean Newton its %.2f" % (
        tot["n"], tot["anchored"] / tot["n"], tot["mid"], tot["decided"] / tot["mid"], tot["signonly"] / tot["mid"],
        tot["newton"] / max(tot["anchored"], 1)))
