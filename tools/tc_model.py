"""CPU model of the coarse stage of the tcgen05 filter engine (csrc/knn_filter_tc.cu): TF32-truncated operands
(products exact, h split into a TF32-exact head and a remainder), the allowance e used by the kernel, and a
sequential replay of the threshold test per query.  Reports the worst |coarse - exact| / e and how many columns
reach the exact slow path per query ("visits") next to the insertions an exact-threshold filter makes.

    python tools/tc_model.py

Used by tests/test_tc_model.py at a small size; the numbers quoted in DESIGN.md come from this script.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcevidence_b200.synthetic import gaussian_chain  # noqa: E402

def trunc_tf32(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)

def model(n, d, nq, L, seed=0, slack_scale=1.25):
    chain, _ = gaussian_chain(n, d, seed=seed)
    x = chain[:, 2:]
    mu = x.mean(0); C = np.cov(x.T); Lc = np.linalg.cholesky(C); W = np.linalg.inv(Lc).T
    y = ((x - mu) @ W)
    yf = y.astype(np.float32)
    h = (-0.5 * (y * y).sum(1)).astype(np.float32)
    h_hi = trunc_tf32(h.copy()); h_lo = h - h_hi
    rt = trunc_tf32(yf.copy()); hl_t = trunc_tf32(h_lo.copy())
    rmax = float(np.sqrt((y*y).sum(1).max()))
    rng = np.random.default_rng(1)
    qi = rng.choice(n, nq, replace=False)
    q = yf[qi]; qt = trunc_tf32(q.copy())
    # coarse in float64 of truncated operands (the tensor core's products are exact; accumulation error is allowed for separately)
    coarse = (qt.astype(np.float64) @ rt.astype(np.float64).T) + h_hi.astype(np.float64)[None, :] + hl_t.astype(np.float64)[None, :]
    # exact fp32 chain
    acc = q[:, 0:1] * yf[None, :, 0] ; acc = (acc.astype(np.float64) + h[None, :].astype(np.float64)).astype(np.float32)  # fma(q0,r0,h) rounded once
    for c in range(1, d):
        acc = (q[:, c:c+1].astype(np.float64) * yf[None, :, c].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    qn = np.sqrt((q.astype(np.float64)**2).sum(1))
    s_qr = slack_scale * (2.0**-9 * (1 + 2.0**-10) + 2.0**-14)
    s_rr = slack_scale * 0.5 * (2.0**-20 + 2.0**-14)
    e = s_qr * qn * rmax + s_rr * rmax * rmax
    ratio = np.abs(coarse - acc.astype(np.float64)) / e[:, None]
    # streaming filter simulation per query
    visits = np.zeros(nq, int); inserts = np.zeros(nq, int); exact_inserts = np.zeros(nq, int)
    for a in range(nq):
        s = acc[a].copy(); cs = coarse[a]
        s[qi[a]] = -np.inf            # self excluded (its visit is counted separately: +1)
        lst = np.full(L, -np.inf, np.float32); th = -np.inf
        # sequential replay; only references that can pass at some point are walked: the first 4096 (lists
        # still filling) and those within the allowance of the final threshold
        kth = np.partition(s, -L)[-L]
        cand = np.nonzero((cs > kth - e[a] - 1e-3) | (np.arange(n) < 4096))[0]   # early refs always examined
        v = 0; ins = 0
        for j in cand:
            if cs[j] > th - e[a]:
                v += 1
                if s[j] > th:
                    m = lst.argmin(); lst[m] = s[j]; th = lst.min(); ins += 1
        visits[a] = v + 1; inserts[a] = ins
        # exact-threshold filter (FFMA engine): visits = inserts
        lst = np.full(L, -np.inf, np.float32); th = -np.inf; ins2 = 0
        for j in cand:
            if s[j] > th:
                m = lst.argmin(); lst[m] = s[j]; th = lst.min(); ins2 += 1
        exact_inserts[a] = ins2
    return dict(n=n, d=d, L=L, max_ratio=float(ratio.max()), mean_e=float(e.mean()), visits=float(visits.mean()),
                inserts=float(inserts.mean()), ffma_inserts=float(exact_inserts.mean()), rmax=rmax)

def bootstrap_model(n, d, nq, L, boot, seed=0, slack_scale=1.25):
    """The multi-tile threshold bootstrap of the kernel: per query and column half, the maxima of the 32-column groups
    of the first `boot` tiles (coarse scores), the (L+1)-th largest of them minus the allowance, the larger of the two
    halves.  Returns whether that start threshold is valid for every query (at least L references other than the query
    itself have an exact score above it) and the slow-path visits of a scan that starts from it, next to those of a
    scan that starts from the first-tile bootstrap."""
    chain, _ = gaussian_chain(n, d, seed=seed)
    x = chain[:, 2:]
    mu = x.mean(0); C = np.cov(x.T); Lc = np.linalg.cholesky(C); W = np.linalg.inv(Lc).T
    y = ((x - mu) @ W)
    yf = y.astype(np.float32)
    h = (-0.5 * (y * y).sum(1)).astype(np.float32)
    h_hi = trunc_tf32(h.copy()); h_lo = h - h_hi
    rt = trunc_tf32(yf.copy()); hl_t = trunc_tf32(h_lo.copy())
    rmax = float(np.sqrt((y * y).sum(1).max()))
    rng = np.random.default_rng(1)
    qi = rng.choice(n, nq, replace=False)
    q = yf[qi]; qt = trunc_tf32(q.copy())
    coarse = (qt.astype(np.float64) @ rt.astype(np.float64).T) + h_hi.astype(np.float64)[None, :] + hl_t.astype(np.float64)[None, :]
    acc = q[:, 0:1] * yf[None, :, 0]; acc = (acc.astype(np.float64) + h[None, :].astype(np.float64)).astype(np.float32)
    for c in range(1, d):
        acc = (q[:, c:c+1].astype(np.float64) * yf[None, :, c].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    qn = np.sqrt((q.astype(np.float64) ** 2).sum(1)) * 1.000001
    e = slack_scale * ((2.0**-9 * (1 + 2.0**-10) + 2.0**-14) * qn * rmax + 0.5 * (2.0**-20 + 2.0**-14) * rmax * rmax)

    def start_threshold(a, tiles):
        cs = coarse[a, :tiles * 128].reshape(tiles, 2, 2, 32)      # [tile][column half][group][column]
        best = -np.inf
        for half in range(2):
            gm = np.sort(cs[:, half].max(axis=2).reshape(-1))[::-1]                 # group maxima of this half
            cols = np.sort(cs[0, half].reshape(-1))[::-1] if tiles == 1 else gm     # one tile: all 64 columns
            if len(cols) > L and np.isfinite(cols[L]):
                best = max(best, cols[L] - e[a] * 1.001 - 1e-30)
        return best

    def visits_from(a, th0):
        s = acc[a].copy(); cs = coarse[a]
        s[qi[a]] = -np.inf
        lst = np.full(L, -np.inf, np.float32); th = np.float32(th0); v = 0
        kth = np.partition(s, -L)[-L]
        cand = np.nonzero((cs > min(kth, th0) - e[a] - 1e-3) | (np.arange(n) < 4096))[0]
        for j in cand:
            if cs[j] > th - e[a]:
                v += 1
                if s[j] > th:
                    m = lst.argmin(); lst[m] = s[j]; th = max(th, lst.min()) if np.isfinite(lst.min()) else th
        return v, lst

    valid = True
    v_boot = np.zeros(nq); v_first = np.zeros(nq); same = True
    for a in range(nq):
        s = acc[a].copy(); s[qi[a]] = -np.inf
        th_b, th_1 = start_threshold(a, boot), start_threshold(a, 1)
        valid &= bool((s > th_b).sum() >= L) and bool((s > th_1).sum() >= L)
        vb, lb = visits_from(a, th_b)
        v1, l1 = visits_from(a, th_1)
        v_boot[a], v_first[a] = vb, v1
        top = np.sort(s)[::-1][:L]
        same &= bool(np.array_equal(np.sort(lb)[::-1], top)) and bool(np.array_equal(np.sort(l1)[::-1], top))
    return dict(n=n, d=d, L=L, boot=boot, valid=valid, lists_exact=same, visits_boot=float(v_boot.mean()),
                visits_first_tile=float(v_first.mean()))


if __name__ == '__main__':
    for (n, d, nq, L) in [(100000, 10, 200, 5), (100000, 5, 200, 5), (100000, 20, 200, 9), (100000, 3, 200, 5), (400000, 10, 100, 5)]:
        t = time.time(); r = model(n, d, nq, L); r['sec'] = round(time.time() - t, 1); print(r, flush=True)
    for (n, d, nq, L, boot) in [(100000, 10, 100, 5, 64), (100000, 20, 100, 11, 64), (50000, 6, 100, 4, 32)]:
        print(bootstrap_model(n, d, nq, L, boot), flush=True)
