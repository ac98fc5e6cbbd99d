"""Batch-size sweep (N4): mce_evidence_sweep against one mce_evidence call per batch size and against one full call.

    python tools/sweep_bench.py [--n 1000000] [--d 10] [--nbatch 10] [--out gpurun_out/sweep.json]
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcevidence_b200 import _lib                      # noqa: E402
from mcevidence_b200.synthetic import gaussian_chain  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=1000000)
ap.add_argument('--d', type=int, default=10)
ap.add_argument('--kmax', type=int, default=4)
ap.add_argument('--nbatch', type=int, default=10)
ap.add_argument('--scale', default='linear', choices=['linear', 'logpower'])
ap.add_argument('--out', default=None)
a = ap.parse_args()
chain, _ = gaussian_chain(a.n, a.d, seed=0)
s, logl, w = np.ascontiguousarray(chain[:, 2:]), -chain[:, 1], np.ascontiguousarray(chain[:, 0])
if a.scale == 'linear':
    rows = [int(x) for x in np.linspace(a.n / a.nbatch, a.n, a.nbatch)]
else:
    rows = [int(10.0 ** x) for x in np.linspace(math.log10(a.n) - 2.0, math.log10(a.n), a.nbatch)]
rows[-1] = a.n
cov = np.cov(s.T)
kw = dict(ndim=a.d, W=np.linalg.inv(np.linalg.cholesky(cov)).T.copy(), mu=s.mean(axis=0),
          log_jacobian=0.5 * np.linalg.slogdet(cov)[1], log_sum_w=math.log(w.sum()))
ctx = _lib.Context(0)


def wall(f, reps=2):
    best, out = 1e30, None
    for _ in range(reps):
        t = time.perf_counter()
        out = f()
        best = min(best, time.perf_counter() - t)
    return best, out


t_full, full = wall(lambda: ctx.evidence(s, logl, w, a.kmax, **kw))
t_sweep, sw = wall(lambda: ctx.evidence_sweep(s, logl, w, a.kmax, rows, **kw))
t_loop, loop = wall(lambda: [ctx.evidence(s[:r], logl[:r], w[:r], a.kmax, **kw) for r in rows], reps=1)
same = all(np.array_equal(x['logE'], y['logE']) for x, y in zip(sw, loop))
rec = {'n': a.n, 'd': a.d, 'kmax': a.kmax, 'rows': rows, 'wall_s_one_full_call': t_full, 'wall_s_sweep': t_sweep,
       'wall_s_one_call_per_batch': t_loop, 'sweep_over_full': t_sweep / t_full, 'loop_over_full': t_loop / t_full,
       'bit_identical_to_per_batch_calls': bool(same), 'sweep_stats': sw[-1]['stats'], 'full_stats': full['stats'],
       'pairs_sum_over_batches': float(sum(float(r) ** 2 for r in rows)), 'pairs_largest': float(a.n) ** 2}
print(json.dumps(rec), flush=True)
if a.out:
    json.dump(rec, open(a.out, 'w'), indent=1)
