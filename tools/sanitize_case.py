"""Small auto / cross / exact-only calls for compute-sanitizer (memcheck, racecheck, synccheck).

    compute-sanitizer --tool memcheck python tools/sanitize_case.py
"""
import math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcevidence_b200 import _lib
from mcevidence_b200.synthetic import gaussian_chain
import argparse
ap = argparse.ArgumentParser()
ap.add_argument('--sync', type=int, default=-1, help='filter_sync of the FFMA filter (1, 2, 3; -1 = automatic)')
ap.add_argument('--engine', type=int, default=0, help='0 FFMA filter, 1 tcgen05 filter')
ap.add_argument('--long', action='store_true', help='tcgen05 engine: add a unit long enough for the multi-tile bootstrap and the slab hit path')
a = ap.parse_args()
ctx = _lib.Context(0)
ctx.set_option('filter_engine', a.engine)
if a.sync >= 0:
    ctx.set_option('filter_sync', a.sync)
print('sanitize_case: filter_engine=%d filter_sync=%d' % (a.engine, a.sync), flush=True)
for (n, d, kmax) in ((3000, 6, 4), (1111, 17, 3), (700, 40, 2)):
    c, _ = gaussian_chain(n, d, seed=n, nextra=2)
    s, l, w = c[:, 2:], -c[:, 1], c[:, 0]
    a1 = ctx.evidence(s, l, w, kmax, ndim=d, log_sum_w=math.log(w.sum()), want_knn=True)
    h = n // 2
    x = ctx.evidence(s[:h], l[:h], w[:h], kmax, ref=s[h:], ndim=d, log_sum_w=0.0)
    e = ctx.evidence(s, l, w, kmax, ndim=d, log_sum_w=math.log(w.sum()), exact_only=True)
    sw = ctx.evidence_sweep(s, l, w, kmax, [n // 3, n], ndim=d, log_sum_w=math.log(w.sum()))
    keep = np.arange(0, n, 2)
    rm = ctx.evidence(s, l[keep], w[keep], kmax, ndim=d, query_rows=keep, log_sum_w=0.0)
    print(n, d, a1['stats']['filter_mode'], a1['logE'], x['logE'], float(np.abs(a1['logE'] - e['logE']).max()),
          float(np.abs(sw[1]['logE'] - a1['logE']).max()), rm['logE'], flush=True)
if a.long:
    c, _ = gaussian_chain(170000, 6, seed=5)
    x = np.ascontiguousarray(c[:, 2:])
    ctx.set_option('ref_chunks', 1)
    dk, ix, st = ctx.knn(x, 2, q_range=(0, 148 * 256))     # one wave of whole-range units: 1329 tiles, boot 110, slabs from tile 1200
    ctx.set_option('ref_chunks', -1)
    print('long unit', st['filter_mode'], float(dk[:, 1].mean()), flush=True)
bt = ctx.evidence_batch([{'query': c[i::3, 2:], 'logl': -c[i::3, 1], 'weights': c[i::3, 0], 'kmax': 2, 'ndim': 5} for i in range(3)])
print('batch', [o['logE'] for o in bt])
print('SANITIZE_CASE_DONE')
