"""Run one evidence call per filter configuration (for `ncu -k regex:knn_filter`).  Development harness.

    ncu --set full -k regex:knn_filter -o gpurun_out/variants python tools/ncu_variants.py --n 200000 --d 10 \
        --cfg 2,8,1 4,8,1 4,4,1 3,8,1
"""
import argparse
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcevidence_b200 import _lib                      # noqa: E402
from mcevidence_b200.synthetic import gaussian_chain  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=200000)
ap.add_argument('--d', type=int, default=10)
ap.add_argument('--kmax', type=int, default=4)
ap.add_argument('--chunks', type=int, default=-1)
ap.add_argument('--reps', type=int, default=1)
ap.add_argument('--qn', type=int, default=0, help='search only the first qn query rows (a rank shard)')
ap.add_argument('--tail', type=int, nargs='*', default=[-1])
ap.add_argument('--cfg', nargs='+', default=['2,8,1'], help="tq,tr,packed[,mode] | auto | tc (tcgen05 engine)")
a = ap.parse_args()
chain, _ = gaussian_chain(a.n, a.d, seed=0)
s, l, w = np.ascontiguousarray(chain[:, 2:]), -chain[:, 1], chain[:, 0]
ctx = _lib.Context(0)
ctx.set_option('ref_chunks', a.chunks)
for c, tail in [(c, t) for c in a.cfg for t in a.tail]:
    ctx.set_option('tail_chunks', tail)
    ctx.set_option('filter_engine', 1 if c == 'tc' else 0)
    parts = [-1, -1, -1, -1] if c in ('auto', 'tc') else [int(x) for x in c.split(',')]
    tq, tr, pk = parts[:3]
    sy = parts[3] if len(parts) > 3 else -1
    ctx.set_option('filter_tq', tq); ctx.set_option('filter_tr', tr); ctx.set_option('filter_packed', pk)
    ctx.set_option('filter_sync', sy)
    best = None
    for _ in range(a.reps):
        out = ctx.evidence(s, l, w, a.kmax, log_sum_w=math.log(w.sum()), q_range=(0, a.qn) if a.qn else None)
        st = out['stats']
        if best is None or st['ms_filter'] < best['ms_filter']:
            best = st
    st = best
    rate = st['pairs'] / (st['ms_filter'] * 1e-3)
    print('cfg=%s used=(%d,%d,%d,%d) chunks=%d grid=%d ctas/sm=%d filter=%.3f ms %.3e pairs/s fma=%.1f TF logE0=%.10f rescanned=%d' % (
        c, st['filter_tq'], st['filter_tr'], st['filter_packed'], st['filter_mode'], st['ref_chunks'], st['filter_grid'], st['filter_ctas_per_sm'],
        st['ms_filter'], rate, 2 * a.d * rate / 1e12, out['logE'][0], st['rows_rescanned']), flush=True)
