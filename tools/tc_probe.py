"""Timing probes of the tcgen05 filter engine: ring depth x perf mode at one shape (development harness).

    python tools/tc_probe.py --n 1000000 --d 10 --qn 37888
perf mode 0 = full kernel, 1 = MMA + TMA pipeline alone (epilogue skips the TMEM loads), 2 = plus the loads.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcevidence_b200 import _lib                      # noqa: E402
from mcevidence_b200.synthetic import gaussian_chain  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=1000000)
ap.add_argument('--d', type=int, default=10)
ap.add_argument('--qn', type=int, default=37888)
ap.add_argument('--k', type=int, default=5)
ap.add_argument('--stages', type=int, nargs='*', default=[2, 3, 4, 6, 8])
ap.add_argument('--modes', type=int, nargs='*', default=[1, 2, 0])
ap.add_argument('--mma-warps', type=int, nargs='*', default=[1, 2])
ap.add_argument('--out', default=None)
a = ap.parse_args()
chain, _ = gaussian_chain(a.n, a.d, seed=0)
x = np.ascontiguousarray(chain[:, 2:])
ctx = _lib.Context(0)
ctx.set_option('filter_engine', 1)
rows = []
for mw in a.mma_warps:
    ctx.set_option('tc_mma_warps', mw)
    for stages in a.stages:
        ctx.set_option('tc_stages', stages)
        for mode in a.modes:
            ctx.set_option('tc_perf_mode', mode)
            ms = min(ctx.knn(x, a.k, q_range=(0, a.qn) if a.qn else None)[2]['ms_filter'] for _ in range(3))
            rows.append({'n': a.n, 'd': a.d, 'qn': a.qn, 'mma_warps': mw, 'stages': stages, 'perf_mode': mode, 'ms_filter': ms})
            print(rows[-1], flush=True)
if a.out:
    json.dump(rows, open(a.out, 'w'), indent=1)
