"""Filter time of the tcgen05 engine for combinations of tc_variant / tc_mma_warps / tc_perf_mode (development harness).

    python tools/tc_variants.py --out gpurun_out/tc_variants.json
One wave (148 CTAs x 256 queries) of the cfg2 shape by default, plus the full problem for the listed variants.
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
ap.add_argument('--variants', type=int, nargs='*', default=[0, 1])
ap.add_argument('--mma-warps', type=int, nargs='*', default=[2, 1])
ap.add_argument('--modes', type=int, nargs='*', default=[0, 3, 1])
ap.add_argument('--full', type=int, nargs='*', default=[0, 1], help='variants also timed on the whole problem')
ap.add_argument('--out', default=None)
a = ap.parse_args()
chain, _ = gaussian_chain(a.n, a.d, seed=0)
x = np.ascontiguousarray(chain[:, 2:])
ctx = _lib.Context(0)
ctx.set_option('filter_engine', 1)
rows = []
ref = None
for var in a.variants:
    ctx.set_option('tc_variant', var)
    for mw in a.mma_warps:
        ctx.set_option('tc_mma_warps', mw)
        for mode in a.modes:
            ctx.set_option('tc_perf_mode', mode)
            ms = min(ctx.knn(x, a.k, q_range=(0, a.qn))[2]['ms_filter'] for _ in range(3))
            tiles = (a.n + 127) // 128
            rows.append({'n': a.n, 'd': a.d, 'qn': a.qn, 'variant': var, 'mma_warps': mw, 'perf_mode': mode, 'ms_filter': ms,
                         'clk_per_tile_at_1965MHz': ms * 1e-3 * 1.965e9 / tiles})
            print(rows[-1], flush=True)
ctx.set_option('tc_perf_mode', 0)
ctx.set_option('tc_mma_warps', 2)
for var in a.full:
    ctx.set_option('tc_variant', var)
    best = None
    for _ in range(3):
        dk, idx, st = ctx.knn(x, a.k)
        best = st['ms_filter'] if best is None else min(best, st['ms_filter'])
    if ref is None:
        ref = (dk.copy(), idx.copy())
    same = bool(np.array_equal(ref[0], dk) and np.array_equal(ref[1], idx))
    rows.append({'n': a.n, 'd': a.d, 'full': True, 'variant': var, 'ms_filter': best, 'identical_to_first': same, 'n_exact': st.get('n_exact')})
    print(rows[-1], flush=True)
if a.out:
    json.dump(rows, open(a.out, 'w'), indent=1)
