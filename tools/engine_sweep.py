"""FFMA filter vs tcgen05 filter over (ndim, N): ms_filter of both engines for the engine-selection rule
(development harness; the chosen rule lives in csrc/api.cu).

    python tools/engine_sweep.py --out gpurun_out/engine_sweep.json
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
ap.add_argument('--dims', type=int, nargs='*', default=[2, 3, 4, 5, 6, 8, 10, 14, 20, 27, 40, 64])
ap.add_argument('--ns', type=int, nargs='*', default=[6000, 15000, 40000, 100000, 300000, 1000000])
ap.add_argument('--k', type=int, default=3)
ap.add_argument('--out', default=None)
a = ap.parse_args()
ctx = _lib.Context(0)
rows = []
for d in a.dims:
    for n in a.ns:
        if n * d > 3e7:
            continue
        chain, _ = gaussian_chain(n, d, seed=1)
        x = np.ascontiguousarray(chain[:, 2:])
        rec = {'d': d, 'n': n}
        for eng, name in ((0, 'ffma'), (1, 'tc')):
            ctx.set_option('filter_engine', eng)
            best = None
            for _ in range(3):
                st = ctx.knn(x, a.k)[2]
                if best is None or st['ms_filter'] < best['ms_filter']:
                    best = st
            rec[name] = best['ms_filter']
            rec[name + '_total'] = best['ms_total'] - best['ms_h2d'] - best['ms_d2h']
            rec[name + '_resc'] = best['rows_rescanned']
        rec['speedup'] = rec['ffma'] / rec['tc']
        rows.append(rec)
        print(rec, flush=True)
if a.out:
    json.dump(rows, open(a.out, 'w'), indent=1)
