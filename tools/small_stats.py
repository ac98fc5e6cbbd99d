"""Stage timings of small (Planck-grid sized) problems: strided vs contiguous host input.  Development harness."""
import math, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcevidence_b200 import _lib
from mcevidence_b200.synthetic import gaussian_chain
ctx = _lib.Context(0)
for n, d in ((15000, 7), (45979, 6), (98886, 8), (5917, 7)):
    c, _ = gaussian_chain(n, d, seed=1, inflate=1.0, integer_weights=True, nextra=3)
    for name, s in (('strided', c[:, 2:]), ('contig ', np.ascontiguousarray(c[:, 2:2 + d]))):
        best = None
        for _ in range(5):
            t0 = time.perf_counter()
            out = ctx.evidence(s, -c[:, 1], c[:, 0], 2, ndim=d, log_sum_w=math.log(c[:, 0].sum()))
            wall = time.perf_counter() - t0
            st = out['stats']
            if best is None or wall < best[0]:
                best = (wall, st)
        wall, st = best
        print('n=%6d d=%d %s wall=%.3f ms total=%.3f h2d=%.3f cov=%.3f whiten=%.3f filter=%.3f rerank=%.3f exact=%.3f reduce=%.3f d2h=%.3f grid=%d chunks=%d launches=%d' % (
            n, d, name, wall * 1e3, st['ms_total'], st['ms_h2d'], st['ms_cov'], st['ms_whiten'], st['ms_filter'], st['ms_rerank'],
            st['ms_exact'], st['ms_reduce'], st['ms_d2h'], st['filter_grid'], st['ref_chunks'], st['launches']), flush=True)
