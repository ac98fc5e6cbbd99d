import sys, os
sys.path.insert(0, '.')
import numpy as np
from mcevidence_b200 import _lib
from mcevidence_b200.synthetic import gaussian_chain
ctx = _lib.Context(0)
for n, d, k in ((8000, 7, 2), (12000, 7, 2), (15000, 7, 2), (18760, 6, 2), (20000, 8, 2), (15000, 4, 2), (20000, 5, 3), (12000, 10, 4), (9000, 14, 4)):
    x = np.ascontiguousarray(gaussian_chain(n, d, seed=1)[0][:, 2:])
    out = []
    for eng in (1, 0):
        ctx.set_option('filter_engine', eng)
        sts = [ctx.knn(x, k)[2] for _ in range(4)]
        st = min(sts, key=lambda s: s['ms_filter'])
        out.append((st['ms_filter'], min(s['ms_total'] - s['ms_h2d'] - s['ms_d2h'] for s in sts)))
    print('n %6d d %2d k %d  tcgen05 filter %.3f total %.3f ms | ffma filter %.3f total %.3f ms' % (n, d, k, out[0][0], out[0][1], out[1][0], out[1][1]), flush=True)
