"""Small auto / cross / exact-only calls for compute-sanitizer (memcheck, racecheck, synccheck).

    compute-sanitizer --tool memcheck python tools/sanitize_case.py
"""
import math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcevidence_b200 import _lib
from mcevidence_b200.synthetic import gaussian_chain
ctx = _lib.Context(0)
for (n, d, kmax) in ((3000, 6, 4), (1111, 17, 3), (700, 40, 2)):
    c, _ = gaussian_chain(n, d, seed=n, nextra=2)
    s, l, w = c[:, 2:], -c[:, 1], c[:, 0]
    a = ctx.evidence(s, l, w, kmax, ndim=d, log_sum_w=math.log(w.sum()), want_knn=True)
    h = n // 2
    x = ctx.evidence(s[:h], l[:h], w[:h], kmax, ref=s[h:], ndim=d, log_sum_w=0.0)
    e = ctx.evidence(s, l, w, kmax, ndim=d, log_sum_w=math.log(w.sum()), exact_only=True)
    for mode in (0, 1):
        ctx.set_option('filter_sync', mode)
        m = ctx.evidence(s, l, w, kmax, ndim=d, log_sum_w=math.log(w.sum()))
        assert np.abs(m['logE'] - a['logE']).max() < 1e-12
    ctx.set_option('filter_sync', -1)
    print(n, d, a['logE'], x['logE'], float(np.abs(a['logE'] - e['logE']).max()), flush=True)
print('SANITIZE_CASE_DONE')
