"""Hand-over timeline of the tcgen05 filter (CTA 0, first 1024 tiles): clock64() of the events listed at TC_MARK in
csrc/knn_filter_tc.cu, for the full kernel (perf mode 0) and the MMA + TMA pipeline alone (perf mode 1).

    python tools/tc_timeline.py --out gpurun_out/tc_timeline.npz
"""
import argparse
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
ap.add_argument('--out', default='gpurun_out/tc_timeline.npz')
ap.add_argument('--mma-warps', type=int, default=2)
ap.add_argument('--skip', type=int, default=0, help='leading tiles of CTA 0 left out of the statistics')
a = ap.parse_args()
chain, _ = gaussian_chain(a.n, a.d, seed=0)
x = np.ascontiguousarray(chain[:, 2:])
ctx = _lib.Context(0)
ctx.set_option('filter_engine', 1)
ctx.set_option('tc_timeline', 1)
ctx.set_option('tc_mma_warps', a.mma_warps)
res = {}
for mode in (0, 1):
    ctx.set_option('tc_perf_mode', mode)
    for _ in range(2):
        st = ctx.knn(x, 5, q_range=(0, a.qn))[2]
    t = ctx.get_debug_array('tc_dump').reshape(-1).view(np.int64).reshape(1024, 16)
    res['mode%d' % mode] = t
    res['ms%d' % mode] = st['ms_filter']
    tt = t[100:900].astype(np.float64)
    per_tile = (tt[-1, 0] - tt[0, 0]) / (len(tt) - 1)
    print('mode %d: filter %.3f ms, %.0f clk per tile' % (mode, st['ms_filter'], per_tile))
    names = ['mma:landed', 'mma:free', 'mma:issued', 'unused', 'unused', 'epi0:full', 'epi0:released', 'epi0:done',
             'epi1:full', 'epi1:released', 'epi1:done', 'prod:slotfree']
    base = tt[:, 0:1]
    rel = tt[:, :12] - base
    for i, nm in enumerate(names):
        print('  %-14s mean %+8.0f  p10 %+8.0f  p90 %+8.0f   (relative to mma:landed of the same tile)' % (
            nm, rel[:, i].mean(), np.percentile(rel[:, i], 10), np.percentile(rel[:, i], 90)))
    # latencies of the chain
    print('  issued -> epi0:full (MMA + commit + wake)      mean %.0f' % (tt[:, 5] - tt[:, 2]).mean())
    print('  issued -> epi1:full                            mean %.0f' % (tt[:, 8] - tt[:, 2]).mean())
    print('  epi0:released(it) -> mma:free(it+2) observed   mean %.0f' % (tt[2:, 1] - tt[:-2, 6]).mean())
    print('  epi0:full -> released (hold)                   mean %.0f' % (tt[:, 6] - tt[:, 5]).mean())
    print('  epi0:full -> done                              mean %.0f' % (tt[:, 7] - tt[:, 5]).mean())
np.savez(a.out, **res)
