"""A/B timing of builds of the tcgen05 filter (tools/build_variant.sh) on one GPU box, one process per library:

    for v in old coop; do python tools/tc_ab.py --lib mcevidence_b200/variants/lib_$v.so --tag $v; done
Prints filter times of one wave (148 CTAs) in perf modes 0 / 3 / 1 and of the whole problem; results of the whole problem are
compared bit for bit with those of the first library run (kept in /tmp).
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
ap.add_argument('--lib', default=None)
ap.add_argument('--tag', default='default')
ap.add_argument('--n', type=int, default=1000000)
ap.add_argument('--d', type=int, default=10)
ap.add_argument('--k', type=int, default=5)
ap.add_argument('--modes', type=int, nargs='*', default=[0, 3, 1])
ap.add_argument('--mma-warps', type=int, default=2)
ap.add_argument('--out', default=None)
ap.add_argument('--probe-mode', type=int, default=0)
a = ap.parse_args()
if a.lib:
    _lib.LIB_PATH = os.path.abspath(a.lib)
chain, _ = gaussian_chain(a.n, a.d, seed=0)
x = np.ascontiguousarray(chain[:, 2:])
ctx = _lib.Context(0)
ctx.set_option('filter_engine', 1)
ctx.set_option('tc_mma_warps', a.mma_warps)
tiles = (a.n + 127) // 128
row = {'tag': a.tag, 'n': a.n, 'd': a.d, 'k': a.k, 'mma_warps': a.mma_warps}
for mode in a.modes:
    ctx.set_option('tc_perf_mode', mode)
    ms = min(ctx.knn(x, a.k, q_range=(0, 148 * 256))[2]['ms_filter'] for _ in range(3))
    row['wave_mode%d_clk_per_tile' % mode] = round(ms * 1e-3 * 1.965e9 / tiles, 1)
ctx.set_option('tc_perf_mode', 0)
best = None
for _ in range(3):
    dk, idx, st = ctx.knn(x, a.k)
    best = st['ms_filter'] if best is None else min(best, st['ms_filter'])
row['full_ms_filter'] = round(best, 3)
try:
    import ctypes
    pr = (ctypes.c_ulonglong * 64)()
    ctx._lib.mce_tc_probe(pr)          # clear
    ctx.set_option('tc_perf_mode', a.probe_mode)
    ctx.knn(x, a.k)
    ctx._lib.mce_tc_probe(pr)
    v = list(pr)
    if any(v):
        row['probe'] = {'slab_entries': v[0], 'slab_clk_per_entry': v[1] / max(v[0], 1), 'slab_taken': v[2],
                        'scan_entries': v[4], 'scan_clk_per_entry': v[3] / max(v[4], 1),
                        'batches': v[5], 'items_per_batch': v[6] / max(v[5], 1), 'clk_per_batch': v[7] / max(v[5], 1),
                        'slab_events': v[8], 'clk_per_slab_event': v[9] / max(v[8], 1),
                        'epi_wait_full_clk_per_tile': [round(2 * x / max(v[34], 1), 1) for x in v[16:32]],
                        'mma_wait_free_clk_per_tile': [round(4 * x / max(v[34], 1), 1) for x in v[32:34]]}
except AttributeError:
    pass
ref = '/tmp/tc_ab_ref_%d_%d_%d.npz' % (a.n, a.d, a.k)
if os.path.exists(ref):
    z = np.load(ref)
    row['identical_to_first'] = bool(np.array_equal(z['dk'], dk) and np.array_equal(z['idx'], idx))
else:
    np.savez(ref, dk=dk, idx=idx)
    row['identical_to_first'] = None
print(json.dumps(row), flush=True)
if a.out:
    with open(a.out, 'a') as f:
        f.write(json.dumps(row) + '\n')
