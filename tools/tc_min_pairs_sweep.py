import sys, json, os, time, math
sys.path.insert(0, '.')
import numpy as np
from mcevidence_b200 import _lib
from mcevidence_b200.grid import evidence_many
from mcevidence_b200.synthetic import gaussian_chain
shapes = json.load(open('mcevidence_b200/planck_grid_shapes.json'))['AllChains']
cells = [(nd, n) for (_, _, nd, ns) in shapes for n in ns]
chains = [gaussian_chain(n, nd, seed=i, inflate=1.0, integer_weights=True, nextra=3)[0] for i, (nd, n) in enumerate(cells)]
ndims = [nd for nd, _ in cells]
ctx = _lib.default_context(0)
for v in (0, 5000, 1200, 600, 300, 150, 80, 40):
    ctx.set_option('tc_min_mpairs', v)
    for lanes in (8, 4):
        evidence_many(chains, kmax=2, ndim=ndims, nstreams=lanes)
        ts = []
        for _ in range(3):
            t = time.perf_counter(); evidence_many(chains, kmax=2, ndim=ndims, nstreams=lanes); ts.append(time.perf_counter() - t)
        print('tc_min_mpairs', v, 'lanes', lanes, 'wall %.1f ms' % (1e3 * min(ts)), flush=True)
