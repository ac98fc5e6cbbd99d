"""tcgen05 filter over problem sizes, forced on (development harness): short scans stay in the start-up regime where
nearly every tile carries hits and the selection warp sets the pace."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcevidence_b200 import _lib
from mcevidence_b200.synthetic import gaussian_chain
if os.environ.get('MCE_LIB'):          # a variant built by tools/build_variant.sh
    _lib.LIB_PATH = os.path.abspath(os.environ['MCE_LIB'])
ctx = _lib.Context(0)
for n, d, k in ((25000, 7, 2), (46000, 8, 2), (99000, 8, 2), (200000, 8, 2), (400000, 10, 4), (1000000, 10, 4), (500000, 27, 5), (1000000, 20, 9)):
    x = np.ascontiguousarray(gaussian_chain(n, d, seed=1)[0][:, 2:])
    out = []
    for eng in ((1,) if os.environ.get('MCE_LIB') else (1, 0)):
        ctx.set_option('filter_engine', eng)
        st = min((ctx.knn(x, k)[2] for _ in range(3)), key=lambda s: s['ms_filter'])
        out.append(st['ms_filter'])
    print('n %8d d %2d k %d  tcgen05 %.3f ms  ffma %s ms' % (n, d, k, out[0], ('%.3f' % out[1]) if len(out) > 1 else '-'), flush=True)
