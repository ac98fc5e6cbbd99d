#!/bin/bash
# quick check of the tcgen05 filter: probe modes at one wave + full cfg2 / cfg5s-shaped calls
T=${1:-x}
mkdir -p gpurun_out
python tools/tc_probe.py --stages 8 --mma-warps 2 --out gpurun_out/r02_tc_probe_$T.json 2>&1 | tail -4
python - <<'P'
import numpy as np, sys, json
sys.path.insert(0,'.')
from mcevidence_b200 import _lib
from mcevidence_b200.synthetic import gaussian_chain
ctx=_lib.Context(0)
for n,d,k in ((1000000,10,4),(1000000,20,9),(500000,27,5),(1000000,5,4)):
    x=np.ascontiguousarray(gaussian_chain(n,d,seed=0)[0][:,2:])
    st=min((ctx.knn(x,k)[2] for _ in range(2)),key=lambda s:s['ms_filter'])
    print(n,d,k,'filter %.2f ms  rerank %.2f exact %.2f rescanned %d chunks %d grid %d mode %d'%(st['ms_filter'],st['ms_rerank'],st['ms_exact'],st['rows_rescanned'],st['ref_chunks'],st['filter_grid'],st['filter_mode']),flush=True)
P
