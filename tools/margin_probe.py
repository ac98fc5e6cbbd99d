import sys; sys.path.insert(0,'.')
import numpy as np
from mcevidence_b200 import _lib
from mcevidence_b200.synthetic import gaussian_chain
ctx=_lib.Context(0)
x=np.ascontiguousarray(gaussian_chain(1000000,10,seed=0)[0][:,2:])
for m in (2,1,0):
    ctx.set_option('list_margin',m)
    st=min((ctx.knn(x,4)[2] for _ in range(2)),key=lambda s:s['ms_filter'])
    print('margin',m,'list_len',st['list_len'],'filter %.2f ms rerank %.2f exact %.2f rescanned %d'%(st['ms_filter'],st['ms_rerank'],st['ms_exact'],st['rows_rescanned']),flush=True)
