"""Run under torchrun on >= 2 GPUs: row-sharded evidence (NCCL broadcast + all-reduce) must equal the
single-GPU result.  Prints 'DIST_CHECK OK' on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py
"""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcevidence_b200 import _lib                      # noqa: E402
from mcevidence_b200.dist import evidence_sharded     # noqa: E402
from mcevidence_b200.grid import evidence_many        # noqa: E402
from mcevidence_b200.synthetic import gaussian_chain  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
ok = True
for cross in (False, True):
    args, kw = (None, None, None, 4), {}
    if rank == 0:
        c, _ = gaussian_chain(200003, 8, seed=3, nextra=1)
        if cross:
            args, kw = (c[:90001, 2:], -c[:90001, 1], c[:90001, 0], 4), {'ref': c[90001:, 2:], 'ndim': 8}
        else:
            args, kw = (c[:, 2:], -c[:, 1], c[:, 0], 4), {'ndim': 8}
        kw['priorvolume'] = 3.0
    out = evidence_sharded(*args, **kw)
    if rank == 0:
        q, l, w = args[0], args[1], args[2]
        one = _lib.default_context(local).evidence(q, l, w, 4, ref=kw.get('ref'), ndim=8, log_sum_w=math.log(w.sum()),
                                                   log_prior_volume=math.log(3.0))
        err = float(np.abs(one['logE'] - out['logE']).max())
        print('cross=%d world=%d logE=%s |sharded - single|=%.2e' % (cross, world, out['logE'], err), flush=True)
        ok &= err < 1e-11
    gathered = [None] * world
    dist.all_gather_object(gathered, [float(x) for x in out['logE']])
    ok &= all(g == gathered[0] for g in gathered)
    # the same with the chain on every rank's host: each rank uploads its own block of rows, the blocks are all-gathered
    c, _ = gaussian_chain(200003, 8, seed=3, nextra=1)
    if cross:
        args2, kw2 = (c[:90001, 2:], -c[:90001, 1], c[:90001, 0], 4), {'ref': c[90001:, 2:], 'ndim': 8}
    else:
        args2, kw2 = (c[:, 2:], -c[:, 1], c[:, 0], 4), {'ndim': 8}
    out2 = evidence_sharded(*args2, priorvolume=3.0, host_on_every_rank=True, **kw2)
    same = bool(np.array_equal(out2['logE'], out['logE']))
    if rank == 0:
        print('cross=%d host arrays on every rank: identical to the broadcast path: %s' % (cross, same), flush=True)
    ok &= same
# chain-per-GPU batching: every rank passes the same list, results gathered
chains = [gaussian_chain(n, d, seed=40 + i, integer_weights=True, inflate=1.0)[0]
          for i, (n, d) in enumerate([(3000, 6), (9000, 7), (5000, 8), (2000, 9), (7000, 7), (4000, 6)])]
loge, var, ms = evidence_many(chains, kmax=2, device=local)
if rank == 0:
    ctx = _lib.default_context(local)
    for c, le in zip(chains, loge):
        one = ctx.evidence(c[:, 2:], -c[:, 1], c[:, 0], 2, log_sum_w=math.log(c[:, 0].sum()))
        ok &= bool(np.abs(one['logE'] - le).max() < 1e-12)
    print('DIST_CHECK', 'OK' if ok else 'FAILED', flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
