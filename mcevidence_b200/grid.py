"""Many independent chains: chain-per-GPU batching.

The reference analyses the Planck grid with MPI ranks over data sets
(``planck_mcevidence.py:149-160, 306-359``): 175 (or 700) independent small evidence problems,
no exchange except a final gather of the result tables (``:436``).  Here the problems are
ordered largest-first and dealt round-robin to the GPUs of the box; each GPU runs its share
back to back on one context, and the per-problem results are gathered with one object
all-gather when a process group is up.  No data-path collective is involved.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib


def assign_largest_first(sizes, world):
    """Indices of the problems each rank runs: sorted by decreasing cost (n^2), dealt greedily to
    the least-loaded rank (the reference's static split is per data set, ``planck_mcevidence.py:149-156``)."""
    order = np.argsort(-np.asarray(sizes, dtype=np.float64) ** 2, kind='stable')
    load = [0.0] * world
    buckets = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        buckets[r].append(int(i))
        load[r] += float(sizes[i]) ** 2
    return buckets


def evidence_many(chains, kmax=2, ndim=None, priorvolume=1.0, device=0, ctx=None, group=None, nstreams=8):
    """Auto-mode log-evidence of every chain in ``chains`` (arrays ``[weight, -logL, theta...]``).

    ``ndim`` / ``priorvolume`` may be scalars or per-chain sequences.  Single process: all chains run
    on ``device``.  With an initialised ``torch.distributed`` group every rank must pass the same
    list; rank r computes its share and all ranks return the full list of ``logE`` arrays.
    A rank's share goes to the library in ONE ``mce_evidence_batch`` call, which overlaps the problems on
    ``nstreams`` internal streams ("batch_lanes"), so that the kernels of small chains -- which cannot fill
    148 SMs on their own -- run side by side.
    """
    n = len(chains)
    ndims = [ndim] * n if (ndim is None or np.isscalar(ndim)) else list(ndim)
    pvs = [priorvolume] * n if np.isscalar(priorvolume) else list(priorvolume)
    rank, world = 0, 1
    dist = None
    if group is not None or _dist_ready():
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = assign_largest_first([len(c) for c in chains], world)[rank]
    def prepare(i):
        # One strided pass over the chain for both leading columns (weight, -logL); the parameter block stays a view and
        # is packed by the library's staging threads.  numpy releases the GIL in these copies, so chains are prepared
        # concurrently.
        c = np.asarray(chains[i], dtype=np.float64)
        nd = c.shape[1] - 2 if ndims[i] is None else int(ndims[i])
        wl = np.ascontiguousarray(c[:, :2].T)
        np.negative(wl[1], out=wl[1])
        return {'query': c[:, 2:], 'logl': wl[1], 'weights': wl[0], 'kmax': kmax, 'ndim': nd,
                'log_sum_w': math.log(np.sum(wl[0])), 'log_prior_volume': math.log(pvs[i])}

    if len(mine) > 8:
        probs = list(_prep_pool().map(prepare, mine))
    else:
        probs = [prepare(i) for i in mine]
    ctx = ctx or _lib.default_context(device)
    ctx.set_option('batch_lanes', max(1, min(int(nstreams), 32)))
    outs = ctx.evidence_batch(probs) if probs else []              # one C call: mce_evidence_batch
    local = {i: (o['logE'], o['var_logE'], o['stats']['ms_total']) for i, o in zip(mine, outs)}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, local, group=group)
        local = {}
        for g in gathered:
            local.update(g)
    return [local[i][0] for i in range(n)], [local[i][1] for i in range(n)], [local[i][2] for i in range(n)]


_PREP_POOL = None


def _prep_pool():
    global _PREP_POOL
    if _PREP_POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _PREP_POOL = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1), thread_name_prefix='mce-prep')
    return _PREP_POOL


def _dist_ready():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized()
    except Exception:
        return False
