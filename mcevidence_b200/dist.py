"""Row-sharded evidence over the GPUs of one box: one process per GPU, ``torch.distributed``.

The neighbour search shards naturally over query rows: every rank
holds the whole sample set (broadcast from rank 0 over NCCL / NVLink), whitens it identically,
searches only its own contiguous block of query rows against the full reference set and reduces
its rows to per-k triples ``(max_t, sum exp(t-max), sum exp(2(t-max)))``.  The only exchange
after the broadcast is an all-gather of those ``3*(kmax-1)`` doubles per rank (combined on the
host: maximum of the maxima, sum of the rescaled sums), so the collective cost is launch latency,
not bandwidth.

The reference has no multi-GPU path; its only parallelism is MPI over independent data sets
(``planck_mcevidence.py:149-160``), which ``mcevidence_b200.grid`` covers chain-per-GPU.

torch is used here for device memory, streams and the process group only.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib


def shard_range(n, rank, world):
    """Contiguous block of rows for ``rank``: sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_partials(partials, group=None, device=None):
    """Combine per-rank ``(nk, 3)`` triples ``(max, sum, sum2)`` into the triples of all rows, identical on every rank.

    One all-gather of the ``3*nk`` doubles of every rank, then ``mce_combine_partials`` on the host (maximum of the
    maxima, sums rescaled and added in rank order: every rank does the same arithmetic, so the result is bit-identical
    everywhere).  An all-reduce MAX followed by an all-reduce SUM of the rescaled sums is the same exchange in two
    collectives and half a dozen small kernels; at 8 GPUs that cost more than the reduction kernel itself.
    Works on CPU tensors (gloo, used by the CPU tests) and CUDA tensors (nccl)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.as_tensor(np.ascontiguousarray(partials, dtype=np.float64)).reshape(-1)
    if device is not None:
        mine = mine.to(device)
    if device is not None:
        allp = torch.empty((world, mine.numel()), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(allp, mine, group=group)
        allp = allp.cpu().numpy()
    else:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        allp = torch.stack(parts).numpy()
    nk = np.asarray(partials).shape[0]
    return _lib.combine_partials(allp.reshape(world, nk, 3))


def evidence_sharded(samples, logl, weights, kmax, ref=None, ndim=None, priorvolume=1.0, sum_w=None,
                     group=None, device=None, compute=None, ctx=None, resident=None, host_on_every_rank=False):
    """Log-evidence with query rows sharded over the ranks of ``group``.

    Rank 0 passes the host arrays (other ranks may pass ``None``); every rank returns the same
    ``dict(logE, var_logE, stats)``.  ``host_on_every_rank``: EVERY rank passes the same host arrays (each process read
    the chain files itself, as the ranks of the reference's MPI driver do, ``planck_mcevidence.py:149-160``); a rank then
    uploads only its own block of rows over its own PCIe link and the blocks are all-gathered over NVLink instead of
    one upload on rank 0 followed by a broadcast.  ``compute(arrays, q_range) -> (partials, log_jacobian, stats)``
    replaces the CUDA call in the CPU tests of the sharding/combination logic.  ``resident``: a dict
    of CUDA tensors already broadcast (``samples``, ``logl``, ``weights``, optional ``ref``) to skip
    the host staging (used by ``bench.py`` for the inputs-resident measurement).
    """
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    use_cuda = compute is None
    if use_cuda and device is None:
        device = torch.device('cuda', torch.cuda.current_device())

    if resident is None:
        # ---- shapes, then the arrays, from rank 0 -------------------------------------------
        meta = torch.zeros(6, dtype=torch.float64)
        if rank == 0:
            samples = np.asarray(samples, dtype=np.float64)
            nd = samples.shape[1] if ndim is None else int(ndim)
            sw = float(np.sum(weights)) if sum_w is None else float(sum_w)
            meta = torch.tensor([samples.shape[0], nd, 0 if ref is None else ref.shape[0], int(kmax), sw,
                                 float(priorvolume)], dtype=torch.float64)
        if use_cuda:
            meta = meta.to(device)
        dist.broadcast(meta, src=0, group=group)
        meta = meta.cpu()
        n, nd, n_ref, kmax = int(meta[0]), int(meta[1]), int(meta[2]), int(meta[3])
        sw, priorvolume = float(meta[4]), float(meta[5])

        def bcast(arr, shape):
            if rank == 0:
                t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64).reshape(shape))
                t = t.to(device, non_blocking=False) if use_cuda else t.clone()
            else:
                t = torch.empty(shape, dtype=torch.float64, device=device if use_cuda else 'cpu')
            dist.broadcast(t, src=0, group=group)
            return t

        def gather_rows(arr, cols):
            """Every rank holds ``arr``: upload rows [rank*c, (rank+1)*c) and all-gather the blocks."""
            c = -(-n_rows[0] // world)
            lo, hi = min(rank * c, n_rows[0]), min((rank + 1) * c, n_rows[0])
            shape = (world * c, cols) if cols else (world * c,)
            full = torch.empty(shape, dtype=torch.float64, device=device if use_cuda else 'cpu')
            mine = full[rank * c:(rank + 1) * c]
            if hi > lo:
                src = torch.from_numpy(np.ascontiguousarray(arr[lo:hi], dtype=np.float64))
                mine[:hi - lo].copy_(src, non_blocking=True)
            if hi - lo < c:
                mine[hi - lo:].zero_()
            if use_cuda:
                dist.all_gather_into_tensor(full, mine, group=group)
            else:
                parts = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine.clone(), group=group)
                full = torch.cat(parts)
            return full[:n_rows[0]]

        n_rows = [n]
        if host_on_every_rank and world > 1:
            samples = np.asarray(samples)
            tens = {'samples': gather_rows(samples[:, :nd], nd), 'logl': gather_rows(logl, 0), 'weights': gather_rows(weights, 0)}
            if n_ref:
                n_rows[0] = n_ref
                tens['ref'] = gather_rows(np.asarray(ref)[:, :nd], nd)
        else:
            tens = {'samples': bcast(samples[:, :nd] if rank == 0 else None, (n, nd)),
                    'logl': bcast(logl, (n,)), 'weights': bcast(weights, (n,))}
            if n_ref:
                tens['ref'] = bcast(ref[:, :nd] if rank == 0 else None, (n_ref, nd))
    else:
        tens = resident
        n, nd = tens['samples'].shape
        n_ref = tens['ref'].shape[0] if tens.get('ref') is not None else 0
        sw = float(sum_w)
    q_range = shard_range(n, rank, world)

    # ---- this rank's rows ----------------------------------------------------------------------
    if compute is not None:
        partials, log_jac, stats = compute({k: v.numpy() for k, v in tens.items()}, q_range)
    else:
        ctx = ctx or _lib.default_context(device.index)
        ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)
        r = tens.get('ref')
        out = ctx.evidence_device(tens['samples'].data_ptr(), n, tens['samples'].stride(0),
                                  tens['logl'].data_ptr(), tens['weights'].data_ptr(), nd, kmax,
                                  ref_ptr=r.data_ptr() if r is not None else 0, n_ref=n_ref,
                                  ld_ref=r.stride(0) if r is not None else 0, q_range=q_range,
                                  want_partials=True, want_logE=False)
        partials, log_jac, stats = out['partials'], out['log_jacobian'], out['stats']

    # ---- exchange: 3*(kmax-1) doubles ------------------------------------------------------------
    total = allreduce_partials(partials, group=group, device=device if use_cuda else None)
    loge, var = _lib.finalize(total, nd, n_ref > 0, n, math.log(sw), log_jac, math.log(priorvolume))
    return {'logE': loge, 'var_logE': var, 'stats': stats, 'q_range': q_range, 'partials': total}
