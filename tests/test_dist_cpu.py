"""Row sharding over ranks: world_size-2 gloo process group on the CPU.  The per-rank CUDA call is
replaced by the oracle (TEST-ONLY) so that the broadcast / shard / all-reduce logic is what is tested."""
import math
import os
import socket

import numpy as np
import pytest

from mcevidence_b200.dist import shard_range
from mcevidence_b200.grid import assign_largest_first


def test_shard_range_partitions_rows():
    for n in (1, 7, 100, 1000003):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_assign_largest_first_balances():
    sizes = [5917, 98886, 15049, 15049, 26862, 6778, 40000, 12000, 30000]
    buckets = assign_largest_first(sizes, 4)
    assert sorted(i for b in buckets for i in b) == list(range(len(sizes)))
    assert buckets[0][0] == 1                     # the largest chain goes first
    load = [sum(sizes[i] ** 2 for i in b) for b in buckets]
    assert max(load) == sizes[1] ** 2             # nothing is stacked on top of the dominant problem


def _oracle_compute(arrays, q_range):
    from oracle import mce_oracle as O
    x = arrays['samples']
    ref = arrays.get('ref')
    d = x.shape[1]
    cs = O.covariance_stats(x if ref is None else np.concatenate((x, ref)))
    y = O.whiten(x, cs['eVec'], cs['eVal'])
    lo, hi = q_range
    kmax = 4
    if ref is None:
        dk, _ = O.knn_brute_fp64(y, y[lo:hi], kmax, exclude_self=True, self_offset=lo)
        cols = range(0, kmax - 1)
    else:
        yr = O.whiten(ref, cs['eVec'], cs['eVal'])
        dk, _ = O.knn_brute_fp64(yr, y[lo:hi], kmax + 1)
        cols = range(1, kmax)
    parts = []
    for c in cols:
        t = O.logsumexp_terms(dk[:, c], arrays['logl'][lo:hi], arrays['weights'][lo:hi], d)
        m = t.max()
        parts.append([m, np.exp(t - m).sum(), np.exp(2 * (t - m)).sum()])
    return np.array(parts), math.log(cs['J']), {}


def _worker(rank, world, port, cross, q, replicated=False):
    import torch.distributed as dist
    from mcevidence_b200.dist import evidence_sharded
    from mcevidence_b200.synthetic import gaussian_chain
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        args = (None, None, None, 4)
        kw = {}
        if rank == 0 or replicated:
            c, _ = gaussian_chain(901 if replicated else 900, 5, seed=7)
            if cross:
                args = (c[:400, 2:], -c[:400, 1], c[:400, 0], 4)
                kw = {'ref': c[400:, 2:]}
            else:
                args = (c[:, 2:], -c[:, 1], c[:, 0], 4)
            kw['priorvolume'] = 2.0
        out = evidence_sharded(*args, compute=_oracle_compute, host_on_every_rank=replicated, **kw)
        q.put((rank, out['logE'], out['q_range']))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('cross', [False, True])
def test_sharded_evidence_two_ranks_gloo(cross):
    import torch.multiprocessing as mp
    from oracle import mce_oracle as O
    from mcevidence_b200.synthetic import gaussian_chain
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cross, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c, _ = gaussian_chain(900, 5, seed=7)
    if cross:
        want = O.evidence(c[:400, 2:], -c[:400, 1], c[:400, 0], 5, 4, s2=c[400:, 2:], priorvolume=2.0, knn='brute')
        assert res[0][2] == (0, 200) and res[1][2] == (200, 400)
    else:
        want = O.evidence(c[:, 2:], -c[:, 1], c[:, 0], 5, 4, priorvolume=2.0, knn='brute')
        assert res[0][2] == (0, 450) and res[1][2] == (450, 900)
    np.testing.assert_array_equal(res[0][1], res[1][1])          # every rank ends with the same answer
    assert np.abs(res[0][1] - want).max() < 1e-10


@pytest.mark.parametrize('cross', [False, True])
def test_sharded_evidence_host_arrays_on_every_rank(cross):
    """Every rank holds the chain (as the ranks of the reference's MPI driver do): each uploads its own block of rows, the
    blocks are all-gathered (odd row count: the last block is padded) and the answer equals the single-process one."""
    import torch.multiprocessing as mp
    from oracle import mce_oracle as O
    from mcevidence_b200.synthetic import gaussian_chain
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cross, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c, _ = gaussian_chain(901, 5, seed=7)
    if cross:
        want = O.evidence(c[:400, 2:], -c[:400, 1], c[:400, 0], 5, 4, s2=c[400:, 2:], priorvolume=2.0, knn='brute')
    else:
        want = O.evidence(c[:, 2:], -c[:, 1], c[:, 0], 5, 4, priorvolume=2.0, knn='brute')
        assert res[0][2] == (0, 451) and res[1][2] == (451, 901)
    np.testing.assert_array_equal(res[0][1], res[1][1])
    assert np.abs(res[0][1] - want).max() < 1e-10
