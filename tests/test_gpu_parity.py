"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the frozen reference outputs.

Tolerances.  Neighbour indices and distances on identical coordinates (identity whitening): bit-exact.
Whitened pipeline: r_k within 1e-12 relative of the exact FP64 oracle, logE within 1e-6 absolute of
the reference (the north-star tolerance; the observed differences are ~1e-13, asserted at 1e-9).
"""
import math
import os

import numpy as np
import pytest

from oracle import mce_oracle as O
from mcevidence_b200 import _lib
from mcevidence_b200.synthetic import gaussian_chain

pytestmark = pytest.mark.gpu

LOGE_TOL = 1e-6          # north-star tolerance vs the reference
LOGE_TIGHT = 1e-9        # what the implementation actually meets


def oracle_knn(x, k, ref=None):
    if ref is not None:
        return O.knn_brute_fp64(ref, x, k)
    d, i = O.knn_brute_fp64(x, x, k - 1, exclude_self=True)
    return (np.concatenate((np.zeros((len(x), 1)), d), axis=1),
            np.concatenate((np.arange(len(x))[:, None], i), axis=1))


@pytest.mark.parametrize('n,d,k,cross', [
    (2, 1, 2, False), (5, 2, 3, False), (129, 1, 2, False), (128, 3, 4, True), (2000, 5, 5, False),
    (3000, 10, 5, False), (3000, 10, 9, True), (2500, 17, 4, False), (2000, 27, 5, True), (1000, 40, 3, False),
    (1500, 64, 2, False), (4097, 6, 17, False), (20000, 8, 3, False)])
def test_knn_bit_exact(ctx, n, d, k, cross):
    """(DkNN, indices) contract of kneighbors (MCEvidence.py:1104): bit-exact against the FP64 oracle."""
    rng = np.random.default_rng(1000 * d + n)
    x = rng.normal(size=(n, d)) * (1 + rng.random(d)) + rng.normal(size=d)
    ref = rng.normal(size=(n + 37, d)) * (1 + rng.random(d)) if cross else None
    dk, ix, st = ctx.knn(x, k, ref=ref, identity=True)
    od, oi = oracle_knn(x, k, ref)
    np.testing.assert_array_equal(ix, oi)
    np.testing.assert_array_equal(dk, od)
    assert st['launches'] >= 4


@pytest.mark.skipif(os.environ.get('MCE_FILTER_ENGINE') == '1', reason='tile options apply to the FFMA engine only')
@pytest.mark.parametrize('d', [1, 2, 3, 16])
def test_wide_tile_at_low_ndim(ctx, d):
    """The wide query tile (TQ=4, two-row prefetch; the default for large inputs at ndim <= 16, forced here):
    spot rows against an exhaustive numpy scan, every row against the exact FP64 kernel on a query shard."""
    n = 400000
    rng = np.random.default_rng(50 + d)
    x = rng.normal(size=(n, d))
    try:
        for name, v in (('filter_tq', 4), ('filter_tr', 8), ('filter_packed', 1), ('filter_sync', 3)):
            ctx.set_option(name, v)
        dk, ix, st = ctx.knn(x, 4, identity=True)
    finally:
        for name in ('filter_tq', 'filter_tr', 'filter_packed', 'filter_sync'):
            ctx.set_option(name, -1)
    assert st['filter_tq'] == 4 and st['filter_mode'] == 3
    assert np.all(np.diff(dk, axis=1) >= 0) and np.array_equal(ix[:, 0], np.arange(n))
    for r in rng.choice(n, 40, replace=False):
        d2 = O.sq_dist_sequential(x[r:r + 1], x)[0]
        d2[r] = np.inf
        nn = np.lexsort((np.arange(n), d2))[:3]
        assert np.array_equal(ix[r, 1:], nn)
        np.testing.assert_array_equal(dk[r, 1:], np.sqrt(d2[nn]))
    lo, hi = 123456, 123456 + 3000
    ek, ei, _ = ctx.knn(x, 4, identity=True, q_range=(lo, hi), exact_only=True)
    np.testing.assert_array_equal(ix[lo:hi], ei)
    np.testing.assert_array_equal(dk[lo:hi], ek)


def test_knn_exact_rescan_path(ctx):
    """The exhaustive FP64 kernel (used for rows the FP32 filter cannot certify) on its own."""
    rng = np.random.default_rng(3)
    x = rng.normal(size=(700, 6))
    for ref in (None, rng.normal(size=(650, 6))):
        dk, ix, st = ctx.knn(x, 6, ref=ref, identity=True, exact_only=True)
        od, oi = oracle_knn(x, 6, ref)
        np.testing.assert_array_equal(ix, oi)
        np.testing.assert_array_equal(dk, od)
        assert st['rows_rescanned'] == 700


def test_duplicates_and_ties(ctx):
    """Exact ties: duplicates are neighbours at distance 0, order by index, rows re-scanned exactly."""
    rng = np.random.default_rng(5)
    x = np.repeat(rng.normal(size=(500, 4)), 4, axis=0)
    dk, ix, st = ctx.knn(x, 6, identity=True)
    od, oi = oracle_knn(x, 6)
    np.testing.assert_array_equal(dk, od)
    np.testing.assert_array_equal(ix, oi)
    assert np.all(dk[:, 1:4] == 0.0) and np.all(dk[:, 4] > 0)
    # integer lattice: massive ties at equal distances
    g = np.stack(np.meshgrid(*[np.arange(9.0)] * 3), -1).reshape(-1, 3)
    dk, ix, st = ctx.knn(g, 8, identity=True)
    od, oi = oracle_knn(g, 8)
    np.testing.assert_array_equal(ix, oi)
    np.testing.assert_array_equal(dk, od)
    assert st['rows_rescanned'] > 0


def test_uncentred_far_from_origin(ctx):
    """FP32 filter on data far from the origin (no centring with an explicit W): stays exact via verification."""
    rng = np.random.default_rng(8)
    x = rng.normal(size=(3000, 5)) * 1e-3 + 50.0
    dk, ix, st = ctx.knn(x, 4, identity=True)
    od, oi = oracle_knn(x, 4)
    np.testing.assert_array_equal(ix, oi)
    np.testing.assert_array_equal(dk, od)
    assert st['rows_rescanned'] > 0        # the FP32 scores cannot separate neighbours here


def test_query_row_ranges(ctx):
    rng = np.random.default_rng(9)
    x = rng.normal(size=(1000, 7))
    od, oi = oracle_knn(x, 5)
    for lo, hi in [(0, 1000), (0, 1), (250, 777), (999, 1000)]:
        dk, ix, _ = ctx.knn(x, 5, identity=True, q_range=(lo, hi))
        np.testing.assert_array_equal(ix, oi[lo:hi])
        np.testing.assert_array_equal(dk, od[lo:hi])


def test_strided_rows_and_ndim_prefix(ctx):
    """First ndim columns of wider rows (MCEvidence.py:902) without a host copy."""
    rng = np.random.default_rng(10)
    wide = rng.normal(size=(800, 11))
    dk, ix, _ = ctx.knn(wide, 4, ndim=6, W=np.eye(6))
    od, oi = oracle_knn(np.ascontiguousarray(wide[:, :6]), 4)
    np.testing.assert_array_equal(ix, oi)
    np.testing.assert_array_equal(dk, od)


def test_very_wide_rows_use_pitched_copy(ctx):
    """ld > 16 d: the host copy packs the first ndim columns (cudaMemcpy2D) instead of moving whole rows."""
    rng = np.random.default_rng(11)
    wide = rng.normal(size=(1200, 40))
    for arr in (wide, wide.astype(np.float32)):
        dk, ix, _ = ctx.knn(arr, 3, ndim=2, W=np.eye(2))
        od, oi = oracle_knn(np.ascontiguousarray(arr[:, :2]).astype(np.float64), 3)
        np.testing.assert_array_equal(ix, oi)
        np.testing.assert_array_equal(dk, od)


@pytest.mark.parametrize('n,d,kmax,cross', [(4000, 5, 4, False), (6000, 10, 4, False), (6000, 10, 4, True),
                                            (5000, 20, 8, False), (4000, 27, 4, True), (3000, 3, 2, False),
                                            (700, 2, 1, False), (900, 12, 16, False)])
def test_evidence_vs_oracle(ctx, n, d, kmax, cross):
    chain, _ = gaussian_chain(n, d, seed=100 + d, nextra=2, mean_shift=1.3)
    s, logl, w = chain[:, 2:], -chain[:, 1], chain[:, 0]
    if cross:
        h = n // 3
        s1, s2, l1, w1 = s[:h], s[h:], logl[:h], w[:h]
    else:
        s1, s2, l1, w1 = s, None, logl, w
    want, parts = O.evidence(s1, l1, w1, d, kmax, s2=s2, priorvolume=3.5, knn='brute', return_parts=True)
    got = ctx.evidence(s1, l1, w1, kmax, ref=s2, ndim=d, log_sum_w=math.log(w1.sum()),
                       log_prior_volume=math.log(3.5), want_knn=True)
    assert got['logE'].shape == (max(2, kmax) - 1,)
    assert np.abs(got['logE'] - want).max() < LOGE_TIGHT
    # kneighbors-shaped arrays: identical neighbours, r_k to 1e-12 relative (Cholesky vs eig whitening)
    np.testing.assert_array_equal(got['idx'], parts['idx'])
    np.testing.assert_allclose(got['dknn'], parts['dknn'], rtol=1e-12, atol=0)
    assert abs(got['log_jacobian'] - math.log(parts['cov']['J'])) < 1e-12
    # variance: pinned by the oracle's definition only (no reference counterpart)
    k0 = parts['k0']
    _, ovar = O.evidence_from_dknn_log(parts['dknn'], l1, w1, d, max(2, kmax), k0, w1.sum(),
                                       math.log(parts['cov']['J']), math.log(3.5))
    np.testing.assert_allclose(got['var_logE'], ovar[1:], rtol=1e-9)


def test_explicit_whitening_matrices(ctx):
    """Reference factor eVec/sqrt(eVal) passed explicitly, and a different W for the ref set."""
    chain, _ = gaussian_chain(3000, 6, seed=4, mean_shift=0.5)
    s1, s2 = chain[:1400, 2:], chain[1400:, 2:]
    l1, w1 = -chain[:1400, 1], chain[:1400, 0]
    cs = O.covariance_stats(np.concatenate((s1, s2)))
    W = cs['eVec'] / np.sqrt(cs['eVal'])[None, :]
    want = O.evidence(s1, l1, w1, 6, 3, s2=s2, knn='brute')
    got = ctx.evidence(s1, l1, w1, 3, ref=s2, W=W, mu=np.concatenate((s1, s2)).mean(0),
                       log_jacobian=math.log(cs['J']), log_sum_w=math.log(w1.sum()))
    assert np.abs(got['logE'] - want).max() < LOGE_TIGHT
    cs2 = O.covariance_stats(s2)
    W2 = cs2['eVec'] / np.sqrt(cs2['eVal'])[None, :]
    dk, _ = O.knn_brute_fp64(s2 @ W2, s1 @ W, 4)
    want2 = O.evidence_from_dknn_linear(dk, l1, w1, 6, 3, 0, w1.sum(), cs['J'], 0.0)[1:]
    got2 = ctx.evidence(s1, l1, w1, 3, ref=s2, W=W, W_ref=W2, log_jacobian=math.log(cs['J']),
                        log_sum_w=math.log(w1.sum()))
    assert np.abs(got2['logE'] - want2).max() < LOGE_TIGHT


@pytest.mark.parametrize('cross', [False, True])
def test_partials_shards_combine_to_full(ctx, cross):
    """Row sharding: per-shard triples combined on the host equal the one-shot result (auto and cross)."""
    chain, _ = gaussian_chain(9000, 8, seed=12)
    s, l, w = chain[:5000, 2:], -chain[:5000, 1], chain[:5000, 0]
    ref = chain[5000:, 2:] if cross else None
    full = ctx.evidence(s, l, w, 4, ref=ref, log_sum_w=math.log(w.sum()), want_knn=True)
    parts = []
    for lo, hi in [(0, 1700), (1700, 1701), (1701, 5000)]:
        o = ctx.evidence(s, l, w, 4, ref=ref, q_range=(lo, hi), want_partials=True, want_knn=True,
                         log_sum_w=math.log(w.sum()))
        parts.append(o['partials'])
        np.testing.assert_array_equal(o['idx'], full['idx'][lo:hi])
        np.testing.assert_array_equal(o['dknn'], full['dknn'][lo:hi])
    tot = _lib.combine_partials(np.array(parts))
    loge, var = _lib.finalize(tot, 8, cross, 5000, math.log(w.sum()), full['log_jacobian'], 0.0)
    assert np.abs(loge - full['logE']).max() < 1e-12
    np.testing.assert_allclose(var, full['var_logE'], rtol=1e-9)


def test_runs_on_caller_stream(ctx):
    """set_stream: work is ordered on the caller's stream (0 = legacy default stream), None restores."""
    import torch
    chain, _ = gaussian_chain(3000, 5, seed=14)
    s, l, w = chain[:, 2:], -chain[:, 1], chain[:, 0]
    base = ctx.evidence(s, l, w, 3, log_sum_w=0.0)['logE']
    side = torch.cuda.Stream()
    try:
        for handle in (torch.cuda.current_stream().cuda_stream, side.cuda_stream):
            ctx.set_stream(handle)
            dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in (('s', s), ('l', l), ('w', w))}
            with torch.cuda.stream(side if handle == side.cuda_stream else torch.cuda.current_stream()):
                dev['s'].mul_(1.0)          # queued on the same stream right before the library call
                out = ctx.evidence_device(dev['s'].data_ptr(), 3000, 5, dev['l'].data_ptr(), dev['w'].data_ptr(), 5, 3)
            np.testing.assert_array_equal(out['logE'], base)
    finally:
        ctx.set_stream(None)


def test_float32_sample_arrays(ctx):
    """float32 samples through the ABI (MCE_FLAG_F32_SAMPLES): widened on load, so the results equal those of
    the widened float64 copy bit for bit -- auto and cross, contiguous and strided."""
    chain, _ = gaussian_chain(5000, 6, seed=15, nextra=2)
    x32 = chain[:, 2:].astype(np.float32)
    l, w = -chain[:, 1], chain[:, 0]
    for ref32 in (None, x32[3000:]):
        q32 = x32 if ref32 is None else x32[:3000]
        lq, wq = (l, w) if ref32 is None else (l[:3000], w[:3000])
        a = ctx.evidence(q32, lq, wq, 4, ref=ref32, ndim=6, log_sum_w=0.0, want_knn=True)
        b = ctx.evidence(q32.astype(np.float64), lq, wq, 4, ref=None if ref32 is None else ref32.astype(np.float64),
                         ndim=6, log_sum_w=0.0, want_knn=True)
        np.testing.assert_array_equal(a['logE'], b['logE'])
        np.testing.assert_array_equal(a['idx'], b['idx'])
        np.testing.assert_array_equal(a['dknn'], b['dknn'])
    dk, ix, _ = ctx.knn(np.ascontiguousarray(x32[:, :6]), 4, identity=True)
    od, oi = oracle_knn(x32[:, :6].astype(np.float64), 4)
    np.testing.assert_array_equal(ix, oi)
    np.testing.assert_array_equal(dk, od)


def test_zero_distance_rows_contribute_nothing(ctx):
    """Duplicated samples have r=0 -> volume 0 (MCEvidence.py:1110): handled as log(0) = -inf terms."""
    chain, _ = gaussian_chain(1500, 4, seed=13)
    chain = np.concatenate((chain, chain[:100]))
    s, l, w = chain[:, 2:], -chain[:, 1], chain[:, 0]
    want = O.evidence(s, l, w, 4, 3, knn='brute')
    got = ctx.evidence(s, l, w, 3, log_sum_w=math.log(w.sum()))
    assert np.all(np.isfinite(got['logE'])) and np.abs(got['logE'] - want).max() < LOGE_TIGHT


def test_argument_errors(ctx):
    x = np.random.default_rng(0).normal(size=(50, 3))
    w = np.ones(50)
    with pytest.raises(_lib.MCEError, match='n_neighbors <= n_samples_fit'):
        ctx.knn(x[:4], 6, identity=True)                     # sklearn raises the same way
    with pytest.raises(_lib.MCEError, match='neighbours'):
        ctx.evidence(x, w, w, 40)                            # kmax beyond MCE_MAX_NEIGHBOURS
    with pytest.raises(_lib.MCEError, match='ndim'):
        ctx.knn(np.zeros((10, 70)), 2, identity=True)
    with pytest.raises(_lib.MCEError, match='row range'):
        ctx.knn(x, 2, identity=True, q_range=(10, 60))
    degenerate = np.column_stack((x[:, 0], np.ones(50), x[:, 1]))    # a constant column: singular covariance
    with pytest.raises(_lib.MCEError) as ei:
        ctx.evidence(degenerate, w, w, 2)
    assert ei.value.code == _lib.ERR_NOT_POSDEF
    # a failed factorisation must not leave the GPU scanning garbage: a large degenerate input fails fast
    import time
    big = np.random.default_rng(1).normal(size=(300000, 4))
    big[:, 2] = 7.0
    t0 = time.perf_counter()
    with pytest.raises(_lib.MCEError):
        ctx.evidence(big, np.zeros(300000), np.ones(300000), 3)
    assert time.perf_counter() - t0 < 5.0
    # the context survives errors
    ctx.knn(x, 2, identity=True)


def test_golden_reference_outputs(ctx, golden, tmp_path):
    """End to end through the drop-in class against the frozen outputs of the unmodified reference."""
    import MCEvidence as M
    g = golden
    np.random.seed(g.np_seed)
    mce = M.MCEvidence(g.method(tmp_path), verbose=0, **g.ctor)
    mle = mce.evidence(**g.ev)
    assert mle.shape == g.mle.shape
    err = np.abs(mle - g.mle).max()
    # the ill-conditioned case is limited by the reference's own eig() accuracy, not by this path
    assert err < (LOGE_TOL if g.name == 'auto_d6_scaled' else LOGE_TIGHT), err
    assert mce.last_stats['launches'] >= 6      # K1..K4 (+ K0 when the library estimates the covariance)


def test_golden_knn_arrays(ctx):
    """DkNN / indices captured inside the reference (sklearn kd-tree) vs mce_evidence(want_knn)."""
    from conftest import Golden
    g = Golden('auto_d8_kd')
    c = g.chains[0]
    out = ctx.evidence(c[:, 2:], -c[:, 1], c[:, 0], g.kmax, ndim=g.ndim, log_sum_w=math.log(c[:, 0].sum()),
                       log_prior_volume=math.log(7.0), want_knn=True)
    np.testing.assert_array_equal(out['idx'], g.raw['idx'])
    np.testing.assert_allclose(out['dknn'], g.raw['dknn'], rtol=1e-12, atol=1e-300)
    assert np.abs(out['logE'] - g.mle).max() < LOGE_TIGHT


def _full_size_properties(ctx, n, d, kmax, cross, seed, nspot=48, nshard=3000, analytic_tol=None):
    """A BASELINE configuration at full size through size-independent properties: sortedness and index ranges, spot
    rows against an exhaustive FP64 numpy scan in the library's own whitened coordinates, a shard of rows against the
    library's exhaustive FP64 search (bit for bit), permutation invariance, and shard combination."""
    chain, info = gaussian_chain(n * (2 if cross else 1), d, seed=seed)
    s, l, w = np.ascontiguousarray(chain[:, 2:]), -chain[:, 1], chain[:, 0]
    s1, l1, w1 = s[:n], l[:n], w[:n]
    s2 = s[n:] if cross else None
    lsw = math.log(w1.sum())
    out = ctx.evidence(s1, l1, w1, kmax, ref=s2, log_sum_w=lsw, want_knn=True)
    dk, ix = out['dknn'], out['idx']
    n_ref = len(s2) if cross else n
    assert dk.shape == (n, kmax + 1)
    assert np.all(np.diff(dk, axis=1) >= 0) and ix.min() >= 0 and ix.max() < n_ref
    c0 = 0 if cross else 1           # first column holding a true neighbour
    if not cross:
        assert np.all(dk[:, 0] == 0) and np.all(dk[:, 1] > 0)
        assert np.array_equal(ix[:, 0], np.arange(n))
        assert not np.any(ix[:, 1:] == np.arange(n)[:, None])
    if analytic_tol is not None and not cross:
        assert np.abs(out['logE'] - info['lnZ']).max() < analytic_tol
    # spot rows against an exhaustive FP64 numpy scan in the library's own whitened coordinates
    cs = O.covariance_stats(s)
    L = np.linalg.cholesky(cs['cov'])
    mu = s.mean(0)
    y1 = np.linalg.solve(L, (s1 - mu).T).T
    y2 = np.linalg.solve(L, (s2 - mu).T).T if cross else y1
    rows = np.random.default_rng(1).choice(n, nspot, replace=False)
    ncol = kmax + 1 - c0
    for r in rows:
        d2 = ((y2 - y1[r]) ** 2).sum(1)
        if not cross:
            d2[r] = np.inf
        nn = np.argsort(d2, kind='stable')[:ncol]
        assert np.array_equal(ix[r, c0:], nn)
        np.testing.assert_allclose(dk[r, c0:], np.sqrt(d2[nn]), rtol=1e-11)
    # a shard of rows through the library's exhaustive FP64 search: identical to the last bit
    lo = n // 3
    ex = ctx.evidence(s1, l1, w1, kmax, ref=s2, log_sum_w=lsw, want_knn=True, q_range=(lo, lo + nshard), exact_only=True)
    np.testing.assert_array_equal(ex['idx'], ix[lo:lo + nshard])
    np.testing.assert_array_equal(ex['dknn'], dk[lo:lo + nshard])
    # permutation invariance of the estimator (different tiles, chunks and summation order)
    perm = np.random.default_rng(2).permutation(n)
    ref_p = s2[np.random.default_rng(3).permutation(n_ref)] if cross else None
    out2 = ctx.evidence(s1[perm], l1[perm], w1[perm], kmax, ref=ref_p, log_sum_w=lsw)
    assert np.abs(out2['logE'] - out['logE']).max() < 1e-9
    # two half-shards combine to the same answer
    parts = [ctx.evidence(s1, l1, w1, kmax, ref=s2, q_range=r, want_partials=True)['partials']
             for r in ((0, n // 2), (n // 2, n))]
    loge, _ = _lib.finalize(_lib.combine_partials(np.array(parts)), d, cross, n, lsw, out['log_jacobian'], 0.0)
    assert np.abs(loge - out['logE']).max() < 1e-11
    assert out['stats']['rows_rescanned'] < 1000
    return out


@pytest.mark.parametrize('engine', [0, 2])
def test_full_size_properties_cfg2(ctx, engine):
    """BASELINE cfg2 (N=1e6, d=10, kmax=4, auto) at full size, with the FFMA filter and with the automatic choice
    (the tcgen05 filter at this size)."""
    ctx.set_option('filter_engine', engine)
    try:
        out = _full_size_properties(ctx, 1000000, 10, 4, False, seed=0, analytic_tol=0.1)
        assert (out['stats']['filter_mode'] == 16) == (engine == 2)
    finally:
        ctx.set_option('filter_engine', 2)


def test_full_size_properties_cfg3(ctx):
    """BASELINE cfg3 shape (cross evidence, d=27, 1e6 x 1e6, kmax=4) at full size."""
    out = _full_size_properties(ctx, 1000000, 27, 4, True, seed=3)
    assert out['stats']['filter_mode'] == 16


def test_full_size_properties_cfg5_shape(ctx):
    """BASELINE cfg5 shape (auto, d=20, kmax=8: nine columns, candidate lists of ten) at N=2e6, the per-GPU reference-set
    size class of the 8-GPU scaling run."""
    out = _full_size_properties(ctx, 2000000, 20, 8, False, seed=5, nspot=32, nshard=2000)
    assert out['stats']['filter_mode'] == 16


@pytest.mark.parametrize('engine', [0, 1])
def test_random_shapes_bit_exact(engine):
    """Thirty seeded random (n, ndim, k, auto / cross) problems per filter engine, with anisotropic, shifted and partly
    duplicated data: neighbour indices and distances bit for bit against the FP64 oracle (MCEvidence.py:1104)."""
    from mcevidence_b200 import _lib
    ctx = _lib.Context(0)
    ctx.set_option('filter_engine', engine)
    rng = np.random.default_rng(20261017)
    for case in range(30):
        n = int(rng.integers(20, 3500))
        d = int(rng.integers(1, 41))
        cross = bool(rng.integers(0, 2))
        nref = int(rng.integers(20, 3500)) if cross else n
        k = int(rng.integers(2, min(10, nref - 1) + 1))
        scale = np.exp(rng.normal(size=d))
        x = rng.normal(size=(n, d)) * scale + rng.normal(size=d) * 3.0
        ref = (rng.normal(size=(nref, d)) * scale + rng.normal(size=d) * 3.0) if cross else None
        if case % 5 == 0:                      # exact duplicates: ties are ordered by index
            x[n // 2:n // 2 + 5] = x[:5]
        if engine == 1 and d > 62:
            continue
        dk, ix, st = ctx.knn(x, k, ref=ref, identity=True)
        od, oi = oracle_knn(x, k, ref)
        assert np.array_equal(ix, oi), (case, n, d, k, cross)
        assert np.array_equal(dk, od), (case, n, d, k, cross)
        assert (st['filter_mode'] == 16) == (engine == 1)
    ctx.close()
