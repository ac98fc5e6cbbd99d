"""The drop-in Python API and command line on the GPU: same calls a user of the reference makes."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import MCEvidence as M
from oracle import mce_oracle as O
from mcevidence_b200.synthetic import gaussian_chain, write_cosmomc_root
from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_cli_on_cosmomc_root(tmp_path):
    """BASELINE config 1 shape: `MCEvidence.py <root> -k 2 --burn 0.3` (and the -b alias) on a
    CosmoMC-format root (4 chain files + .ranges), cosmological parameters only."""
    root = str(tmp_path / 'base_plikHM_TT_lowTEB')
    names, vol = write_cosmomc_root(root, nchains=4, nrows=1500, seed=3)
    chains = [np.loadtxt('%s_%d.txt' % (root, i + 1)) for i in range(4)]
    burned = np.concatenate([c[int(len(c) * 0.3):] for c in chains])
    want = O.evidence(burned[:, 2:], -burned[:, 1], burned[:, 0], 6, 2, priorvolume=vol, knn='brute')
    got = M.main([root, '-k', '2', '--burn', '0.3', '-vb', '0'])
    assert got.shape == (1,) and abs(got[0] - want[0]) < 1e-9
    got_alias = M.main([root, '-k', '2', '-b', '0.3', '-vb', '0'])
    assert got_alias[0] == got[0]
    # --allparams uses every varied column; -ic picks one chain file
    burned1 = chains[1][int(len(chains[1]) * 0.3):]
    want_all = O.evidence(burned1[:, 2:], -burned1[:, 1], burned1[:, 0], len(names), 3,
                          priorvolume=24.0 ** len(names), knn='brute')
    got_all = M.main([root, '-k', '3', '--burn', '0.3', '--allparams', '-ic', '2', '-vb', '0'])
    assert np.abs(got_all - want_all).max() < 1e-9
    # as a script
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'MCEvidence.py'), root, '-k', '2', '--burn', '0.3'],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert 'ln(B)[k=1] = ' in p.stderr + p.stdout
    logged = float((p.stderr + p.stdout).split('ln(B)[k=1] = ')[1].split()[0])
    assert abs(logged - want[0]) < 1e-9


def test_cross_mode_through_class():
    c, _ = gaussian_chain(6000, 7, seed=21, nextra=3)
    np.random.seed(77)
    m = M.MCEvidence([c], split=True, kmax=4, ndim=7, priorvolume=4.0, verbose=0)
    got = m.evidence()
    np.random.seed(77)
    n = len(c)
    ix = np.random.choice(range(n), size=int(n * 0.5), replace=False)
    nix = np.setxor1d(range(n), ix)
    want = O.evidence(c[ix, 2:], -c[ix, 1], c[ix, 0], 7, 4, s2=c[nix, 2:], priorvolume=4.0, knn='brute')
    assert np.abs(got - want).max() < 1e-9
    assert m.nsample == [3000, 3000]


def test_batches_and_info():
    c, _ = gaussian_chain(3000, 5, seed=22)
    m = M.MCEvidence([c], kmax=3, verbose=0, nbatch=3, brange=[2.5, 3.4], bscale='logpower')
    res, info = m.evidence(info=True)
    assert res.shape == (3, 2) and info['Nsamples'] == '3000'
    cs = O.covariance_stats(c[:, 2:])
    for i, p in enumerate(np.linspace(2.5, 3.4, 3)):
        S = int(10 ** p)
        y = O.whiten(c[:S, 2:], cs['eVec'], cs['eVal'])
        d, _ = O.knn_brute_fp64(y, y, 3, exclude_self=True)
        dk = np.concatenate((np.zeros((S, 1)), d), axis=1)
        want = O.evidence_from_dknn_linear(dk, -c[:S, 1], c[:S, 0], 5, 3, 1, c[:, 0].sum(), cs['J'], 0.0)[1:]
        assert np.abs(res[i] - want).max() < 1e-9


def test_importance_sampling_hook():
    """isfunc rescales adjusted_weights only -> SumW (MCEvidence.py:265-270, 1126; planck_mcevidence.py:344)."""
    c, _ = gaussian_chain(2000, 4, seed=23)
    f = lambda s: 0.5 * ((s[:, 0] - 0.1) / 2.0) ** 2
    a = M.MCEvidence([c], kmax=2, verbose=0).evidence()
    b = M.MCEvidence([c], kmax=2, verbose=0, isfunc=f).evidence()
    aw = c[:, 0] * np.exp(-f(c[:, 2:]))
    assert abs((b - a)[0] - (math.log(aw.sum()) - math.log(c[:, 0].sum()))) < 1e-10


def test_grid_of_chains_batched():
    """Chain-per-GPU batching (planck grid shape): many small independent problems."""
    from mcevidence_b200.grid import evidence_many
    chains = [gaussian_chain(n, d, seed=30 + i, integer_weights=True, inflate=1.0)[0]
              for i, (n, d) in enumerate([(1475, 6), (5917, 7), (3000, 8), (2200, 9), (4000, 7)])]
    loge, var, ms = evidence_many(chains, kmax=2)
    for c, le in zip(chains, loge):
        want = O.evidence(c[:, 2:], -c[:, 1], c[:, 0], c.shape[1] - 2, 2, knn='brute')
        assert np.abs(le - want).max() < 1e-9
    # overlapped on several contexts of the same GPU: same numbers
    many = chains * 4
    a1, _, _ = evidence_many(many, kmax=2, nstreams=1)
    a4, _, _ = evidence_many(many, kmax=2, nstreams=4)
    for x, y in zip(a1, a4):
        np.testing.assert_array_equal(x, y)
    # per-chain ndim / prior volume, and an error inside a batch names the failing problem
    loge2, _, _ = evidence_many(chains[:2], kmax=3, ndim=[4, 5], priorvolume=[2.0, 3.0])
    for c, nd, pv, le in zip(chains[:2], (4, 5), (2.0, 3.0), loge2):
        want = O.evidence(c[:, 2:], -c[:, 1], c[:, 0], nd, 3, priorvolume=pv, knn='brute')
        assert np.abs(le - want).max() < 1e-9
    from mcevidence_b200 import _lib
    bad = np.column_stack((np.ones(50), np.zeros(50), np.ones((50, 2))))      # constant columns: singular covariance
    with pytest.raises(_lib.MCEError, match='problem 1'):
        evidence_many([chains[0], bad], kmax=2, nstreams=1)


def test_planck_grid_driver(tmp_path):
    """The grid driver end to end on a miniature CosmoMC grid: tables, csv files and values vs the oracle."""
    import pandas as pd
    import planck_mcevidence as P
    from mcevidence_b200.synthetic import write_cosmomc_grid
    rootdir, out = str(tmp_path / 'grid'), str(tmp_path / 'out')
    grid = write_cosmomc_grid(rootdir, nchains=2, nrows=700)
    dsets = ['plikHM_TT_lowTEB', 'plikHM_TT_lowTEB_post_BAO']
    P.main(['--rootdir', rootdir, '-o', out, '-k', '2', '-b', '0.3', '-v', '0', '--no-h0prior', '--datasets'] + dsets)
    vol_base = {dd: grid[(dd, 'base')][2] for dd in dsets}
    for dd in dsets:
        df = pd.read_csv(os.path.join(out, 'AllChains', 'csv', 'mce_%s.csv' % dd), index_col=0)
        assert list(df.index) == ['base', 'base_omegak', 'base_mnu']
        assert list(df.columns) == ['AllChains', 'delta_lnE_k1', 'PriorVol', 'ndim', 'N_read', 'N_used']
        for mm in df.index:
            root, names, vol = grid[(dd, mm)]
            chains = [np.loadtxt('%s_%d.txt' % (root, i + 1)) for i in range(2)]
            burned = np.concatenate([c[int(len(c) * 0.3):] for c in chains])
            nd = int(df.loc[mm, 'ndim'])
            want = O.evidence(burned[:, 2:], -burned[:, 1], burned[:, 0], nd, 2, priorvolume=vol / vol_base[dd],
                              knn='brute')
            assert abs(df.loc[mm, 'AllChains'] - want[0]) < 1e-8
            assert np.isclose(df.loc[mm, 'PriorVol'], vol / vol_base[dd])
        assert np.isclose(df['delta_lnE_k1'].max(), 0.0)
    assert os.path.exists(os.path.join(out, 'AllChains', 'mce_planck_fullgrid.txt'))
    assert os.path.exists(os.path.join(out, 'AllChains', 'delta_lnE_all.txt'))
    # per-chain run with the H0 importance weight: Mean / Err columns (planck_mcevidence.py:366-372)
    P.main(['--rootdir', rootdir, '-o', out, '-nc', '2', '-nd', '1', '-v', '0', '--h0prior', '--datasets'] + dsets)
    df = pd.read_csv(os.path.join(out, 'SingleChains', 'csv', 'mce_%s.csv' % dsets[0]), index_col=0)
    assert list(df.columns)[:5] == ['chain1', 'chain2', 'Mean_lnE_k1', 'Err_lnE_k1', 'delta_lnE_k1']
    root = grid[(dsets[0], 'base')][0]
    c = np.loadtxt(root + '_1.txt')
    aw = c[:, 0] * np.exp(-0.5 * ((c[:, 2 + 9] - 73.24) / 1.74) ** 2)
    want = O.evidence(c[:, 2:], -c[:, 1], c[:, 0], 6, 2, adjusted_weights=aw, knn='brute')
    assert abs(df.loc['base', 'chain1'] - want[0]) < 1e-8
    assert np.isclose(df.loc['base', 'Err_lnE_k1'], df.loc['base', ['chain1', 'chain2']].std() / np.sqrt(2.0))


@pytest.mark.parametrize('engine', [0, 1])
@pytest.mark.parametrize('cross', [False, True])
def test_batch_size_sweep_is_bit_identical(engine, cross):
    """N4: mce_evidence_sweep (one upload, one whitening, one filter pass cut at the tile-aligned batch sizes) against
    one mce_evidence call per prefix with the same explicit whitening -- bit for bit (MCEvidence.py:1039-1045)."""
    from mcevidence_b200 import _lib
    n, d, kmax = 5300, 6, 4
    c, _ = gaussian_chain(n + (2100 if cross else 0), d, seed=71, mean_shift=0.2)
    s, logl, w = c[:n, 2:], -c[:n, 1], c[:n, 0]
    ref = c[n:, 2:] if cross else None
    cov = np.cov(c[:, 2:].T)
    W = np.linalg.inv(np.linalg.cholesky(cov)).T.copy()
    mu = c[:, 2:].mean(axis=0)
    logj = 0.5 * np.linalg.slogdet(cov)[1]
    rows = [5300, 37, 1000, 1000, 2511, 256, 4097, 5299, 130]      # unsorted, repeated, below one tile, tile aligned
    ctx = _lib.Context(0)
    ctx.set_option('filter_engine', engine)
    kw = dict(ndim=d, W=W, mu=mu, log_jacobian=logj, log_sum_w=math.log(w.sum()), log_prior_volume=0.3)
    got = ctx.evidence_sweep(s, logl, w, kmax, rows, ref=ref, **kw)
    assert got[0]['stats']['filter_mode'] == (16 if engine else got[0]['stats']['filter_mode'])
    ctx.set_option('sweep_seed', 0)          # one launch, lists restarted per range: the same numbers
    for g, g0 in zip(got, ctx.evidence_sweep(s, logl, w, kmax, rows, ref=ref, **kw)):
        np.testing.assert_array_equal(g['logE'], g0['logE'])
    ctx.set_option('sweep_seed', 1)
    for nb, g in zip(rows, got):
        want = ctx.evidence(s[:nb], logl[:nb], w[:nb], kmax, ref=ref, **kw)
        np.testing.assert_array_equal(g['logE'], want['logE'])
        np.testing.assert_array_equal(g['var_logE'], want['var_logE'])
    # W == NULL: the covariance of ALL rows (covtype='all'), not of the prefix
    got2 = ctx.evidence_sweep(s, logl, w, kmax, [900, n], ref=ref, ndim=d, log_sum_w=math.log(w.sum()))
    full = ctx.evidence(s, logl, w, kmax, ref=ref, ndim=d, log_sum_w=math.log(w.sum()))
    np.testing.assert_array_equal(got2[1]['logE'], full['logE'])
    want = O.evidence(s[:900], logl[:900], w[:900], d, kmax, s2=ref, knn='brute', cov_from=c[:, 2:] if cross else s,
                      sum_w=w.sum())
    assert np.abs(got2[0]['logE'] - want).max() < 1e-9
    # the largest batch may be smaller than the chain: rows beyond it are neither searched nor uploaded (explicit W) ...
    part = ctx.evidence_sweep(s, logl, w, kmax, [700, 2000], ref=ref, **kw)
    for nb, g in zip((700, 2000), part):
        np.testing.assert_array_equal(g['logE'], ctx.evidence(s[:nb], logl[:nb], w[:nb], kmax, ref=ref, **kw)['logE'])
    # ... but still enter the covariance when the library estimates it (covtype='all')
    part2 = ctx.evidence_sweep(s, logl, w, kmax, [800, 2100], ref=ref, ndim=d, log_sum_w=math.log(w.sum()))
    for nb, g in zip((800, 2100), part2):
        want = O.evidence(s[:nb], logl[:nb], w[:nb], d, kmax, s2=ref, knn='brute', cov_from=c[:, 2:] if cross else s,
                          sum_w=w.sum())
        assert np.abs(g['logE'] - want).max() < 1e-9
    with pytest.raises(_lib.MCEError):
        ctx.evidence_sweep(s, logl, w, kmax, [3], ref=None if not cross else ref[:2], **kw)   # fewer rows than neighbours
    ctx.close()


def test_batch_lanes_match_single_calls():
    """mce_evidence_batch overlaps problems on internal streams: results equal n mce_evidence calls bit for bit,
    whatever the number of lanes, and a failing problem is named."""
    from mcevidence_b200 import _lib
    # two cells large enough for the tcgen05 engine (>= 3e8 pairs) run on lanes next to FFMA-engine cells
    shapes = [(1475, 6), (9000, 7), (3000, 8), (2200, 9), (4000, 7), (700, 6), (15000, 6), (5100, 9), (800, 7), (12000, 8),
              (30000, 7), (26000, 9)]
    chains = [gaussian_chain(n, d, seed=130 + i, integer_weights=True, inflate=1.0, nextra=3)[0]
              for i, (n, d) in enumerate(shapes)]
    probs = [{'query': c[:, 2:], 'logl': -c[:, 1], 'weights': c[:, 0], 'kmax': 3, 'ndim': d,
              'log_sum_w': math.log(c[:, 0].sum()), 'log_prior_volume': 0.1 * i}
             for i, (c, (_, d)) in enumerate(zip(chains, shapes))]
    ctx = _lib.Context(0)
    single = [ctx.evidence(p['query'], p['logl'], p['weights'], 3, ndim=p['ndim'], log_sum_w=p['log_sum_w'],
                           log_prior_volume=p['log_prior_volume']) for p in probs]
    assert single[-1]['stats']['filter_mode'] == 16 and single[0]['stats']['filter_mode'] != 16
    for lanes in (1, 3, 8):
        ctx.set_option('batch_lanes', lanes)
        for rep in range(2):                       # second pass: warm workspaces
            outs = ctx.evidence_batch(probs)
            for o, s1 in zip(outs, single):
                np.testing.assert_array_equal(o['logE'], s1['logE'])
                np.testing.assert_array_equal(o['var_logE'], s1['var_logE'])
                assert o['log_jacobian'] == s1['log_jacobian']
    ctx.set_option('batch_stage', 0)               # uploads straight from the caller's arrays: the same numbers
    for o, s1 in zip(ctx.evidence_batch(probs), single):
        np.testing.assert_array_equal(o['logE'], s1['logE'])
    ctx.set_option('batch_stage', 1)
    cross = [{'query': c[:len(c) // 2, 2:], 'ref': c[len(c) // 2:, 2:], 'logl': -c[:len(c) // 2, 1], 'weights': c[:len(c) // 2, 0],
              'kmax': 2, 'ndim': d} for c, (_, d) in zip(chains[:5], shapes)]
    for p, o in zip(cross, ctx.evidence_batch(cross)):
        want = ctx.evidence(p['query'], p['logl'], p['weights'], 2, ref=p['ref'], ndim=p['ndim'])
        np.testing.assert_array_equal(o['logE'], want['logE'])
    bad = np.column_stack((np.ones(50), np.zeros(50), np.ones((50, 2))))
    probs.insert(4, {'query': bad[:, 2:], 'logl': -bad[:, 1], 'weights': bad[:, 0], 'kmax': 3, 'ndim': 2})
    with pytest.raises(_lib.MCEError, match='problem 4'):
        ctx.evidence_batch(probs)
    ctx.close()


def test_row_maps_gather_on_the_device():
    """N2: burn-in / thinning / split as row maps of the chain (query_rows / ref_rows), gathered by the covariance and
    whitening kernels as they load -- bit-identical to the host-gathered arrays (MCEvidence.py:174-249, 417-445)."""
    from mcevidence_b200 import _lib
    rng = np.random.default_rng(5)
    c, _ = gaussian_chain(9000, 7, seed=77, nextra=4, mean_shift=0.3)
    base = c[1200:, 2:]                                   # burn-in = a view of the chain; rows keep the nuisance columns
    nb = base.shape[0]
    keep = np.flatnonzero(rng.poisson(0.6, size=nb) > 0)            # thinning keeps a subset ...
    perm = rng.permutation(len(keep))
    r1, r2 = keep[perm[:len(keep) // 2]], keep[np.sort(perm[len(keep) // 2:])]   # ... the split draws s1 unordered
    w = rng.integers(1, 5, size=nb).astype(float)
    logl = -c[1200:, 1]
    ctx = _lib.Context(0)
    kw = dict(ndim=7, log_prior_volume=0.2)
    # auto mode
    a = ctx.evidence(base, logl[keep], w[keep], 3, query_rows=keep, log_sum_w=math.log(w[keep].sum()), want_knn=True, **kw)
    b = ctx.evidence(base[keep], logl[keep], w[keep], 3, log_sum_w=math.log(w[keep].sum()), want_knn=True, **kw)
    np.testing.assert_array_equal(a['logE'], b['logE'])
    np.testing.assert_array_equal(a['dknn'], b['dknn'])
    np.testing.assert_array_equal(a['idx'], b['idx'])
    assert a['log_jacobian'] == b['log_jacobian']
    # cross mode, both halves rows of the same base (uploaded once), and of different bases
    a = ctx.evidence(base, logl[r1], w[r1], 3, ref=base, query_rows=r1, ref_rows=r2, log_sum_w=1.0, **kw)
    b = ctx.evidence(base[r1], logl[r1], w[r1], 3, ref=base[r2], log_sum_w=1.0, **kw)
    np.testing.assert_array_equal(a['logE'], b['logE'])
    other = np.ascontiguousarray(base[:, :7])
    a2 = ctx.evidence(base, logl[r1], w[r1], 3, ref=other, query_rows=r1, ref_rows=r2, log_sum_w=1.0, **kw)
    np.testing.assert_array_equal(a2['logE'], b['logE'])
    # sweep over prefixes of the mapped rows
    sw = ctx.evidence_sweep(base, logl[keep], w[keep], 3, [700, len(keep)], query_rows=keep, log_sum_w=0.5, **kw)
    np.testing.assert_array_equal(sw[1]['logE'], ctx.evidence(base[keep], logl[keep], w[keep], 3, log_sum_w=0.5, **kw)['logE'])
    with pytest.raises(_lib.MCEError, match='query_rows'):
        ctx.evidence(base, logl[:3], w[:3], 2, query_rows=[0, nb, 1], ndim=7)
    ctx.close()
