"""Python glue of MCEvidence.evidence() on the CPU: the CUDA context is replaced by a TEST-ONLY stand-in
built on the oracle (conftest.OracleContext), so that argument plumbing, sign conventions, covtype
branches and batching are checked against the frozen reference outputs without a GPU."""
import numpy as np
import pytest

import MCEvidence as M
from mcevidence_b200 import _lib
from conftest import OracleContext


@pytest.fixture
def oracle_ctx(monkeypatch):
    monkeypatch.setattr(_lib, 'default_context', lambda device=0: OracleContext())


def test_evidence_glue_matches_reference(golden, tmp_path, oracle_ctx):
    g = golden
    np.random.seed(g.np_seed)
    mce = M.MCEvidence(g.method(tmp_path), verbose=0, **g.ctor)
    assert mce.kmax == g.kmax and mce.ndim == g.ndim
    mle = mce.evidence(**g.ev)
    assert mle.shape == g.mle.shape and mle.shape[-1] == g.kmax - 1
    assert np.abs(mle - g.mle).max() < 1e-9
    out, info = mce.evidence(info=True, **g.ev)
    assert set(info) == {'NparamsMC', 'Nsamples_read', 'Nparams_read', 'NparamsCosmo', 'Nsamples'}
    assert mce.var_logE.shape == mle.shape


def test_covtype_single_and_batches(oracle_ctx):
    """covtype='single' (own eigen-system per partition) and brange batching, against the oracle pieces."""
    from oracle import mce_oracle as O
    from mcevidence_b200.synthetic import gaussian_chain
    import math
    c, _ = gaussian_chain(1200, 4, seed=3, mean_shift=2.0)
    # auto, single == all (one partition)
    a = M.MCEvidence([c], kmax=3, verbose=0).evidence(covtype='all')
    b = M.MCEvidence([c], kmax=3, verbose=0).evidence(covtype='single')
    c0 = M.MCEvidence([c], kmax=3, verbose=0, covtype='single').evidence(covtype=None)
    assert np.abs(a - b).max() < 1e-9 and np.abs(c0 - b).max() == 0
    # cross, single: s2 whitened with its own eigen-system, J from s1 (MCEvidence.py:1080-1086)
    np.random.seed(5)
    m = M.MCEvidence([c], kmax=3, split=True, verbose=0)
    got = m.evidence(covtype='single')
    s1, s2 = m.gd.data['s1'], m.gd.data['s2']
    cs1, cs2 = O.covariance_stats(s1.samples), O.covariance_stats(s2.samples)
    y1 = O.whiten(s1.samples, cs1['eVec'], cs1['eVal'])
    y2 = O.whiten(s2.samples, cs2['eVec'], cs2['eVal'])
    dk, _ = O.knn_brute_fp64(y2, y1, 4)
    want = O.evidence_from_dknn_linear(dk, -s1.loglikes, s1.weights, 4, 3, 0, s1.weights.sum(), cs1['J'], 0.0)[1:]
    assert np.abs(got - want).max() < 1e-9
    # batches: first S rows of s1, covariance of ALL samples, SumW of all of s1 (MCEvidence.py:1034-1045,1126)
    mb = M.MCEvidence([c], kmax=3, verbose=0, nbatch=2, brange=[2.5, 3.0], bscale='logpower')
    res = mb.evidence()
    assert res.shape == (2, 2)
    cs = O.covariance_stats(c[:, 2:])
    for i, S in enumerate([int(10 ** 2.5), 1000]):
        y = O.whiten(c[:S, 2:], cs['eVec'], cs['eVal'])
        d, _ = O.knn_brute_fp64(y, y, 3, exclude_self=True)
        dk = np.concatenate((np.zeros((S, 1)), d), axis=1)
        want = O.evidence_from_dknn_linear(dk, -c[:S, 1], c[:S, 0], 4, 3, 1, c[:, 0].sum(), cs['J'], 0.0)[1:]
        assert np.abs(res[i] - want).max() < 1e-9
    with pytest.raises(ValueError):
        M.MCEvidence([c], verbose=0).evidence(covtype='bogus')


def test_library_limits_raise_before_the_gpu():
    """ndim > MCE_MAX_DIM and more than MCE_MAX_NEIGHBOURS neighbours are rejected in the constructor."""
    import MCEvidence as M
    rng = np.random.default_rng(0)
    wide = np.column_stack((np.ones(200), np.zeros(200), rng.normal(size=(200, 70))))
    with pytest.raises(ValueError, match='MCE_MAX_DIM'):
        M.MCEvidence([wide], kmax=2, verbose=0)
    M.MCEvidence([wide], kmax=2, ndim=64, verbose=0)
    with pytest.raises(ValueError, match='MCE_MAX_NEIGHBOURS'):
        M.MCEvidence([wide], kmax=18, ndim=5, verbose=0)
    M.MCEvidence([wide], kmax=17, ndim=5, verbose=0)
    with pytest.raises(ValueError, match='MCE_MAX_NEIGHBOURS'):
        M.MCEvidence([wide], kmax=17, ndim=5, split=True, verbose=0)
