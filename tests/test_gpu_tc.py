"""Parity of the tcgen05 engine of the neighbour filter (``filter_engine`` = 1) with the FP64 oracle and with the
FFMA engine.  The library picks the engine per problem (``filter_engine`` = 2, the default); these tests force the
tensor-core engine on every shape, including the small ones the automatic rule leaves to the FFMA filter.
Bar: bit-exact neighbours, as for the FFMA engine, and a coarse-score error that stays inside the allowance used by
the filter (``tc_check_ratio`` < 1).
"""
import os

import numpy as np
import pytest

from oracle import mce_oracle as O
from mcevidence_b200.synthetic import gaussian_chain

pytestmark = [pytest.mark.gpu]


@pytest.fixture
def tc(ctx):
    ctx.set_option('filter_engine', 1)
    try:
        yield ctx
    finally:
        for name, v in (('filter_engine', 2), ('tc_check', 0), ('tc_slack_percent', 100), ('ref_chunks', -1)):
            ctx.set_option(name, v)


def oracle_knn(x, k, ref=None):
    if ref is not None:
        return O.knn_brute_fp64(ref, x, k)
    d, i = O.knn_brute_fp64(x, x, k - 1, exclude_self=True)
    return (np.concatenate((np.zeros((len(x), 1)), d), axis=1),
            np.concatenate((np.arange(len(x))[:, None], i), axis=1))


@pytest.mark.parametrize('n,d,k,cross', [
    (5, 2, 3, False), (129, 1, 2, False), (128, 3, 4, True), (2000, 5, 5, False), (3000, 10, 5, False),
    (3000, 10, 9, True), (2500, 17, 4, False), (2000, 27, 5, True), (1000, 40, 3, False), (1500, 64, 2, False),
    (4097, 6, 17, False), (20000, 8, 3, False)])
def test_tc_knn_bit_exact(tc, n, d, k, cross):
    """Same contract and cases as test_gpu_parity.test_knn_bit_exact, with every pair also checked against the
    error allowance of the coarse score."""
    rng = np.random.default_rng(1000 * d + n)
    x = rng.normal(size=(n, d)) * (1 + rng.random(d)) + rng.normal(size=d)
    ref = rng.normal(size=(n + 37, d)) * (1 + rng.random(d)) if cross else None
    tc.set_option('tc_check', 1)
    dk, ix, st = tc.knn(x, k, ref=ref, identity=True)
    assert st['filter_mode'] == 16
    od, oi = oracle_knn(x, k, ref)
    np.testing.assert_array_equal(ix, oi)
    np.testing.assert_array_equal(dk, od)
    assert tc.get_debug('tc_check_pairs') > 0
    assert tc.get_debug('tc_check_ratio') < 1.0


@pytest.mark.parametrize('chunks', [1, 3, 7])
def test_tc_reference_split(tc, chunks):
    """Units restricted to a reference range: the re-rank merges the candidate planes of all ranges."""
    rng = np.random.default_rng(7)
    x = rng.normal(size=(3000, 9))
    tc.set_option('ref_chunks', chunks)
    dk, ix, st = tc.knn(x, 6, identity=True)
    od, oi = oracle_knn(x, 6)
    np.testing.assert_array_equal(ix, oi)
    np.testing.assert_array_equal(dk, od)


@pytest.mark.parametrize('d', [5, 10, 27])
def test_tc_matches_ffma_engine(tc, d):
    """Large input: both engines return the same neighbours, and logE agrees to the last bits."""
    chain, _ = gaussian_chain(300000, d, seed=11)
    x, logl, w = np.ascontiguousarray(chain[:, 2:]), -chain[:, 1], chain[:, 0]
    r1 = tc.evidence(x, logl, w, 4, want_knn=True)
    assert r1['stats']['filter_mode'] == 16
    tc.set_option('filter_engine', 0)
    r0 = tc.evidence(x, logl, w, 4, want_knn=True)
    assert r0['stats']['filter_mode'] != 16
    np.testing.assert_array_equal(r1['idx'], r0['idx'])
    np.testing.assert_array_equal(r1['dknn'], r0['dknn'])
    np.testing.assert_allclose(r1['logE'], r0['logE'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('n,d,k', [(400000, 10, 2), (300000, 20, 9), (420000, 6, 5)])
def test_tc_long_units_bootstrap_and_slabs(tc, n, d, k):
    """Units long enough for the multi-tile threshold bootstrap (its 9- and 13-deep networks) and for the slab hit path
    that follows the start-up tiles (whole-range units: no reference split): neighbours equal the FFMA engine's bit for
    bit."""
    chain, _ = gaussian_chain(n, d, seed=23)
    x = np.ascontiguousarray(chain[:, 2:])
    tc.set_option('ref_chunks', 1)
    d1, i1, st = tc.knn(x, k)
    assert st['filter_mode'] == 16
    tc.set_option('filter_engine', 0)
    tc.set_option('ref_chunks', -1)
    d0, i0, st0 = tc.knn(x, k)
    assert st0['filter_mode'] != 16
    np.testing.assert_array_equal(i1, i0)
    np.testing.assert_array_equal(d1, d0)


def test_tc_duplicates_and_padding(tc):
    """Duplicate rows (distance 0 ties ordered by index) and a reference count that is not a tile multiple."""
    rng = np.random.default_rng(5)
    x = rng.normal(size=(1000, 4))
    x[500:520] = x[100:120]
    dk, ix, _ = tc.knn(x, 5, identity=True)
    od, oi = oracle_knn(x, 5)
    np.testing.assert_array_equal(ix, oi)
    np.testing.assert_array_equal(dk, od)
