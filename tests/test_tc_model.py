"""The error allowance of the tcgen05 filter engine against an emulation of its coarse scores (TF32-truncated
operands, split h): the allowance must cover the emulated error with margin, and must stay small enough that the
exact slow path is visited only a little more often than an exact-threshold filter inserts (tools/tc_model.py)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
import tc_model  # noqa: E402


@pytest.mark.parametrize('d,L', [(6, 5), (10, 5), (27, 9)])
def test_allowance_covers_tf32_truncation(d, L):
    r = tc_model.model(20000, d, 40, L, seed=3)
    assert r['max_ratio'] < 0.8                       # worst emulated |coarse - exact| / e
    assert r['inserts'] == r['ffma_inserts']          # no reference an exact-threshold filter keeps is dropped
    assert r['visits'] < 1.6 * r['inserts'] + 2       # the allowance adds few slow-path visits at these ndim


@pytest.mark.parametrize('d,L,boot', [(6, 4, 16), (10, 5, 32), (20, 11, 32)])
def test_multi_tile_bootstrap_threshold_is_valid(d, L, boot):
    """The start threshold of the multi-tile bootstrap (group maxima of the first `boot` tiles per column half, the
    (L+1)-th largest minus the allowance) never exceeds what L references other than the query itself reach exactly, the
    scan that starts from it ends with the exact top-L list, and it sends fewer columns to the slow path than a scan
    that starts from the first tile alone."""
    r = tc_model.bootstrap_model(30000, d, 40, L, boot, seed=5)
    assert r['valid'] and r['lists_exact']
    assert r['visits_boot'] < r['visits_first_tile']
