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
