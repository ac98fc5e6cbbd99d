"""The C-ABI shared library: loads, exports every declared symbol, struct layouts match the header,
host-side arithmetic is right, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import math
import os
import re
import subprocess

import numpy as np
import pytest

from mcevidence_b200 import _lib
from conftest import ROOT, have_gpu

HEADER = os.path.join(ROOT, 'include', 'mcevidence_b200.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mce_[a-z_0-9]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    names = declared_functions()
    assert {'mce_evidence', 'mce_knn', 'mce_evidence_batch', 'mce_ctx_create', 'mce_finalize'} <= set(names)
    for n in names:
        assert hasattr(lib, n), 'symbol %s declared in the header is not exported' % n
    assert lib.mce_version() == 2


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof as seen by a C compiler vs the ctypes mirrors in mcevidence_b200/_lib.py."""
    probe = tmp_path / 'probe.c'
    fields = {'mce_problem': [f for f, _ in _lib.Problem._fields_],
              'mce_stats': [f for f, _ in _lib.Stats._fields_],
              'mce_result': [f for f, _ in _lib.Result._fields_]}
    body = ['#include <stdio.h>', '#include <stddef.h>', '#include "mcevidence_b200.h"', 'int main(void){']
    for s, fs in fields.items():
        body.append('printf("%s %%zu\\n", sizeof(%s));' % (s, s))
        for f in fs:
            body.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (s, f, s, f))
    body.append('return 0;}')
    probe.write_text('\n'.join(body))
    exe = str(tmp_path / 'probe')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.dirname(HEADER), str(probe), '-o', exe])
    got = dict(line.split() for line in subprocess.check_output([exe]).decode().splitlines())
    mirror = {'mce_problem': _lib.Problem, 'mce_stats': _lib.Stats, 'mce_result': _lib.Result}
    for s, cls in mirror.items():
        assert int(got[s]) == C.sizeof(cls), s
        for f in fields[s]:
            assert int(got['%s.%s' % (s, f)]) == getattr(cls, f).offset, (s, f)


def test_finalize_matches_reference_formula():
    """mce_finalize against the linear-space arithmetic of MCEvidence.py:1107-1131."""
    from oracle import mce_oracle as O
    rng = np.random.default_rng(0)
    S, d, kmax = 500, 7, 4
    for cross in (0, 1):
        dknn = np.sort(rng.random((S, kmax + 1)) + 0.05, axis=1)
        logl = -rng.random(S) * 30
        w = rng.integers(1, 6, S).astype(float)
        k0 = 0 if cross else 1
        sum_w, jac, log_pv = w.sum(), 2.7, math.log(3.0)
        want = O.evidence_from_dknn_linear(dknn, logl, w, d, kmax, k0, sum_w, jac, log_pv)[1:]
        parts = []
        for k in range(1, kmax):
            t = O.logsumexp_terms(dknn[:, k], logl, w, d)
            m = t.max()
            parts.append([m, np.exp(t - m).sum(), np.exp(2 * (t - m)).sum()])
        loge, var = _lib.finalize(np.array(parts), d, cross, S, math.log(sum_w), math.log(jac), log_pv)
        assert np.abs(loge - want).max() < 1e-11
        _, ovar = O.evidence_from_dknn_log(dknn, logl, w, d, kmax, k0, sum_w, math.log(jac), log_pv)
        np.testing.assert_allclose(var, ovar[1:], rtol=1e-12)


def test_combine_partials():
    rng = np.random.default_rng(1)
    t = rng.normal(size=(3, 1000)) * 20
    chunks = np.array_split(np.arange(1000), 7)
    parts = np.array([[[t[k, c].max(), np.exp(t[k, c] - t[k, c].max()).sum(),
                        np.exp(2 * (t[k, c] - t[k, c].max())).sum()] for k in range(3)] for c in chunks])
    parts[2, 1] = [-np.inf, 0.0, 0.0]          # an empty shard
    t[1, chunks[2]] = -np.inf
    out = _lib.combine_partials(parts)
    for k in range(3):
        m = t[k].max()
        assert out[k, 0] == m
        np.testing.assert_allclose(out[k, 1], np.exp(t[k] - m).sum(), rtol=1e-13)
        np.testing.assert_allclose(out[k, 2], np.exp(2 * (t[k] - m)).sum(), rtol=1e-13)


@pytest.mark.skipif(have_gpu(), reason='checks the behaviour on a machine without a GPU')
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(_lib.MCEError, match='no CPU fallback'):
        _lib.Context(0)
    import MCEvidence as M
    from mcevidence_b200.synthetic import gaussian_chain
    c, _ = gaussian_chain(200, 3, seed=0)
    with pytest.raises(_lib.MCEError):
        M.MCEvidence([c], verbose=0).evidence()


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(ImportError, match='no CPU fallback'):
        _lib.load_library()


def test_split_planner_invariants():
    """mce_plan_split (host arithmetic of the filter's work decomposition): structural invariants."""
    slots = 148 * 4
    for n in (1, 300, 15000, 125000, 1000000, 16000000):
        for d in (1, 7, 10, 27, 64):
            for n_ref in (n, max(1, n // 3), 4 * n + 17):
                nfull, chunks, cost = _lib.plan_split(n, n_ref, d)
                nblocks = -(-n // 256)
                assert 0 <= nfull <= nblocks and nfull % slots == 0        # whole waves only
                assert 1 <= chunks <= 64 and cost > 0
                if nfull == nblocks:
                    assert chunks == 1
    # large inputs: whole-set units for the full waves, a modest split of the remainder
    nfull, chunks, _ = _lib.plan_split(1000000, 1000000, 10)
    assert nfull == (3907 // slots) * slots and 2 <= chunks <= 16
    # fewer blocks than CTA slots: split until every slot has a unit
    nfull, chunks, _ = _lib.plan_split(15000, 15000, 10)
    assert nfull == 0 and 59 * chunks >= slots * 0.9
    # the insertion cost weighs more at low ndim: never more splits than at high ndim for the same shape
    c_lo = _lib.plan_split(125000, 1000000, 2)[1]
    c_hi = _lib.plan_split(125000, 1000000, 30)[1]
    assert c_lo <= c_hi
