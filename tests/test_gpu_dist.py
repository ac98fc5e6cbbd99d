"""Multi-GPU path on real devices (NCCL): skipped unless the box has at least two GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        from mcevidence_b200 import _lib
        return _lib.load_library().mce_device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason='needs two GPUs')
def test_sharded_equals_single_gpu_nccl():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29533', os.path.join(ROOT, 'tools', 'dist_check.py')]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert 'DIST_CHECK OK' in p.stdout


def _group_cases():
    import math
    import numpy as np
    from mcevidence_b200.synthetic import gaussian_chain
    c, _ = gaussian_chain(30011, 9, seed=41, nextra=2, mean_shift=0.2)
    s, logl, w = c[:, 2:], -c[:, 1], c[:, 0]
    return s, logl, w, math.log(w.sum())


def test_device_group_of_one_equals_context():
    """mce_group_* with a single device: the same numbers as the plain context calls (C ABI, one process)."""
    import numpy as np
    from mcevidence_b200 import _lib
    s, logl, w, lsw = _group_cases()
    g = _lib.Group(1)
    assert g.size == 1 and not g.uses_nccl
    ctx = _lib.Context(0)
    a = g.evidence(s, logl, w, 4, ndim=9, log_sum_w=lsw)
    b = ctx.evidence(s, logl, w, 4, ndim=9, log_sum_w=lsw)
    np.testing.assert_array_equal(a['logE'], b['logE'])
    x = g.evidence(s[:15000], logl[:15000], w[:15000], 3, ref=s[15000:], ndim=9, log_sum_w=1.0)
    y = ctx.evidence(s[:15000], logl[:15000], w[:15000], 3, ref=s[15000:], ndim=9, log_sum_w=1.0)
    np.testing.assert_array_equal(x['logE'], y['logE'])
    probs = [{'query': s[i::3], 'logl': logl[i::3], 'weights': w[i::3], 'kmax': 2, 'ndim': 7} for i in range(3)]
    for o, o1 in zip(g.evidence_batch(probs), ctx.evidence_batch(probs)):
        np.testing.assert_array_equal(o['logE'], o1['logE'])
    with pytest.raises(_lib.MCEError):
        _lib.Group(_ngpu() + 1)
    g.close()
    ctx.close()


@pytest.mark.skipif(_ngpu() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('no_nccl', ['0', '1'])
def test_device_group_shards_rows_over_gpus(no_nccl, monkeypatch):
    """mce_group_evidence over every GPU of the box (ncclBroadcast of the samples, or one upload per device with
    MCE_GROUP_NO_NCCL=1) against the single-GPU call: partial sums combine to the same logE within rounding."""
    import numpy as np
    from mcevidence_b200 import _lib
    monkeypatch.setenv('MCE_GROUP_NO_NCCL', no_nccl)
    s, logl, w, lsw = _group_cases()
    g = _lib.Group()
    assert g.size == _ngpu() and g.uses_nccl == (no_nccl == '0')
    ctx = _lib.Context(0)
    for kw in (dict(), dict(ref=s[20000:])):
        n1 = 30011 if not kw else 20000
        a = g.evidence(s[:n1], logl[:n1], w[:n1], 4, ndim=9, log_sum_w=lsw, **kw)
        b = ctx.evidence(s[:n1], logl[:n1], w[:n1], 4, ndim=9, log_sum_w=lsw, **kw)
        assert np.abs(a['logE'] - b['logE']).max() < 1e-12
        assert a['stats']['pairs'] == b['stats']['pairs']
    probs = [{'query': s[i::5], 'logl': logl[i::5], 'weights': w[i::5], 'kmax': 2, 'ndim': 6 + i % 3} for i in range(5)]
    for o, o1 in zip(g.evidence_batch(probs), ctx.evidence_batch(probs)):
        np.testing.assert_array_equal(o['logE'], o1['logE'])
    g.close()
    ctx.close()
