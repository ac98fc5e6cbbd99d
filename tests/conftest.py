import glob
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    """Tests marked gpu are skipped (not errored) on a machine without a CUDA device."""
    if not any('gpu' in item.keywords for item in items) or have_gpu():
        return
    skip = pytest.mark.skip(reason='needs a CUDA sm_100 device (B200); the product path has no CPU fallback')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


class Golden:
    """One frozen run of the unmodified reference (written by oracle/gen_golden.py)."""

    def __init__(self, name):
        g = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=True)
        self.name = name
        self.raw = g
        self.chains = [g['chain%d' % i] for i in range(int(g['nchains']))]
        self.ctor = dict(zip([str(k) for k in g['ctor_keys']], [eval(str(v)) for v in g['ctor_vals']]))
        self.ev = dict(zip([str(k) for k in g['ev_keys']], [eval(str(v)) for v in g['ev_vals']]))
        self.mle = g['mle']
        self.np_seed = int(g['np_seed'])
        self.as_files = bool(g['as_files'])
        self.fmt = str(g['fmt'])
        self.ndim = int(g['ndim'])
        self.kmax = int(g['kmax'])
        self.fit_method = str(g['fit_method'])

    def write_files(self, tmpdir):
        root = os.path.join(str(tmpdir), 'chain')
        for i, c in enumerate(self.chains):
            np.savetxt('%s_%d.txt' % (root, i + 1), c, fmt=self.fmt)
        return root

    def method(self, tmpdir=None):
        """The first constructor argument the reference was given."""
        if self.as_files:
            return self.write_files(tmpdir)
        return [c.copy() for c in self.chains]


@pytest.fixture(params=golden_names())
def golden(request):
    return Golden(request.param)


def have_gpu():
    try:
        from mcevidence_b200 import _lib
        return _lib.load_library().mce_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope='session')
def ctx():
    from mcevidence_b200 import _lib
    return _lib.default_context(0)


class OracleContext:
    """Stand-in for _lib.Context.evidence built on the CPU oracle.  TEST-ONLY: lets the CPU suite
    exercise the Python glue of MCEvidence.evidence() (argument plumbing, covtype branches, batching)."""

    supports_row_maps = True

    def evidence(self, query, logl, weights, kmax, ref=None, ndim=None, W=None, W_ref=None, mu=None,
                 log_jacobian=0.0, log_sum_w=0.0, log_prior_volume=0.0, q_range=None, query_rows=None, ref_rows=None,
                 **_):
        from oracle import mce_oracle as O
        kmax = max(2, int(kmax))
        q = np.asarray(query, dtype=np.float64)[:, :ndim]
        r = None if ref is None else np.asarray(ref, dtype=np.float64)[:, :ndim]
        if query_rows is not None:          # row maps (N2): the gather the device does as it loads
            q = q[np.asarray(query_rows)]
        if ref_rows is not None:
            r = r[np.asarray(ref_rows)]
        if W is None:
            cs = O.covariance_stats(q if r is None else np.concatenate((q, r)))
            yq = O.whiten(q, cs['eVec'], cs['eVal'])
            yr = None if r is None else O.whiten(r, cs['eVec'], cs['eVal'])
            log_jacobian = math.log(cs['J'])
        else:
            m = 0.0 if mu is None else np.asarray(mu)
            yq = (q - m) @ np.asarray(W)
            yr = None if r is None else (r - m) @ np.asarray(W if W_ref is None else W_ref)
        if r is None:
            d, _ = O.knn_brute_fp64(yq, yq, kmax, exclude_self=True)
            dknn = np.concatenate((np.zeros((len(yq), 1)), d), axis=1)
            k0 = 1
        else:
            dknn, _ = O.knn_brute_fp64(yr, yq, kmax + 1)
            k0 = 0
        mle, var = O.evidence_from_dknn_log(dknn, np.asarray(logl), np.asarray(weights), ndim, kmax, k0,
                                            math.exp(log_sum_w), log_jacobian, log_prior_volume)
        return {'logE': mle[1:], 'var_logE': var[1:], 'log_jacobian': log_jacobian, 'stats': {}}

    def evidence_sweep(self, query, logl, weights, kmax, batch_rows, ref=None, **kw):
        """Prefix by prefix with the same explicit whitening: what mce_evidence_sweep must reproduce."""
        return [self.evidence(np.asarray(query)[:int(n)], np.asarray(logl)[:int(n)], np.asarray(weights)[:int(n)], kmax,
                              ref=ref, **kw) for n in batch_rows]
