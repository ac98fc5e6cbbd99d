"""Freeze outputs of the UNMODIFIED reference into tests/golden/ (run in the build container).

The reference has no tests or golden vectors, so the parity
pin is the reference itself, imported from ``/root/reference`` through the
two-line ``DistanceMetric`` shim of SURVEY.md F10 and run on seeded synthetic
chains.  ``/root/reference`` does not exist on the GPU box, hence the frozen
fixtures.  Usage::

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz
    python oracle/gen_golden.py NAME ...   # only the named fixtures

Each fixture stores the input chain(s), the call arguments and what the
reference returned (``MLE`` per k; for the small cases also ``DkNN`` and
``indices`` captured from the sklearn call at MCEvidence.py:1104).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REFERENCE_DIR = os.environ.get('MCE_REFERENCE_DIR', '/root/reference')


def import_reference():
    """Import the reference module unmodified."""
    import sklearn.metrics
    import sklearn.neighbors
    sklearn.neighbors.DistanceMetric = sklearn.metrics.DistanceMetric
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import importlib
    ref = importlib.import_module('MCEvidence')
    assert os.path.dirname(os.path.abspath(ref.__file__)) == os.path.abspath(REFERENCE_DIR), ref.__file__
    return ref


class KnnSpy:
    """Records what ``NearestNeighbors.kneighbors`` returned inside the reference."""

    def __init__(self):
        import sklearn.neighbors
        self.cls = sklearn.neighbors.NearestNeighbors
        self.orig = self.cls.kneighbors
        self.calls = []

    def __enter__(self):
        spy = self

        def wrapped(self_nn, *a, **kw):
            out = spy.orig(self_nn, *a, **kw)
            spy.calls.append((np.array(out[0]), np.array(out[1]), np.array(a[0]) if a else None,
                              np.array(self_nn._fit_X), self_nn._fit_method))
            return out
        self.cls.kneighbors = wrapped
        return self

    def __exit__(self, *exc):
        self.cls.kneighbors = self.orig


FMT = '%.10e'


ONLY = set()      # names given on the command line: regenerate just these fixtures


def run_case(ref, name, chains, np_seed, keep_knn=False, evidence_kwargs=None, as_files=False, **ctor):
    """Run the reference on ``chains`` (arrays, or files ``root_<i>.txt`` when ``as_files``)."""
    if ONLY and name not in ONLY:
        return
    evidence_kwargs = evidence_kwargs or {}
    if as_files:
        import tempfile
        tmp = tempfile.mkdtemp(prefix='mce_golden_')
        root = os.path.join(tmp, 'chain')
        rounded = []
        for i, c in enumerate(chains):
            np.savetxt('%s_%d.txt' % (root, i + 1), c, fmt=FMT)
            rounded.append(np.loadtxt('%s_%d.txt' % (root, i + 1)))
        chains = rounded
        method = root
    else:
        method = [c.copy() for c in chains]
    np.random.seed(np_seed)
    with KnnSpy() as spy:
        mce = ref.MCEvidence(method, verbose=0, **ctor)
        mle = mce.evidence(**evidence_kwargs)
    dknn, idx, query, fitx, fit_method = spy.calls[-1]
    out = {
        'np_seed': np_seed,
        'mle': np.asarray(mle),
        'nchains': len(chains),
        'ctor_keys': np.array(sorted(ctor.keys())),
        'ctor_vals': np.array([repr(ctor[k]) for k in sorted(ctor.keys())]),
        'ev_keys': np.array(sorted(evidence_kwargs.keys())),
        'ev_vals': np.array([repr(evidence_kwargs[k]) for k in sorted(evidence_kwargs.keys())]),
        'fit_method': fit_method, 'as_files': as_files, 'fmt': FMT,
        'ndim': mce.ndim, 'kmax': mce.kmax,
        'nsample': np.array(mce.nsample),
        # post burn-in / thinning / split partitions exactly as the reference built them
        # (first parameter column is enough to identify rows; weights change under thinning)
        's1_col0': mce.gd.data['s1'].samples[:, 0], 's1_weights': mce.gd.data['s1'].weights,
    }
    if mce.gd.data['s2'].samples is not None:
        out.update({'s2_col0': mce.gd.data['s2'].samples[:, 0], 's2_weights': mce.gd.data['s2'].weights})
    for i, c in enumerate(chains):
        out['chain%d' % i] = c
    if keep_knn:
        out.update({'dknn': dknn, 'idx': idx})
    path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
    np.savez_compressed(path, **out)
    print('%-28s mle=%s  fit=%s  S=%d  -> %s (%.0f kB)' % (
        name, np.array2string(np.asarray(mle), precision=10), fit_method, len(dknn), os.path.relpath(path, ROOT),
        os.path.getsize(path) / 1024))


def main():
    from mcevidence_b200.synthetic import gaussian_chain
    ref = import_reference()
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)

    # auto mode, importance weights, kd-tree branch (d <= 15), ndim < ncolumns
    c, _ = gaussian_chain(3000, 8, seed=11, nextra=3, mean_shift=0.7)
    run_case(ref, 'auto_d8_kd', [c], 101, keep_knn=True, kmax=4, ndim=8, priorvolume=7.0)

    # auto mode, integer weights, two chains passed as arrays: burnlen/thinlen are silently
    # ignored for array input (setup() calls chains2samples() without kwargs, MCEvidence.py:151)
    c1, _ = gaussian_chain(1800, 5, seed=21, cov_seed=20, integer_weights=True, inflate=1.0)
    c2, _ = gaussian_chain(1700, 5, seed=22, cov_seed=20, integer_weights=True, inflate=1.0)
    run_case(ref, 'auto_d5_arrays_burn_ignored', [c1, c2], 102, kmax=3, burnlen=0.3)

    # file root input with burn-in fraction (MCEvidence.py:362-377); a single file keeps the
    # glob order (MCEvidence.py:596, unsorted) out of the comparison
    c1, _ = gaussian_chain(3000, 5, seed=23, integer_weights=True, inflate=1.0)
    run_case(ref, 'auto_d5_file_burn', [c1], 109, as_files=True, kmax=3, burnlen=0.3)

    # auto mode, brute branch of sklearn (d > 15): DGEMM-trick distances
    c, _ = gaussian_chain(2500, 17, seed=31)
    run_case(ref, 'auto_d17_brute', [c], 103, kmax=5, priorvolume=3.0)

    # cross mode (split=True): unseeded global RNG split -> np.random.seed before construction
    c, _ = gaussian_chain(4000, 6, seed=41, nextra=2, mean_shift=-1.5)
    run_case(ref, 'cross_d6', [c], 104, keep_knn=True, kmax=4, ndim=6, split=True, priorvolume=2.5)

    # cross mode in the brute branch, uneven split
    c, _ = gaussian_chain(3000, 18, seed=51)
    run_case(ref, 'cross_d18_brute', [c], 105, kmax=3, split=True, s1frac=0.4)

    # file input: Poisson thinning (0 < thinlen < 1, MCEvidence.py:417-445) after a
    # burn-in given in rows (an int: a float row count breaks the slice at :377)
    c, _ = gaussian_chain(5000, 4, seed=61, integer_weights=True, inflate=1.0)
    run_case(ref, 'auto_d4_file_poisson_thin', [c], 106, as_files=True, kmax=2, burnlen=500, thinlen=0.5)

    # kmax below 2 is clamped to 2 (MCEvidence.py:694); pos_lnp and pvolume overrides
    c, _ = gaussian_chain(1500, 3, seed=71)
    c[:, 1] = -c[:, 1]
    run_case(ref, 'auto_d3_poslnp', [c], 107, kmax=1, evidence_kwargs={'pos_lnp': True, 'pvolume': 11.0})

    # strongly anisotropic scales and a large mean (cosmology-like parameter magnitudes)
    c, _ = gaussian_chain(2500, 6, seed=81)
    scale = np.array([2e-4, 1e-3, 5e-4, 1e-2, 1.0, 30.0])
    shift = np.array([0.022, 0.12, 1.04, 0.08, 3.1, 67.0])
    c[:, 2:8] = c[:, 2:8] * scale + shift
    run_case(ref, 'auto_d6_scaled', [c], 108, kmax=3, priorvolume=1e-3)

    # batch-size sweep (A7): first S rows per batch, covariance of ALL samples, SumW of all of s1
    # (MCEvidence.py:808-840, 1034-1045, 1126); returns MLE[:, 1:] of shape (nbatch, kmax-1)
    c, _ = gaussian_chain(2600, 5, seed=91, nextra=1)
    run_case(ref, 'auto_d5_batches', [c], 110, kmax=3, ndim=5, nbatch=3, brange=[2.6, 3.4], bscale='logpower')

    # covtype='single' in cross mode: s2 whitened with its own eigen-system, J of s1 (MCEvidence.py:1052-1054,
    # 1080-1086) -- the two partitions live in different bases, so the explicit W / W_ref path is exercised
    c, _ = gaussian_chain(3200, 5, seed=92, mean_shift=0.4)
    run_case(ref, 'cross_d5_covsingle', [c], 111, kmax=3, split=True, evidence_kwargs={'covtype': 'single'})

    # Whole-number thinning (thinlen >= 1, an int: a float breaks np.linspace / slicing under current numpy).
    # Integer weights: getdist's algorithm (MCEvidence.py:481-532).  It uses the removed alias np.int, and the bare
    # `except` at :283 would silently route the call to weighted_thin instead; the alias is restored for these two
    # runs so that the unmodified source follows the path it specifies (same kind of shim as DistanceMetric).
    added_alias = not hasattr(np, 'int')
    if added_alias:
        np.int = int
    try:
        c, _ = gaussian_chain(4000, 4, seed=95, integer_weights=True, inflate=1.0)
        run_case(ref, 'auto_d4_file_int_thin', [c], 112, as_files=True, kmax=3, burnlen=200, thinlen=3)
        # non-integer weights: the heaviest sample of each bin (weighted_thin, MCEvidence.py:447-479)
        c, _ = gaussian_chain(4000, 4, seed=96, mean_shift=0.3)
        run_case(ref, 'auto_d4_file_weighted_thin', [c], 113, as_files=True, kmax=3, burnlen=0.2, thinlen=4)
    finally:
        if added_alias:
            del np.int


if __name__ == '__main__':
    ONLY.update(sys.argv[1:])      # python oracle/gen_golden.py [fixture names]: default = all
    main()
