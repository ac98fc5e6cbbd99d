"""BASELINE config 1 shape: `MCEvidence.py <root> -k 2 --burn 0.3` on a CosmoMC-format root (4 chain files of
~6.7k rows, 6 cosmological + nuisance/derived columns, integer weights, .ranges), end to end in one process:
text parsing + burn-in + evidence.  Our CLI (GPU) against the CPU port of the reference path on the same files
(np.loadtxt + numpy/sklearn, n_jobs=-1).  The Planck chains themselves are not bundled.

    python tools/cli_bench.py [--rows 6700] [--extra 30]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import MCEvidence as M                                  # noqa: E402
from mcevidence_b200 import priors                      # noqa: E402
from mcevidence_b200.synthetic import gaussian_chain    # noqa: E402
from oracle import mce_oracle as O                      # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--rows', type=int, default=6700)
ap.add_argument('--extra', type=int, default=30, help='nuisance / derived columns after the 6 cosmological ones')
ap.add_argument('--reps', type=int, default=5)
a = ap.parse_args()

tmp = tempfile.mkdtemp(prefix='mce_cli_')
root = os.path.join(tmp, 'base_plikHM_TT_lowTEB')
names = ['omegabh2', 'omegach2', 'theta', 'tau', 'logA', 'ns'] + ['nuis%d' % i for i in range(a.extra)]
for ic in range(4):
    c, _ = gaussian_chain(a.rows + 13 * ic, 6, seed=100 + ic, cov_seed=7, inflate=1.0, integer_weights=True,
                          nextra=a.extra, mean_shift=0.3)
    np.savetxt('%s_%d.txt' % (root, ic + 1), c, fmt='%.8e')
with open(root + '.ranges', 'w') as fh:
    for nm in names:
        fh.write('%-12s -12 12\n' % nm)
argv = [root, '-k', '2', '--burn', '0.3', '-vb', '0']
M.main(argv)                                            # warm-up: CUDA context, library load
t_gpu = []
for _ in range(a.reps):
    t0 = time.perf_counter()
    mle = M.main(argv)
    t_gpu.append(time.perf_counter() - t0)


def cpu_port():
    par = priors.params_info(root, cosmo=True)
    chains = [np.loadtxt(f) for f in sorted(__import__('glob').glob(root + '_?.txt'))]
    s = np.concatenate([c[int(len(c) * 0.3):] for c in chains])
    return O.evidence(s[:, 2:], -s[:, 1], s[:, 0], par['ndim'], 2, priorvolume=par['volume'], knn='sklearn',
                      n_jobs=-1, scalar_loop=True)


from oracle import ref_loader                           # noqa: E402


def reference_cli():
    """The UNMODIFIED reference on the same files: its own file reader, burn-in, params_info and evidence()."""
    import logging
    R = ref_loader.load()
    logging.disable(logging.CRITICAL)
    try:
        par = R.params_info(root, cosmo=True)
        return R.MCEvidence(root, ndim=par['ndim'], priorvolume=par['volume'], kmax=2, burnlen=0.3, verbose=0).evidence(nproc=-1)
    finally:
        logging.disable(logging.NOTSET)


kind = 'reference' if ref_loader.available() else 'port'
run_cpu = reference_cli if kind == 'reference' else cpu_port
run_cpu()
t_cpu = []
for _ in range(max(2, a.reps // 2)):
    t0 = time.perf_counter()
    want = run_cpu()
    t_cpu.append(time.perf_counter() - t0)
print(json.dumps({'workload': 'cfg1 shape: 4 chains x %d rows x %d columns, ndim 6, kmax 2, burn 0.3' % (a.rows, 8 + a.extra),
                  'ours_cli_s_best': min(t_gpu), 'ours_cli_s_all': t_gpu, 'cpu_s_best': min(t_cpu), 'cpu_kind': kind,
                  'cpu_cores': os.cpu_count(), 'logE': float(mle[0]), 'abs_diff_vs_cpu': float(abs(mle[0] - np.ravel(want)[0])),
                  'file_mb': sum(os.path.getsize('%s_%d.txt' % (root, i + 1)) for i in range(4)) / 1e6}))
