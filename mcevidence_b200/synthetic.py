"""Seeded synthetic chains with an analytic evidence.

No MCMC chains are bundled with the reference, so every benchmark
and parity case is driven from these generators.  The chain layout is the
CosmoMC column order the reference reads (`MCEvidence.py:126-128`):
column 0 = weight, column 1 = -log(likelihood), columns 2.. = parameters.
"""
from __future__ import annotations

import math
import os

import numpy as np


def gaussian_chain(nsamples, ndim, seed=0, inflate=1.1, offset=0.0, mean_shift=0.0,
                   nextra=0, integer_weights=False, chunk=1 << 20, dtype=np.float64,
                   cov_seed=None):
    """Correlated Gaussian posterior sampled from an inflated proposal.

    Returns ``(chain, info)``; ``chain`` has shape ``(nsamples, 2 + ndim + nextra)``.
    ``info['lnZ']`` is the analytic log-evidence for prior volume 1:
    ``(d/2) ln 2pi + (1/2) ln det C + offset``.

    * ``inflate`` (s): samples are drawn from N(0, s^2 C); the importance weight
      ``exp(-m2/2 (1 - 1/s^2))`` (scaled to max 1) restores the target.
    * ``integer_weights``: CosmoMC-style positive integer multiplicities instead
      (1..5, independent of position; the posterior is then the proposal and
      ``lnZ`` refers to ``logL = -m2/(2 s^2)``).
    * ``nextra``: extra (nuisance) columns appended after the ``ndim`` parameters,
      as in chains read with ``ndim`` smaller than the number of columns.
    * ``cov_seed``: seed of the target covariance when it must be shared by
      several independently drawn chains (default: same stream as ``seed``).
    """
    crng = np.random.default_rng(seed if cov_seed is None else cov_seed)
    a = crng.normal(size=(ndim, ndim))
    rng = crng if cov_seed is None else np.random.default_rng(seed)
    cov = a @ a.T / ndim + np.eye(ndim)
    chol = np.linalg.cholesky(cov)
    s = float(inflate)
    chain = np.empty((nsamples, 2 + ndim + nextra), dtype=dtype)
    for lo in range(0, nsamples, chunk):
        hi = min(nsamples, lo + chunk)
        z = s * rng.normal(size=(hi - lo, ndim))
        m2 = np.einsum('ij,ij->i', z, z)
        chain[lo:hi, 2:2 + ndim] = z @ chol.T + mean_shift
        if integer_weights:
            chain[lo:hi, 0] = rng.integers(1, 6, size=hi - lo)
            chain[lo:hi, 1] = m2 / (2.0 * s * s) - offset
        else:
            chain[lo:hi, 0] = np.exp(-0.5 * m2 * (1.0 - 1.0 / (s * s)))
            chain[lo:hi, 1] = 0.5 * m2 - offset
        if nextra:
            chain[lo:hi, 2 + ndim:] = rng.normal(size=(hi - lo, nextra))
    if not integer_weights:
        chain[:, 0] /= chain[:, 0].max()
    sign, logdet = np.linalg.slogdet(cov)
    lnz = 0.5 * ndim * math.log(2.0 * math.pi) + 0.5 * logdet + offset
    if integer_weights:
        lnz += ndim * math.log(s)
    return chain, {'lnZ': lnz, 'cov': cov, 'ndim': ndim, 'seed': seed}


#: parameter names used for CosmoMC-shaped synthetic roots: six cosmological
#: names the reference recognises (`MCEvidence.py:84-86`) followed by nuisance ones.
COSMO_NAMES = ['omegabh2', 'omegach2', 'theta', 'tau', 'logA', 'ns', 'omegak', 'mnu', 'nnu']
NUISANCE_NAMES = ['calPlanck', 'acib217', 'xi', 'asz143', 'aps100', 'aps143',
                  'aps217', 'aksz', 'kgal100', 'kgal143']


def write_cosmomc_root(root, nchains=4, nrows=6700, ncosmo=6, nnuis=4, seed=0,
                       fixed=('nnu',)):
    """Write ``root_1.txt .. root_<nchains>.txt`` plus ``root.ranges`` in CosmoMC format.

    Shapes follow BASELINE config 1 (four chains of ~6.7k rows, six cosmological
    parameters plus nuisance columns, integer weights).  Returns the list of
    parameter names and the prior volume over the cosmological block.
    """
    names = COSMO_NAMES[:ncosmo] + NUISANCE_NAMES[:nnuis]
    ndim = len(names)
    fixed = tuple(f for f in fixed if f not in names)
    os.makedirs(os.path.dirname(os.path.abspath(root)), exist_ok=True)
    for ic in range(nchains):
        chain, _ = gaussian_chain(nrows + 17 * ic, ndim, seed=seed * 1000 + ic + 1,
                                  cov_seed=seed, inflate=1.0, integer_weights=True,
                                  mean_shift=0.3)
        np.savetxt('%s_%d.txt' % (root, ic + 1), chain, fmt='%.10e')
    lo, hi = -12.0, 12.0
    with open(root + '.ranges', 'w') as fh:
        for nm in names:
            fh.write('%-12s %.6e %.6e\n' % (nm, lo, hi))
        for nm in fixed:
            fh.write('%-12s %.6e %.6e\n' % (nm, 3.046, 3.046))
    return names, (hi - lo) ** ncosmo


def write_cosmomc_grid(rootdir, datasets=('plikHM_TT_lowTEB', 'plikHM_TT_lowTEB_post_BAO'),
                       models=(('base', 6), ('base_omegak', 7), ('base_mnu', 7)), nchains=2, nrows=900, seed=0):
    """A miniature ``COM_CosmoParams_fullGrid_R2.00`` tree: ``<rootdir>/<model>/<data_dir>/<model>_<data>_<i>.txt``
    plus ``.ranges`` and ``.paramnames`` (with a derived ``H0*`` column appended), as read by
    ``planck_mcevidence.py``.  ``models`` lists ``(name, number of cosmological parameters)``.
    Returns ``{(data, model): (root, names, prior volume)}``."""
    out = {}
    for idd, dd in enumerate(datasets):
        for imm, (mm, ncosmo) in enumerate(models):
            root = os.path.join(rootdir, mm, dd.split('_post_')[0], '%s_%s' % (mm, dd))
            names, vol = write_cosmomc_root(root, nchains=nchains, nrows=nrows + 37 * imm, ncosmo=ncosmo, nnuis=3,
                                            seed=seed + 10 * idd + imm)
            for ic in range(nchains):                       # append a derived H0* column
                f = '%s_%d.txt' % (root, ic + 1)
                c = np.loadtxt(f)
                h0 = 67.0 + 3.0 * c[:, 2] + 0.5 * c[:, 3]
                np.savetxt(f, np.column_stack((c, h0)), fmt='%.10e')
            with open(root + '.paramnames', 'w') as fh:
                for nm in names:
                    fh.write('%s\t%s\n' % (nm, nm))
                fh.write('H0*\tH_0\n')
            out[(dd, mm)] = (root, names, vol)
    return out
