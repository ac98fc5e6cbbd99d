"""``MCEvidence``: the reference's estimator class with the CUDA hot path behind it.

Same constructor and ``evidence()`` signature, attributes and return values as the
reference class (``MCEvidence.py:613-1168``).  What differs is where the arithmetic
runs: the covariance / whitening / neighbour search / volume reduction of
``MCEvidence.py:1033-1131`` is one call into ``libmcevidence_b200.so``
(``Context.evidence`` -> ``mce_evidence``) instead of numpy + scikit-learn.  There is
no CPU fallback.
"""
from __future__ import annotations

import logging
import math
from collections import namedtuple

import numpy as np

from . import _lib
from .chains import FORMAT, MCSamples

__all__ = ['MCEvidence']


class MCEvidence(object):
    def __init__(self, method, ischain=True, isfunc=None,
                 thinlen=0.0, burnlen=0.0,
                 split=False, s1frac=0.5, shuffle=True,
                 ndim=None, kmax=5,
                 priorvolume=1, debug=False,
                 nsample=None, covtype='single',
                 nbatch=1,
                 brange=None,
                 bscale='',
                 verbose=1, args={},
                 device=0,
                 **gdkwargs):
        """Evidence estimation from MCMC chains (Heavens et al. 2017, arXiv:1704.03472).

        Parameters are those of the reference (``MCEvidence.py:614-654``): ``method`` is a chain
        root / file name(s) / list, tuple or dict of ``(n, 2+p)`` arrays with columns
        ``[weight, -logL, theta...]``; ``kmax`` gives results for k = 1..kmax-1; ``split=True``
        computes the cross evidence between two random halves; ``thinlen`` / ``burnlen`` as in
        the chain reader; ``nbatch`` / ``brange`` / ``bscale`` select sub-sample sizes.
        ``device`` (new) is the CUDA device ordinal.
        """
        logging.basicConfig(level=logging.DEBUG, format=FORMAT)
        self.logger = logging.getLogger(__name__)
        self.verbose = verbose
        self.debug = bool(debug or verbose > 1)
        log_level = logging.DEBUG if self.debug else (logging.INFO if verbose == 1 else logging.WARNING)
        self.logger.setLevel(log_level)

        self.info = {}
        self.split = split
        self.covtype = covtype
        self.nbatch = nbatch
        self.brange = brange
        self.bscale = bscale if not isinstance(self.brange, int) else 'constant'
        self.snames = ['s1'] + (['s2'] if self.split else [])
        self.idbatch = np.arange(self.nbatch, dtype=int)
        self.powers = np.zeros((self.nbatch, len(self.snames)))
        self.bsize = np.zeros((self.nbatch, len(self.snames)), dtype=int)
        self.nchain = np.zeros((self.nbatch, len(self.snames)), dtype=int)
        self.kmax = max(2, kmax)                                   # MCEvidence.py:694
        self.priorvolume = priorvolume
        self.ischain = ischain
        self.device = device
        self.fname = method if isinstance(method, str) else None

        if not ischain:
            # a sampler object/class producing chains (MCEvidence.py:715-743; broken upstream)
            self.nsample = 100000 if nsample is None else nsample
            sampler = method(*args) if isinstance(method, type) else method
            self.method = sampler
            method = sampler.Sampler(nsamples=self.nsample)

        gdkwargs.setdefault('thinlen', thinlen)
        gdkwargs.setdefault('burnlen', burnlen)
        gdkwargs.setdefault('log_level', log_level)
        split_var = namedtuple('split_var', 'split frac shuffle')
        csplit = split_var(split=self.split, frac=s1frac, shuffle=shuffle)
        self.gd = MCSamples(method, csplit=csplit, debug=self.debug, **gdkwargs)

        if isfunc:
            self.gd.importance_sample(isfunc, name='s1')
            if self.split:
                self.gd.importance_sample(isfunc, name='s2')

        self.info['NparamsMC'] = self.gd.nparamMC
        self.info['Nsamples_read'] = self.gd.get_shape()[0]
        self.info['Nparams_read'] = self.gd.get_shape()[1]
        self.nsample = [self.gd.get_shape(name=s)[0] for s in self.snames]
        if ndim is None:
            ndim = self.gd.nparamMC
        self.ndim = ndim
        # Limits of the CUDA library that scikit-learn does not have (include/mcevidence_b200.h): checked here, before
        # anything is sent to the GPU, so that the error names the argument instead of coming back from the device.
        if self.ndim > _lib.MCE_MAX_DIM:
            raise ValueError('ndim=%d: the B200 filter kernels serve at most %d parameters (MCE_MAX_DIM); select fewer '
                             'columns with ndim= / --ndim' % (self.ndim, _lib.MCE_MAX_DIM))
        kcap = _lib.MCE_MAX_NEIGHBOURS + (0 if self.split else 1)
        if self.kmax > kcap:
            raise ValueError('kmax=%d: at most %d true neighbours per sample are searched (MCE_MAX_NEIGHBOURS), i.e. '
                             'kmax <= %d in %s mode' % (self.kmax, _lib.MCE_MAX_NEIGHBOURS, kcap,
                                                        'cross' if self.split else 'auto'))
        self.info['NparamsCosmo'] = self.ndim
        self.info['Nsamples'] = ', '.join([str(x) for x in self.nsample])
        self.logger.info('chain array dimensions: %s x %s =' % (self.nsample, self.ndim))
        self.set_batch()
        self.last_stats = None

    # -- bookkeeping kept from the reference ------------------------------------------------
    def summary(self):
        print()
        print('ndim={}'.format(self.ndim))
        print('nsample={}'.format(self.nsample))
        print('kmax={}'.format(self.kmax))
        print('brange={}'.format(self.brange))
        print('bsize={}'.format(self.bsize))
        print('powers={}'.format(self.powers))
        print('nchain={}'.format(self.nchain))
        print()

    def get_batch_range(self):
        if self.brange is None:
            return None, None
        powmin, powmax = np.array(self.brange).min(), np.array(self.brange).max()
        if powmin == powmax and self.nbatch > 1:
            self.logger.error('nbatch>1 but batch range is set to zero.')
            raise ValueError('nbatch>1 but batch range is set to zero.')
        return powmin, powmax

    def set_batch(self, bscale=None):
        """Sub-sample sizes per batch (``MCEvidence.py:808-840``); default one batch of everything."""
        if bscale is None:
            bscale = self.bscale
        else:
            self.bscale = bscale
        if self.brange is None:
            self.bsize = self.brange
            for ix, nn in enumerate(self.nsample):
                self.nchain[0, ix] = nn
                self.powers[0, ix] = np.log10(nn)
        elif bscale == 'logpower':
            powmin, powmax = self.get_batch_range()
            for ix, nn in enumerate(self.nsample):
                self.powers[:, ix] = np.linspace(powmin, powmax, self.nbatch)
                self.bsize[:, ix] = np.array([int(pow(10.0, x)) for x in self.powers[:, ix]])
            self.nchain = self.bsize
        elif bscale == 'linear':
            powmin, powmax = self.get_batch_range()
            for ix, nn in enumerate(self.nsample):
                self.bsize[:, ix] = np.linspace(powmin, powmax, self.nbatch, dtype=int)
                self.powers[:, ix] = np.array([int(math.log10(x)) for x in self.bsize[:, ix]])
            self.nchain = self.bsize
        else:  # constant
            self.bsize[:, :] = self.brange
            self.powers[:, :] = self.idbatch[:, None]
            for ix, nn in enumerate(self.nsample):
                self.nchain[:, ix] = self.bsize[:, ix].cumsum()

    # -- numpy restatements kept as public helpers (used for covtype='single' and batching) ----
    def diagonalise_chain(self, s, eigenVec, eigenVal):
        """``s @ eVec / sqrt(eVal)`` (``MCEvidence.py:842-849``)."""
        return np.dot(s, eigenVec) / np.sqrt(eigenVal)[None, :self.ndim]

    def get_covariance(self, s=None):
        """Unweighted covariance of all samples, eigen-system and Jacobian (``MCEvidence.py:851-882``)."""
        if s is None:
            s = self.gd.all_sample_arrays()[0][:, 0:self.ndim]
        cov = np.atleast_2d(np.cov(s.T))
        e_val, e_vec = np.linalg.eig(cov)
        if (e_val < 0).any():
            self.logger.warning('Some of the eigenvalues of the covariance matrix are negative and/or complex')
            return {'cov': cov, 'posdef': False, 'J': 1, 'eVec': e_vec, 'eVal': e_val}
        return {'cov': cov, 'posdef': True, 'J': math.sqrt(np.linalg.det(cov)), 'eVec': e_vec, 'eVal': e_val}

    def get_samples(self, nsamples, istart=0, rand=False, name='s1', prewhiten=True):
        """Rows ``istart..istart+nsamples`` (or a random draw) of a partition, first ``ndim`` columns
        (``MCEvidence.py:884-947``).  ``prewhiten`` whitens with the partition's own covariance."""
        ntot = self.gd.get_shape(name)[0]
        s, lnp, w = self.gd.arrays(name)
        s = s[:, 0:self.ndim]
        if nsamples > 0:
            if rand and self.brange is not None:
                if nsamples > ntot:
                    raise ValueError('partition %s nsamples=%s, ntotal_chain=%s' % (name, nsamples, ntot))
                idx = np.random.randint(0, high=ntot, size=nsamples)
                s, lnp, w = s[idx, :], lnp[idx], w[idx]
            elif istart != 0 or nsamples != ntot:
                idx = np.arange(istart, nsamples + istart)
                s, lnp, w = s[idx, :], lnp[idx], w[idx]
        jac, e_vec, e_val = 1, None, None
        if prewhiten:
            cs = self.get_covariance(s=s)
            e_vec, e_val = cs['eVec'], cs['eVal']
            if cs['posdef']:
                jac = cs['J']
                s = self.diagonalise_chain(s, e_vec, e_val)
        return s, lnp, w, {'J': jac, 'eVec': e_vec, 'eVal': e_val}

    def _device_views(self):
        """``[(base, rows) of s1, (base, rows) of s2]`` when the partitions are still row maps of the chain (files read
        with thinning and / or a split) and the context can follow them, else None."""
        ctx = _lib.default_context(self.device)
        if not getattr(ctx, 'supports_row_maps', False):
            return None
        v1 = self.gd.data['s1'].device_view()
        if v1 is None:
            return None
        if not self.split:
            return [v1]
        v2 = self.gd.data['s2'].device_view()
        return None if v2 is None else [v1, v2]

    @staticmethod
    def _reference_factor(cs):
        """The reference's whitening matrix ``eVec / sqrt(eVal)`` for an explicit-W call."""
        if not cs['posdef']:
            raise np.linalg.LinAlgError('covariance matrix is not positive definite '
                                        '(negative eigenvalue, MCEvidence.py:867)')
        return np.ascontiguousarray(np.real(cs['eVec']) / np.sqrt(np.real(cs['eVal']))[None, :])

    # -- the hot path ---------------------------------------------------------------------------
    def evidence(self, verbose=None, rand=False, info=False, covtype='all',
                 profile=False, pvolume=None, pos_lnp=False,
                 nproc=-1, prewhiten=True):
        """k-th nearest-neighbour estimate of the log-evidence (``MCEvidence.py:950-1168``).

        Returns ``MLE[0, 1:]`` -- ln E for k = 1..kmax-1 -- or ``MLE[:, 1:]`` when batching
        (``brange`` set), and ``(MLE, self.info)`` with ``info=True``.  ``pvolume`` overrides the
        constructor's prior volume, ``pos_lnp=True`` flips the sign of the likelihood column,
        ``covtype`` is ``'all'`` (covariance of s1 and s2 together) or ``'single'``; ``nproc``
        (scikit-learn's ``n_jobs``), ``profile`` and ``prewhiten`` are accepted and ignored.
        The delta-method variance of each estimate is left in ``self.var_logE``.
        """
        if verbose is None:
            verbose = self.verbose
        log_pv = math.log(self.priorvolume if pvolume is None else pvolume)   # MCEvidence.py:1017-1020
        kmax, ndim = self.kmax, self.ndim
        if covtype is None:
            covtype = self.covtype
        ctx = _lib.default_context(self.device)

        mle = np.zeros((self.nbatch, kmax))
        var = np.zeros((self.nbatch, kmax))
        full = self.brange is None
        sum_w = np.sum(self.gd.data['s1'].adjusted_weights)                   # MCEvidence.py:1126
        cov_all = None
        batches = list(zip(self.idbatch, self.nchain))
        if not full and covtype == 'all' and not rand and len(batches) > 1 and hasattr(ctx, 'evidence_sweep'):
            # N4: every batch is a prefix of s1 whitened with the covariance of ALL samples (MCEvidence.py:1034-1045),
            # so one upload, one whitening and one pass of the filter over the largest prefix serve all of them
            s1_all, logl, weight = self.gd.arrays('s1')
            samples = s1_all[:, 0:ndim]
            if pos_lnp:
                logl = -logl                                                  # MCEvidence.py:1060
            samples2 = self.get_samples(0, prewhiten=False, name='s2')[0] if self.split else None
            cov_all = self.get_covariance()
            cov_mu = self.gd.all_sample_arrays()[0][:, 0:ndim].mean(axis=0)
            rows = [int(nsample[0]) for _, nsample in batches]
            outs = ctx.evidence_sweep(samples, logl, weight, kmax, rows, ref=samples2, ndim=ndim,
                                      W=self._reference_factor(cov_all), mu=cov_mu, log_jacobian=math.log(cov_all['J']),
                                      log_sum_w=math.log(sum_w), log_prior_volume=log_pv)
            for (ipow, _), out in zip(batches, outs):
                mle[ipow, 1:] = out['logE']
                var[ipow, 1:] = out['var_logE']
            self.last_stats = outs[-1]['stats']
            batches = []
        views = self._device_views() if (full and covtype == 'all' and batches) else None
        if views is not None:
            # N2: burn-in / thinning / split stay row maps of the chain; the device gathers the rows as it loads
            # (MCEvidence.py:174-249, 350-391, 417-445 decide WHICH rows on the host, with numpy's global RNG)
            d1 = self.gd.data['s1']
            logl = d1.loglikes if pos_lnp else -d1.loglikes                   # MCEvidence.py:399, 1060
            kw = {'query_rows': views[0][1]}
            ref = None
            if self.split:
                ref, kw['ref_rows'] = views[1][0][:, 0:ndim], views[1][1]
                self.logger.info('using XMCEvidence. NN distance is estimated using nsamples=(%s, %s)'
                                 % (len(views[0][1]), len(views[1][1])))
            out = ctx.evidence(views[0][0][:, 0:ndim], logl, d1.weights, kmax, ref=ref, ndim=ndim,
                               log_sum_w=math.log(sum_w), log_prior_volume=log_pv, **kw)
            ipow = batches[0][0]
            mle[ipow, 1:] = out['logE']
            var[ipow, 1:] = out['var_logE']
            self.last_stats = out['stats']
            batches = []
        for ipow, nsample in batches:
            s_n = int(nsample[0])
            samples, logl, weight, _ = self.get_samples(s_n, istart=0, rand=rand, prewhiten=False, name='s1')
            if pos_lnp:
                logl = -logl                                                  # MCEvidence.py:1060
            samples2 = None
            if self.split:
                samples2 = self.get_samples(0, prewhiten=False, name='s2')[0]
                self.logger.info('using XMCEvidence. NN distance is estimated using nsamples=(%s, %s)'
                                 % (s_n, samples2.shape[0]))
            kw = {}
            if covtype == 'all' and full:
                pass   # library estimates cov(s1 U s2) on the GPU and whitens with its Cholesky factor
            elif covtype == 'all':
                # batching: the covariance is still that of ALL samples (MCEvidence.py:1034-1037, 860-861)
                if cov_all is None:
                    cov_all = self.get_covariance()
                    cov_mu = self.gd.all_sample_arrays()[0][:, 0:ndim].mean(axis=0)
                kw = {'W': self._reference_factor(cov_all), 'mu': cov_mu, 'log_jacobian': math.log(cov_all['J'])}
            elif covtype == 'single':
                cs1 = self.get_covariance(s=samples)                          # MCEvidence.py:1052-1054
                kw = {'W': self._reference_factor(cs1), 'log_jacobian': math.log(cs1['J'])}
                if self.split:
                    # s2 is whitened with its own eigen-system, J stays that of s1 (:1080-1086);
                    # the two bases differ, so no common centring is applied
                    kw['W_ref'] = self._reference_factor(self.get_covariance(s=samples2))
                else:
                    kw['mu'] = samples.mean(axis=0)
            else:
                raise ValueError("covtype must be 'all' or 'single'")
            out = ctx.evidence(samples, logl, weight, kmax, ref=samples2, ndim=ndim,
                               log_sum_w=math.log(sum_w), log_prior_volume=log_pv, **kw)
            mle[ipow, 1:] = out['logE']
            var[ipow, 1:] = out['var_logE']
            self.last_stats = out['stats']
            self.logger.debug('SumW={} \t S={} \t logJ={} \t logPriorVolume={}'.format(
                sum_w, s_n, out['log_jacobian'], log_pv))

        if full:
            mle, var = mle[0, 1:], var[0, 1:]                                 # MCEvidence.py:1155-1157
        else:
            mle, var = mle[:, 1:], var[:, 1:]
        self.var_logE = var
        if verbose > 0:
            for k in range(1, self.kmax):
                self.logger.info('   ln(B)[k={}] = {}'.format(k, mle[..., k - 1]))
        if info:
            return mle, self.info
        return mle
