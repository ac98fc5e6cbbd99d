"""Chain container: loading, burn-in, thinning and the s1/s2 split (stays on the host).

Fresh implementation of the behaviour of ``SamplesMIXIN`` / ``MCSamples``
(reference ``MCEvidence.py:99-606``), which defines the arrays handed to the CUDA
hot path.  Column convention of a chain array
(``MCEvidence.py:126-128``): 0 = weight, 1 = -log(likelihood), 2.. = parameters.

Behaviours kept for result parity with the reference:

* a list/tuple of arrays is concatenated **without** burn-in or thinning (the
  reference drops the keyword arguments on that path, ``MCEvidence.py:151``);
  burn-in and thinning apply to chains read from files (``:606``);
* the split draws from the global ``np.random`` stream exactly as
  ``np.random.choice(range(n), int(n*s1frac), replace=False)`` does (``:225``),
  Poisson thinning as one ``np.random.poisson`` draw per row in row order (``:427``);
* ``arrays()`` returns ``-loglikes`` (``:399``).

Deliberate differences: chain files found by a glob are read in sorted order
(the reference order is whatever ``glob`` returns, ``:596``); integer thinning
(``thinlen >= 1``) works (it crashes under numpy >= 1.24 in the reference,
SURVEY.md F10); a dict of arrays is accepted; a burn-in given in rows may be a
float (the CLI passes one, ``:1379-1383``).
"""
from __future__ import annotations

import glob as _glob
import logging
import os

import numpy as np

FORMAT = "%(levelname)s:%(filename)s.%(funcName)s():%(lineno)-8s %(message)s"


class data_set(object):
    """One partition of the samples (``MCEvidence.py:99-104``).

    The partition is either an array (``samples``) or -- next-row N2, burn-in / thinning / split applied on the
    device -- a *row map* into the concatenated chain: ``base`` (the parameter columns of the chain, a view) and
    ``rows`` (the rows burn-in, thinning and the random split kept, in the reference's order).  ``samples``
    materialises ``base[rows]`` on first use, so every caller of the reference's attribute keeps working; the hot
    path asks :meth:`device_view` instead and lets the covariance / whitening kernels gather the rows as they load,
    so no gathered host copy of the chain is ever made.
    """

    def __init__(self, d):
        self._samples = d.get('samples')
        self._base = d.get('base')
        self._rows = d.get('rows')
        self.weights = d['weights']
        self.loglikes = d['loglikes']
        self.adjusted_weights = d['aweights']

    @property
    def samples(self):
        if self._samples is None and self._base is not None:
            self._samples = self._base if self._rows is None else self._base[self._rows]
        return self._samples

    @samples.setter
    def samples(self, value):
        self._samples, self._base, self._rows = value, None, None

    @property
    def nrows(self):
        if self._base is not None:
            return self._base.shape[0] if self._rows is None else len(self._rows)
        return 0 if self._samples is None else self._samples.shape[0]

    @property
    def ncols(self):
        a = self._base if self._base is not None else self._samples
        return 0 if a is None else a.shape[1]

    def device_view(self):
        """``(base, rows)`` when the partition is still a row map of the chain (``rows`` None: all rows), else None."""
        if self._base is None or self._rows is None:
            return None
        return self._base, self._rows


def read_chain_file(fname, nthreads=0):
    """Parse a whitespace-separated chain file into a C-contiguous float64 ``(rows, columns)`` array.

    Same values as the reference's ``np.loadtxt`` (``MCEvidence.py:564``: correctly rounded
    decimal->double, ``#`` comments, blank lines skipped), parsed by the multi-threaded reader of the
    shared library (``mce_read_chain_text``); anything it does not accept (ragged rows, exotic number
    formats) is handed to ``np.loadtxt`` so that error behaviour stays numpy's.
    """
    try:
        from . import _lib
        return _lib.read_chain_text(fname, nthreads)
    except Exception:
        return np.ascontiguousarray(np.atleast_2d(np.loadtxt(fname)), dtype=np.float64)


class LoggingHandler(object):
    """Mixin of the reference (``MCEvidence.py:91-97``): ``set_logger`` attaches a logger named after the calling
    source line."""

    def set_logger(self):
        self.logger = logging.getLogger(self.log_message())

    def log_message(self):
        import inspect
        return str(inspect.stack()[2][4])


_READ_POOL = None


def _read_pool():
    """Process-wide pool of file-reading threads (created on first use; a grid run reads hundreds of roots)."""
    global _READ_POOL
    if _READ_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _READ_POOL = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1), thread_name_prefix='mce-read')
    return _READ_POOL


class SamplesMIXIN(object):
    """Loading-independent half of the chain container."""

    def __init__(self):
        raise NotImplementedError()

    def load_from_file(self, fname, **kwargs):
        raise NotImplementedError()

    # -- construction -------------------------------------------------------
    def setup(self, str_or_dict, **kwargs):
        self.iw = kwargs.pop('iw', 0)
        self.ilike = kwargs.pop('ilike', 1)
        self.itheta = kwargs.pop('itheta', 2)
        level = kwargs.pop('log_level', logging.INFO)
        logging.basicConfig(level=level, format=FORMAT)
        self.logger = logging.getLogger(__name__)
        if self.debug:
            self.logger.setLevel(logging.DEBUG)

        if isinstance(str_or_dict, str):
            self.logger.info('Loading chain from ' + str_or_dict)
            self.data = self.load_from_file(str_or_dict, **kwargs)
        elif isinstance(str_or_dict, (list, tuple)) and len(str_or_dict) and isinstance(str_or_dict[0], str):
            # explicit list of file names (MCEvidence.py:576-579, 708-709)
            self.data = self.load_from_file(list(str_or_dict), **kwargs)
        elif isinstance(str_or_dict, (dict, list, tuple)):
            if isinstance(str_or_dict, dict):
                self.chains = list(str_or_dict.values())
            else:
                self.chains = list(str_or_dict)
            # arrays: burn-in / thinning keywords are not forwarded (MCEvidence.py:151)
            self.data = self.chains2samples()
        else:
            self.logger.error('first argument should be a file name string, list, tuple or dict.')
            raise TypeError('unsupported chain input of type %s' % type(str_or_dict))

        ndim = self.get_shape()[1]
        if getattr(self, 'names', None) is None:
            self.names = ['p%s' % i for i in range(ndim)]
        if getattr(self, 'labels', None) is None:
            self.labels = ['%s_%s' % (self.px, i) for i in range(ndim)]
        if not hasattr(self, 'trueval'):
            self.trueval = None
        self.nparamMC = self.get_shape()[1]

    def chains2samples(self, **kwargs):
        """Burn each chain, concatenate, thin, split (``MCEvidence.py:174-219``)."""
        if self.chains is None or len(self.chains) == 0:
            self.logger.error('The chains array is empty!')
            raise ValueError('no chains to combine')
        burnlen = kwargs.pop('burnlen', 0)
        thinlen = kwargs.pop('thinlen', 0)
        self.chains = [np.atleast_2d(np.asarray(c, dtype=np.float64)) for c in self.chains]
        self.nchains = len(self.chains)
        self.ichain = np.concatenate([(i + 1) * np.ones(len(c)) for i, c in enumerate(self.chains)])
        if burnlen > 0:
            self.chains = [self.removeBurn(burnlen, chain=c) for c in self.chains]
        self.chain_offsets = np.cumsum(np.array([0] + [c.shape[0] for c in self.chains]))
        self.samples = np.concatenate(self.chains)          # the one host pass over the chain (burn-in = views)
        rows, weights = None, None
        if np.abs(thinlen) > 0 and thinlen != 1:             # thin() is a no-op for nthin == 1 (MCEvidence.py:296-297)
            # thinning keeps a row map and the new weights; the rows themselves are gathered on the device
            weights, rows = self.get_thin_index(thinlen, self.samples[:, self.iw])
            self.logger.info('Thinning with thin length={} #old_chain={},#new_chain={}'.format(
                thinlen, len(self.samples), len(weights)))
            rows = np.asarray(rows, dtype=np.int64)
            weights = np.asarray(weights, dtype=np.float64)
        self.chains = None
        return self.chain_split(self.samples, rows=rows, weights=weights)

    def chain_split(self, s, rows=None, weights=None):
        """Random s1/s2 split from the global numpy RNG (``MCEvidence.py:221-249``).

        ``rows`` / ``weights``: the rows of ``s`` kept by thinning and their new weights (None: every row with its
        own weight column).  The partitions are row maps into ``s`` (see :class:`data_set`)."""
        nrow = len(s) if rows is None else len(rows)
        w_all = s[:, self.iw] if weights is None else weights
        base = s[:, self.itheta:]

        def part(sel):
            r = sel if rows is None else rows[sel]
            return {'weights': w_all[sel], 'loglikes': s[r, self.ilike], 'base': base, 'rows': r, 'ichain': sel}

        if self.split:
            ix = np.random.choice(nrow, size=int(nrow * self.s1frac), replace=False)
            mask = np.ones(nrow, dtype=bool)
            mask[ix] = False
            not_ix = np.flatnonzero(mask)            # == np.setxor1d(range(nrow), ix)
            self.logger.info('{} chain with nrow={} split to ns1={}, ns2={}'.format(
                self.nchains, nrow, len(ix), len(not_ix)))
            s1_dict, s2_dict = part(ix), part(not_ix)
        else:
            if rows is None:
                s1_dict = {'weights': s[:, self.iw], 'loglikes': s[:, self.ilike], 'samples': base,
                           'ichain': range(len(s))}
            else:
                s1_dict = part(np.arange(nrow))
            s2_dict = {'weights': None, 'loglikes': None, 'samples': None, 'ichain': None}
        s1_dict['aweights'] = np.copy(s1_dict['weights'])
        s2_dict['aweights'] = np.copy(s2_dict['weights']) if s2_dict['weights'] is not None else None
        return {'s1': data_set(s1_dict), 's2': data_set(s2_dict)}

    # -- accessors ----------------------------------------------------------
    def get_shape(self, name='s1'):
        def shape_of(ds):
            return (ds.nrows, ds.ncols)
        if name in ('s1', 's2'):
            return shape_of(self.data[name])
        s1 = shape_of(self.data['s1'])
        s2 = shape_of(self.data['s2'])
        return (s1[0] + s2[0], s1[1])

    def arrays(self, name='s1'):
        """``(samples, lnp = -loglikes, weights)`` of a partition (``MCEvidence.py:394-405``)."""
        if name in ('s1', 's2'):
            s = self.data[name].samples
            if s is None:
                return None, None, None
            return s, -self.data[name].loglikes, self.data[name].weights
        return self.all_sample_arrays()

    def all_sample_arrays(self):
        s, lnp, w = self.arrays('s1')
        s2, lnp2, w2 = self.arrays('s2')
        if s2 is None:
            return s, lnp, w
        return np.concatenate((s, s2)), np.concatenate((lnp, lnp2)), np.concatenate((w, w2))

    def importance_sample(self, func, name='s1'):
        """Rescale ``adjusted_weights`` by ``exp(-func(samples))`` (``MCEvidence.py:265-270``)."""
        self.logger.info('Importance sampling partition: {}'.format(name))
        neg_log_likes = func(self.data[name].samples)
        self.data[name].adjusted_weights *= np.exp(-neg_log_likes)

    # -- burn-in -------------------------------------------------------------
    def removeBurn(self, remove, chain=None, name=None):
        """Drop the first rows: a fraction if ``remove < 1`` else a row count (``MCEvidence.py:350-391``)."""
        nstart = remove
        if chain is None and name is None:
            return nstart
        if remove < 1:
            if chain is not None:
                nstart = int(chain.shape[0] * remove)
            if name is not None:
                nstart = int(len(self.data[name].loglikes) * remove)
        else:
            nstart = int(remove)
        self.logger.info('Removing %s lines as burn in' % nstart)
        if chain is not None:
            return chain[nstart:, :]
        d = self.data[name]
        d.samples = d.samples[nstart:, :]
        d.loglikes = d.loglikes[nstart:]
        d.weights = d.weights[nstart:]
        return None

    # -- thinning ------------------------------------------------------------
    def get_thin_index(self, nthin, weights):
        if nthin < 0:
            raise NotImplementedError('thinning by the autocorrelation length (thinlen < 0) is not implemented '
                                      '(nor is it in the reference)')
        if nthin < 1:
            thin_ix, new_weights = self.poisson_thin(nthin, weights=weights)
        else:
            try:
                thin_ix, new_weights = self.thin_indices(nthin, weights=weights)
            except ValueError:
                thin_ix, new_weights = self.weighted_thin(nthin, weights=weights)
        return new_weights, thin_ix

    def thin(self, nthin=1, name=None, chain=None):
        """Thin a chain array, the concatenated samples or a partition (``MCEvidence.py:289-348``)."""
        if nthin == 1:
            return None
        if chain is not None:
            weights = chain[:, self.iw]
            new_weights, thin_ix = self.get_thin_index(nthin, weights)
            output = chain[thin_ix, :]
            output[:, self.iw] = new_weights
        elif name is None:
            weights = self.samples[:, self.iw]
            new_weights, thin_ix = self.get_thin_index(nthin, weights)
            self.samples = self.samples[thin_ix, :]
            self.samples[:, self.iw] = new_weights
            output = self.samples
        else:
            d = self.data[name]
            weights = d.weights
            new_weights, thin_ix = self.get_thin_index(nthin, weights)
            d.weights = new_weights
            d.samples = d.samples[thin_ix, :]
            d.loglikes = d.loglikes[thin_ix]
            d.adjusted_weights = d.weights.copy()
            output = d
        self.logger.info('Thinning with thin length={} #old_chain={},#new_chain={}'.format(
            nthin, len(weights), len(new_weights)))
        return output

    def poisson_thin(self, thin_retain_frac, name='s1', weights=None):
        """New weight ~ Poisson(weight * frac), rows with 0 dropped (``MCEvidence.py:417-445``)."""
        if weights is None:
            weights = self.data[name].weights.copy()
        w = weights * thin_retain_frac
        new_w = np.random.poisson(w).astype(np.float64)   # same stream as one scalar draw per row
        thin_ix = np.where(new_w > 0)[0]
        return thin_ix, new_w[thin_ix]

    def weighted_thin(self, thin_unit, name='s1', weights=None):
        """Keep the heaviest sample of each of N/thin_unit equal bins (``MCEvidence.py:447-479``)."""
        if weights is None:
            weights = self.data[name].weights.copy()
        n = len(weights)
        if thin_unit == 0:
            return np.arange(n), weights
        n2 = int(n * thin_unit) if thin_unit < 1 else int(n // thin_unit)
        n2 = max(n2, 1)
        bins = np.linspace(-1, n, n2 + 1)
        ind = np.digitize(np.arange(n), bins)
        order = np.lexsort((np.arange(n), -np.asarray(weights), ind))   # first max per bin
        first = np.ones(n, dtype=bool)
        first[1:] = ind[order][1:] != ind[order][:-1]
        thin_ix = np.sort(order[first]).astype(np.intp)
        return thin_ix, weights[thin_ix]

    def thin_indices(self, factor, name='s1', weights=None):
        """getdist's integer-weight thinning (``MCEvidence.py:481-532``): unit-weight samples every
        ``factor`` units of cumulative weight.  Raises ``ValueError`` for non-integer weights/factor."""
        if weights is None:
            weights = self.data[name].weights.copy()
        numrows = len(weights)
        norm1 = np.sum(weights)
        iweights = weights.astype(np.int64)
        norm = np.sum(iweights)
        if abs(norm - norm1) > 1e-4:
            raise ValueError('Can only thin with integer weights')
        if factor != int(factor):
            raise ValueError('Thin factor must be integer')
        factor = int(factor)
        if factor >= np.max(iweights):
            cumsum = np.cumsum(iweights) // factor
            _, thin_ix = np.unique(cumsum, return_index=True)
        else:
            # sample k (k = 1, 2, ...) sits at cumulative weight k*factor: the row whose
            # cumulative-weight interval (c_{i-1}, c_i] contains it
            csum = np.cumsum(iweights)
            targets = np.arange(1, norm // factor + 1, dtype=np.int64) * factor
            thin_ix = np.searchsorted(csum, targets, side='left')
            thin_ix = thin_ix[thin_ix < numrows]
        return thin_ix, iweights[thin_ix]


class MCSamples(SamplesMIXIN):
    """File / array chain reader (``MCEvidence.py:536-606``)."""

    def __init__(self, str_or_dict, trueval=None, debug=False, csplit=None, names=None, labels=None, px='x',
                 **kwargs):
        self.debug = debug
        self.names = None
        self.labels = None
        self.trueval = trueval
        self.px = px
        if csplit is None:
            self.split, self.s1frac, self.shuffle = False, 0.5, True
        else:
            self.split, self.s1frac, self.shuffle = csplit.split, csplit.frac, csplit.shuffle
        self.setup(str_or_dict, **kwargs)

    def read_list_to_array(self, flist):
        """The chain files of one root, parsed concurrently (the parser releases the GIL): a root is typically
        4-8 files of about a megabyte, too small for the threads of a single parse to pay off."""
        for f in flist:
            self.logger.info('loading: ' + f)
        if len(flist) == 1:
            return [read_chain_file(flist[0])]
        ncpu = os.cpu_count() or 1
        per_file = max(1, ncpu // len(flist))
        return list(_read_pool().map(lambda f: read_chain_file(f, per_file), flist))

    def load_from_file(self, fname, **kwargs):
        """``fname``: a file, a list of files, a wildcard, or a CosmoMC root (``root_?.txt``;
        ``idchain > 0`` selects ``root_<idchain>.txt``) -- ``MCEvidence.py:567-606``."""
        idchain = kwargs.pop('idchain', 0)
        idpattern = kwargs.pop('idpattern', '_?.txt')
        if isinstance(fname, (list, tuple)):
            flist = list(fname)
        elif os.path.isfile(fname):
            flist = [fname]
        elif '*' in fname or '?' in fname:
            flist = sorted(_glob.glob(fname))
        elif idchain > 0:
            flist = [fname + '_{}.txt'.format(idchain)]
        else:
            self.logger.info(' loading files: ' + fname + idpattern)
            flist = sorted(_glob.glob(fname + idpattern))
        if not flist:
            raise IOError('Can not read chain: no files match %r' % (fname,))
        try:
            self.chains = self.read_list_to_array(flist)
        except Exception:
            print('Can not read chain from the following list of files: ', flist)
            raise
        return self.chains2samples(**kwargs)
