"""ctypes binding of ``libmcevidence_b200.so`` (C ABI in ``include/mcevidence_b200.h``).

This is the whole Python<->CUDA boundary of the package: plain pointers and
sizes, no torch types.  There is no CPU fallback -- if the shared library is
missing or no sm_100 device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

MCE_MAX_DIM = 64
MCE_MAX_NEIGHBOURS = 16
FLAG_DEVICE_PTRS = 1
FLAG_PARTIALS = 2
FLAG_EXACT_ONLY = 4
FLAG_F32_SAMPLES = 8

ERR_NOT_POSDEF = -3

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmcevidence_b200.so')

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class Problem(C.Structure):
    _fields_ = [
        ('query', C.c_void_p), ('ref', C.c_void_p), ('logl', C.c_void_p), ('weights', C.c_void_p),
        ('n_query', C.c_int64), ('n_ref', C.c_int64), ('ld_query', C.c_int64), ('ld_ref', C.c_int64),
        ('q_begin', C.c_int64), ('q_end', C.c_int64),
        ('d', C.c_int32), ('kmax', C.c_int32), ('flags', C.c_int32), ('reserved', C.c_int32),
        ('W', C.c_void_p), ('W_ref', C.c_void_p), ('mu', C.c_void_p),
        ('log_jacobian', C.c_double), ('log_sum_w', C.c_double), ('log_prior_volume', C.c_double),
        ('query_rows', C.c_void_p), ('ref_rows', C.c_void_p), ('n_query_base', C.c_int64), ('n_ref_base', C.c_int64),
    ]


class Stats(C.Structure):
    _fields_ = [
        ('ms_total', C.c_double), ('ms_h2d', C.c_double), ('ms_cov', C.c_double), ('ms_whiten', C.c_double),
        ('ms_filter', C.c_double), ('ms_rerank', C.c_double), ('ms_exact', C.c_double), ('ms_reduce', C.c_double),
        ('ms_d2h', C.c_double),
        ('pairs', C.c_int64), ('rows_rescanned', C.c_int64),
        ('launches', C.c_int32), ('filter_tq', C.c_int32), ('filter_tr', C.c_int32), ('filter_packed', C.c_int32),
        ('ref_chunks', C.c_int32), ('list_len', C.c_int32), ('filter_grid', C.c_int32), ('filter_smem', C.c_int32),
        ('filter_ctas_per_sm', C.c_int32), ('rmax', C.c_float), ('filter_mode', C.c_int32),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class Result(C.Structure):
    _fields_ = [
        ('logE', C.c_void_p), ('var_logE', C.c_void_p), ('dknn', C.c_void_p), ('idx', C.c_void_p),
        ('partials', C.c_void_p), ('log_jacobian', C.c_double), ('stats', Stats),
    ]


class MCEError(RuntimeError):
    def __init__(self, code, message):
        super().__init__('mcevidence_b200 error %d: %s' % (code, message))
        self.code = code


_lib = None
_lib_lock = threading.Lock()


def load_library():
    """Load the CUDA shared library; raise if it has not been built (no fallback)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                'mcevidence_b200: %s not found. Build it with `make -C mcevidence_b200/csrc -j8` '
                '(or `python -c "import __graft_entry__ as g; g.build()"`). There is no CPU fallback.' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.mce_version.restype = C.c_int
        lib.mce_last_error.restype = C.c_char_p
        lib.mce_device_count.restype = C.c_int
        lib.mce_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        lib.mce_ctx_destroy.argtypes = [C.c_void_p]
        lib.mce_ctx_destroy.restype = None
        lib.mce_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.mce_ctx_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.mce_ctx_get_debug.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]
        lib.mce_ctx_get_debug_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        lib.mce_evidence.argtypes = [C.c_void_p, C.POINTER(Problem), C.POINTER(Result)]
        lib.mce_evidence_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(Problem), C.POINTER(Result)]
        lib.mce_evidence_sweep.argtypes = [C.c_void_p, C.POINTER(Problem), C.c_int, c_int64_p, C.POINTER(Result)]
        lib.mce_knn.argtypes = [C.c_void_p, C.POINTER(Problem), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        lib.mce_group_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
        lib.mce_group_destroy.argtypes = [C.c_void_p]
        lib.mce_group_destroy.restype = None
        lib.mce_group_size.argtypes = [C.c_void_p]
        lib.mce_group_uses_nccl.argtypes = [C.c_void_p]
        lib.mce_group_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.mce_group_evidence.argtypes = [C.c_void_p, C.POINTER(Problem), C.POINTER(Result)]
        lib.mce_group_evidence_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(Problem), C.POINTER(Result)]
        lib.mce_combine_partials.argtypes = [C.c_int, C.c_int, c_double_p, c_double_p]
        lib.mce_finalize.argtypes = [c_double_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double,
                                     C.c_double, c_double_p, c_double_p]
        lib.mce_fma_peak.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p]
        lib.mce_unit_peak.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p]
        lib.mce_plan_split.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int), c_double_p]
        lib.mce_read_chain_text.argtypes = [C.c_char_p, C.c_int, C.POINTER(c_double_p), c_int64_p, c_int64_p]
        lib.mce_free_host.argtypes = [C.c_void_p]
        lib.mce_free_host.restype = None
        _lib = lib
        return lib


def _check(lib, rc):
    if rc != 0:
        raise MCEError(rc, lib.mce_last_error().decode('utf-8', 'replace'))


def _ptr(a):
    return None if a is None else a.ctypes.data


def _f64(a, name, allow_f32=False):
    """float64 (or, for sample arrays, float32) array whose rows are contiguous (strided rows go through ``ld``)."""
    a = np.asarray(a)
    if a.dtype != np.float64 and not (allow_f32 and a.dtype == np.float32):
        a = a.astype(np.float64)
    if a.ndim == 2:
        isz = a.dtype.itemsize
        if a.strides[1] != isz or a.strides[0] % isz or a.strides[0] < isz * a.shape[1]:
            a = np.ascontiguousarray(a)
    elif not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    return a


class Context:
    """One CUDA stream + workspace on one device (``mce_ctx``)."""

    supports_row_maps = True     # query_rows / ref_rows: burn-in, thinning and split applied by the device (N2)

    def __init__(self, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        _check(self._lib, self._lib.mce_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, '_h', None):
            self._lib.mce_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        """Run on an existing CUDA stream given by its raw handle (``torch.cuda.current_stream().cuda_stream``).
        ``0`` is the legacy default stream (torch's default), passed to the library as ``cudaStreamLegacy``;
        ``None`` restores the context's own non-blocking stream.  Work is ordered after everything already
        queued on that stream (e.g. an NCCL broadcast of the inputs)."""
        if cuda_stream is None:
            handle = 0
        else:
            handle = int(cuda_stream) or 1          # cudaStreamLegacy == (cudaStream_t)0x1
        _check(self._lib, self._lib.mce_ctx_set_stream(self._h, C.c_void_p(handle)))

    def set_option(self, name, value):
        _check(self._lib, self._lib.mce_ctx_set_option(self._h, name.encode(), int(value)))

    def get_debug(self, name):
        """``mce_ctx_get_debug``: diagnostics of the last call (e.g. ``tc_check_ratio``)."""
        v = C.c_double(0.0)
        _check(self._lib, self._lib.mce_ctx_get_debug(self._h, name.encode(), C.byref(v)))
        return v.value

    # ------------------------------------------------------------------
    def get_debug_array(self, name):
        """``mce_ctx_get_debug_array``: e.g. ``tc_dump`` -> float32 (256, 128) coarse scores of unit 0, tile 0."""
        n = C.c_int64(0)
        _check(self._lib, self._lib.mce_ctx_get_debug_array(self._h, name.encode(), None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        if n.value:
            _check(self._lib, self._lib.mce_ctx_get_debug_array(self._h, name.encode(), _ptr(out), n.value, C.byref(n)))
        return out.reshape(-1, 128) if n.value else out

    def _problem(self, query, ref, logl, weights, ndim, kmax, W, W_ref, mu, log_jacobian, log_sum_w,
                 log_prior_volume, q_range, flags, keep, query_rows=None, ref_rows=None):
        """``query_rows`` / ``ref_rows``: ``query`` / ``ref`` are BASE arrays and the sample sets are their rows
        ``query_rows`` / ``ref_rows`` (burn-in, thinning and split applied by the device as it loads)."""
        p = Problem()
        # float32 sample arrays are passed as they are (MCE_FLAG_F32_SAMPLES) when query and ref agree
        f32 = np.asarray(query).dtype == np.float32 and (ref is None or np.asarray(ref).dtype == np.float32)
        query = _f64(query, 'query', allow_f32=f32)
        isz = query.dtype.itemsize
        if f32:
            flags |= FLAG_F32_SAMPLES
        keep.append(query)
        if query.ndim != 2:
            raise ValueError('query must be 2-D')
        ndim = int(query.shape[1] if ndim is None else ndim)
        if ndim > query.shape[1]:
            raise ValueError('ndim=%d exceeds the %d sample columns' % (ndim, query.shape[1]))
        p.query = _ptr(query)
        p.n_query, p.ld_query = query.shape[0], query.strides[0] // isz
        if query_rows is not None:
            query_rows = np.ascontiguousarray(query_rows, dtype=np.int64).reshape(-1)
            keep.append(query_rows)
            p.query_rows, p.n_query_base, p.n_query = _ptr(query_rows), query.shape[0], query_rows.shape[0]
        if ref is not None:
            ref = _f64(ref, 'ref', allow_f32=f32)
            keep.append(ref)
            if ref.ndim != 2 or ref.shape[1] < ndim:
                raise ValueError('ref must be 2-D with at least ndim columns')
            p.ref = _ptr(ref)
            p.n_ref, p.ld_ref = ref.shape[0], ref.strides[0] // isz
            if ref_rows is not None:
                ref_rows = np.ascontiguousarray(ref_rows, dtype=np.int64).reshape(-1)
                keep.append(ref_rows)
                p.ref_rows, p.n_ref_base, p.n_ref = _ptr(ref_rows), ref.shape[0], ref_rows.shape[0]
        for name, arr in (('logl', logl), ('weights', weights)):
            if arr is not None:
                arr = _f64(arr, name).reshape(-1)
                if arr.shape[0] != p.n_query:
                    raise ValueError('%s has %d entries for %d samples' % (name, arr.shape[0], p.n_query))
                keep.append(arr)
                setattr(p, name, _ptr(arr))
        for name, arr, shape in (('W', W, (ndim, ndim)), ('W_ref', W_ref, (ndim, ndim)), ('mu', mu, (ndim,))):
            if arr is not None:
                arr = np.ascontiguousarray(arr, dtype=np.float64)
                if arr.shape != shape:
                    raise ValueError('%s must have shape %s' % (name, shape))
                keep.append(arr)
                setattr(p, name, _ptr(arr))
        if q_range is not None:
            p.q_begin, p.q_end = int(q_range[0]), int(q_range[1])
        p.d, p.kmax, p.flags = ndim, int(kmax), int(flags)
        p.log_jacobian, p.log_sum_w, p.log_prior_volume = float(log_jacobian), float(log_sum_w), float(log_prior_volume)
        nq = (p.q_end - p.q_begin) if p.q_end else p.n_query
        return p, nq

    def evidence(self, query, logl, weights, kmax, ref=None, ndim=None, W=None, W_ref=None, mu=None,
                 log_jacobian=0.0, log_sum_w=0.0, log_prior_volume=0.0, q_range=None, want_knn=False,
                 want_partials=False, exact_only=False, query_rows=None, ref_rows=None):
        """``mce_evidence`` on host arrays.  Returns a dict with ``logE`` (kmax-1), ``var_logE``,
        ``log_jacobian``, ``stats`` and optionally ``dknn``/``idx``/``partials``."""
        keep = []
        flags = (FLAG_PARTIALS if want_partials else 0) | (FLAG_EXACT_ONLY if exact_only else 0)
        p, nq = self._problem(query, ref, logl, weights, ndim, kmax, W, W_ref, mu, log_jacobian, log_sum_w,
                              log_prior_volume, q_range, flags, keep, query_rows, ref_rows)
        kk = max(2, int(kmax))
        r = Result()
        loge = np.empty(kk - 1)
        var = np.empty(kk - 1)
        r.logE, r.var_logE = _ptr(loge), _ptr(var)
        out = {}
        if want_knn:
            out['dknn'] = np.empty((nq, kk + 1))
            out['idx'] = np.empty((nq, kk + 1), dtype=np.int64)
            r.dknn, r.idx = _ptr(out['dknn']), _ptr(out['idx'])
        if want_partials:
            out['partials'] = np.empty((kk - 1, 3))
            r.partials = _ptr(out['partials'])
        _check(self._lib, self._lib.mce_evidence(self._h, C.byref(p), C.byref(r)))
        out.update({'logE': loge, 'var_logE': var, 'log_jacobian': r.log_jacobian, 'stats': r.stats.as_dict()})
        return out

    def evidence_batch(self, problems):
        """``mce_evidence_batch``: a list of dicts with the keyword arguments of :meth:`evidence`
        (``query, logl, weights, kmax`` and optionally ``ref, ndim, log_sum_w, log_prior_volume``), run back
        to back on this context (chain-per-GPU batching).  Returns a list of result dicts."""
        n = len(problems)
        keep, outs = [], []
        parr = (Problem * n)()
        rarr = (Result * n)()
        for i, kw in enumerate(problems):
            kw = dict(kw)
            p, _ = self._problem(kw['query'], kw.get('ref'), kw['logl'], kw['weights'], kw.get('ndim'), kw['kmax'],
                                 kw.get('W'), kw.get('W_ref'), kw.get('mu'), kw.get('log_jacobian', 0.0),
                                 kw.get('log_sum_w', 0.0), kw.get('log_prior_volume', 0.0), None, 0, keep)
            parr[i] = p
            kk = max(2, int(kw['kmax']))
            loge, var = np.empty(kk - 1), np.empty(kk - 1)
            rarr[i].logE, rarr[i].var_logE = _ptr(loge), _ptr(var)
            outs.append({'logE': loge, 'var_logE': var})
        _check(self._lib, self._lib.mce_evidence_batch(self._h, n, parr, rarr))
        for i, o in enumerate(outs):
            o['log_jacobian'] = rarr[i].log_jacobian
            o['stats'] = rarr[i].stats.as_dict()
        return outs

    def evidence_sweep(self, query, logl, weights, kmax, batch_rows, ref=None, ndim=None, W=None, W_ref=None, mu=None,
                       log_jacobian=0.0, log_sum_w=0.0, log_prior_volume=0.0, want_partials=False, query_rows=None,
                       ref_rows=None):
        """``mce_evidence_sweep``: the evidence of the first ``batch_rows[b]`` rows of ``query`` for every b (the batch
        loop of ``MCEvidence.py:1039-1045``) from one upload, one whitening and one pass of the filter.  Returns a list
        of result dicts (``logE``, ``var_logE``, ``log_jacobian``, ``stats``), one per batch size."""
        keep = []
        rows = np.ascontiguousarray(batch_rows, dtype=np.int64).reshape(-1)
        nb = int(rows.shape[0])
        p, _ = self._problem(query, ref, logl, weights, ndim, kmax, W, W_ref, mu, log_jacobian, log_sum_w,
                             log_prior_volume, None, FLAG_PARTIALS if want_partials else 0, keep, query_rows, ref_rows)
        kk = max(2, int(kmax))
        rarr = (Result * nb)()
        outs = []
        for b in range(nb):
            o = {'logE': np.empty(kk - 1), 'var_logE': np.empty(kk - 1)}
            rarr[b].logE, rarr[b].var_logE = _ptr(o['logE']), _ptr(o['var_logE'])
            if want_partials:
                o['partials'] = np.empty((kk - 1, 3))
                rarr[b].partials = _ptr(o['partials'])
            outs.append(o)
        _check(self._lib, self._lib.mce_evidence_sweep(self._h, C.byref(p), nb, rows.ctypes.data_as(c_int64_p), rarr))
        for b, o in enumerate(outs):
            o['log_jacobian'] = rarr[b].log_jacobian
            o['stats'] = rarr[b].stats.as_dict()
        return outs

    def evidence_batch_device(self, problems):
        """``mce_evidence_batch`` on DEVICE pointers: a list of dicts with ``query_ptr, n_query, ld_query, logl_ptr,
        weights_ptr, ndim, kmax`` (auto mode) and optionally ``log_sum_w, log_prior_volume``; the arrays live on this
        context's device (resident-input benchmark of the chain grid).  Returns a list of result dicts."""
        n = len(problems)
        parr = (Problem * n)()
        rarr = (Result * n)()
        outs = []
        for i, kw in enumerate(problems):
            p = parr[i]
            p.query, p.n_query, p.ld_query = int(kw['query_ptr']), int(kw['n_query']), int(kw['ld_query'])
            p.logl, p.weights = int(kw['logl_ptr']), int(kw['weights_ptr'])
            p.d, p.kmax, p.flags = int(kw['ndim']), int(kw['kmax']), FLAG_DEVICE_PTRS
            p.log_sum_w, p.log_prior_volume = float(kw.get('log_sum_w', 0.0)), float(kw.get('log_prior_volume', 0.0))
            kk = max(2, int(kw['kmax']))
            loge, var = np.empty(kk - 1), np.empty(kk - 1)
            rarr[i].logE, rarr[i].var_logE = _ptr(loge), _ptr(var)
            outs.append({'logE': loge, 'var_logE': var})
        _check(self._lib, self._lib.mce_evidence_batch(self._h, n, parr, rarr))
        for i, o in enumerate(outs):
            o['log_jacobian'] = rarr[i].log_jacobian
            o['stats'] = rarr[i].stats.as_dict()
        return outs

    def knn(self, query, n_neighbors, ref=None, ndim=None, W=None, W_ref=None, mu=None, q_range=None,
            exact_only=False, identity=False, query_rows=None, ref_rows=None):
        """``mce_knn``: ``(dknn, idx, stats)`` with sklearn's ``kneighbors`` contract.
        ``identity=True`` searches the coordinates as given (W = I, mu = 0)."""
        keep = []
        query_a = np.asarray(query)
        nd = int(query_a.shape[1] if ndim is None else ndim)
        if identity:
            W = np.eye(nd)
        p, nq = self._problem(query, ref, None, None, nd, 2, W, W_ref, mu, 0.0, 0.0, 0.0, q_range,
                              FLAG_EXACT_ONLY if exact_only else 0, keep, query_rows, ref_rows)
        dknn = np.empty((nq, int(n_neighbors)))
        idx = np.empty((nq, int(n_neighbors)), dtype=np.int64)
        st = Stats()
        _check(self._lib, self._lib.mce_knn(self._h, C.byref(p), int(n_neighbors), _ptr(dknn), _ptr(idx), C.byref(st)))
        return dknn, idx, st.as_dict()

    def evidence_device(self, query_ptr, n_query, ld_query, logl_ptr, weights_ptr, ndim, kmax, ref_ptr=0,
                        n_ref=0, ld_ref=0, q_range=None, log_sum_w=0.0, log_prior_volume=0.0, W=None, W_ref=None,
                        mu=None, log_jacobian=0.0, want_partials=False, want_logE=True, exact_only=False):
        """``mce_evidence`` on DEVICE pointers (integers, e.g. ``tensor.data_ptr()``) that live on this
        context's device; nothing is copied host<->device except the few result scalars.  Used for the
        resident-input benchmark and by the row-sharded multi-GPU path (``mcevidence_b200.dist``)."""
        keep = []
        p = Problem()
        p.query, p.n_query, p.ld_query = int(query_ptr), int(n_query), int(ld_query)
        if ref_ptr:
            p.ref, p.n_ref, p.ld_ref = int(ref_ptr), int(n_ref), int(ld_ref)
        p.logl, p.weights = int(logl_ptr), int(weights_ptr)
        if q_range is not None:
            p.q_begin, p.q_end = int(q_range[0]), int(q_range[1])
        p.d, p.kmax = int(ndim), int(kmax)
        p.flags = FLAG_DEVICE_PTRS | (FLAG_PARTIALS if want_partials else 0) | (FLAG_EXACT_ONLY if exact_only else 0)
        for name, arr, shape in (('W', W, (ndim, ndim)), ('W_ref', W_ref, (ndim, ndim)), ('mu', mu, (ndim,))):
            if arr is not None:
                arr = np.ascontiguousarray(arr, dtype=np.float64)
                if arr.shape != shape:
                    raise ValueError('%s must have shape %s' % (name, shape))
                keep.append(arr)
                setattr(p, name, _ptr(arr))
        p.log_jacobian, p.log_sum_w, p.log_prior_volume = float(log_jacobian), float(log_sum_w), float(log_prior_volume)
        kk = max(2, int(kmax))
        r = Result()
        out = {}
        if want_logE:
            out['logE'], out['var_logE'] = np.empty(kk - 1), np.empty(kk - 1)
            r.logE, r.var_logE = _ptr(out['logE']), _ptr(out['var_logE'])
        if want_partials:
            out['partials'] = np.empty((kk - 1, 3))
            r.partials = _ptr(out['partials'])
        _check(self._lib, self._lib.mce_evidence(self._h, C.byref(p), C.byref(r)))
        out.update({'log_jacobian': r.log_jacobian, 'stats': r.stats.as_dict()})
        return out

    def fma_peak(self, variant=1, iters=4096):
        tf, ms = C.c_double(), C.c_double()
        _check(self._lib, self._lib.mce_fma_peak(self._h, int(variant), int(iters), C.byref(tf), C.byref(ms)))
        return tf.value, ms.value


    def unit_peak(self, unit, iters=2048):
        """``mce_unit_peak``: (rate, ms) of FMNMX3 (0), tcgen05.ld with 8 / 4 warps (1 / 2), tf32 tcgen05.mma (3)."""
        rate, ms = C.c_double(), C.c_double()
        _check(self._lib, self._lib.mce_unit_peak(self._h, int(unit), int(iters), C.byref(rate), C.byref(ms)))
        return rate.value, ms.value


def combine_partials(parts):
    """``mce_combine_partials`` on an array of shape (nparts, nk, 3)."""
    lib = load_library()
    parts = np.ascontiguousarray(parts, dtype=np.float64)
    nparts, nk, _ = parts.shape
    out = np.empty((nk, 3))
    _check(lib, lib.mce_combine_partials(nparts, nk, parts.ctypes.data_as(c_double_p), out.ctypes.data_as(c_double_p)))
    return out


class Group:
    """``mce_group``: the GPUs of one box driven from this one process, one host thread per device inside the library
    (row-sharded :meth:`evidence`, chain-per-GPU :meth:`evidence_batch`).  ``devices``: a count, a list of ids, or None
    for every device."""

    def __init__(self, devices=None):
        self._lib = load_library()
        h = C.c_void_p()
        if devices is None:
            n, ids = 0, None
        elif isinstance(devices, int):
            n, ids = int(devices), None
        else:
            devs = [int(x) for x in devices]
            n, ids = len(devs), (C.c_int * len(devs))(*devs)
        _check(self._lib, self._lib.mce_group_create(n, ids, C.byref(h)))
        self._h = h
        self._helper = Context.__new__(Context)      # for _problem(): argument marshalling only, no device handle
        self._helper._lib, self._helper._h = self._lib, None

    @property
    def size(self):
        return int(self._lib.mce_group_size(self._h))

    @property
    def uses_nccl(self):
        return bool(self._lib.mce_group_uses_nccl(self._h))

    def set_option(self, name, value):
        _check(self._lib, self._lib.mce_group_set_option(self._h, name.encode(), int(value)))

    def close(self):
        if self._h is not None:
            self._lib.mce_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def evidence(self, query, logl, weights, kmax, ref=None, ndim=None, W=None, W_ref=None, mu=None, log_jacobian=0.0,
                 log_sum_w=0.0, log_prior_volume=0.0, want_partials=False):
        """``mce_group_evidence``: query rows sharded over the devices; same result dict as :meth:`Context.evidence`."""
        keep = []
        p, _ = self._helper._problem(query, ref, logl, weights, ndim, kmax, W, W_ref, mu, log_jacobian, log_sum_w,
                                     log_prior_volume, None, FLAG_PARTIALS if want_partials else 0, keep)
        kk = max(2, int(kmax))
        r = Result()
        out = {'logE': np.empty(kk - 1), 'var_logE': np.empty(kk - 1)}
        r.logE, r.var_logE = _ptr(out['logE']), _ptr(out['var_logE'])
        if want_partials:
            out['partials'] = np.empty((kk - 1, 3))
            r.partials = _ptr(out['partials'])
        _check(self._lib, self._lib.mce_group_evidence(self._h, C.byref(p), C.byref(r)))
        out.update({'log_jacobian': r.log_jacobian, 'stats': r.stats.as_dict()})
        return out

    def evidence_batch(self, problems):
        """``mce_group_evidence_batch``: independent problems dealt largest-first over the devices."""
        n = len(problems)
        keep, outs = [], []
        parr = (Problem * n)()
        rarr = (Result * n)()
        for i, kw in enumerate(problems):
            p, _ = self._helper._problem(kw['query'], kw.get('ref'), kw['logl'], kw['weights'], kw.get('ndim'), kw['kmax'],
                                         kw.get('W'), kw.get('W_ref'), kw.get('mu'), kw.get('log_jacobian', 0.0),
                                         kw.get('log_sum_w', 0.0), kw.get('log_prior_volume', 0.0), None, 0, keep)
            parr[i] = p
            kk = max(2, int(kw['kmax']))
            loge, var = np.empty(kk - 1), np.empty(kk - 1)
            rarr[i].logE, rarr[i].var_logE = _ptr(loge), _ptr(var)
            outs.append({'logE': loge, 'var_logE': var})
        _check(self._lib, self._lib.mce_group_evidence_batch(self._h, n, parr, rarr))
        for i, o in enumerate(outs):
            o['log_jacobian'] = rarr[i].log_jacobian
            o['stats'] = rarr[i].stats.as_dict()
        return outs


def finalize(partials, ndim, cross, n_query_total, log_sum_w, log_jacobian, log_prior_volume):
    """``mce_finalize``: partial triples -> (logE, var_logE)."""
    lib = load_library()
    partials = np.ascontiguousarray(partials, dtype=np.float64)
    nk = partials.shape[0]
    loge, var = np.empty(nk), np.empty(nk)
    _check(lib, lib.mce_finalize(partials.ctypes.data_as(c_double_p), nk, int(ndim), int(bool(cross)),
                                 int(n_query_total), float(log_sum_w), float(log_jacobian),
                                 float(log_prior_volume), loge.ctypes.data_as(c_double_p),
                                 var.ctypes.data_as(c_double_p)))
    return loge, var


def plan_split(n_query, n_ref, ndim, list_len=5, queries_per_cta=256, ctas_per_sm=4, wide=False, sm_count=148):
    """``mce_plan_split``: ``(nfull, chunks, cost)`` of the filter's work decomposition (host arithmetic)."""
    lib = load_library()
    nfull, chunks, cost = C.c_int(), C.c_int(), C.c_double()
    _check(lib, lib.mce_plan_split(int(sm_count), int(ndim), int(n_query), int(n_ref), int(list_len),
                                   int(queries_per_cta), int(ctas_per_sm), int(bool(wide)), C.byref(nfull),
                                   C.byref(chunks), C.byref(cost)))
    return nfull.value, chunks.value, cost.value


def read_chain_text(path, nthreads=0):
    """``mce_read_chain_text``: multi-threaded parse of a text chain into a float64 (rows, columns) array."""
    lib = load_library()
    data = c_double_p()
    nrows, ncols = C.c_int64(), C.c_int64()
    rc = lib.mce_read_chain_text(os.fsencode(path), int(nthreads), C.byref(data), C.byref(nrows), C.byref(ncols))
    if rc != 0:
        raise MCEError(rc, 'cannot parse %s as a rectangular whitespace-separated table' % path)
    try:
        arr = np.ctypeslib.as_array(data, shape=(nrows.value, ncols.value)).copy()
    finally:
        lib.mce_free_host(data)
    return arr


_default_ctx = {}


def default_context(device=0):
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx
