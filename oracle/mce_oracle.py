"""CPU restatement of the MCEvidence kNN-evidence hot path (TEST INFRASTRUCTURE ONLY).

This module is the parity oracle for the CUDA path.  It is *not* product code:
only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / reference
legs of ``bench.py`` may import it.  The product path (``mcevidence_b200``)
never routes through it and has no CPU fallback.

Every function cites the reference lines it restates (paths relative to
``/root/reference``).  The arithmetic of the neighbour search itself lives in
the un-vendored third-party dependency scikit-learn (unpinned in the reference,
``setup.py:31-37``; 1.9.0 in this image), reached from ``MCEvidence.py:1093-1104``;
``knn_sklearn`` calls it exactly as the reference does, and ``knn_brute_fp64``
is an independent exact restatement of what that call returns.

Parity pin: the reference ships no tests or golden vectors.
``oracle/gen_golden.py`` therefore imports the unmodified reference in the
build container (through the two-line ``DistanceMetric`` shim),
runs it on seeded synthetic chains and freezes its outputs under
``tests/golden/``; ``tests/test_oracle.py`` checks this restatement against
those frozen reference outputs.  The variance output has no reference
counterpart and is pinned by this file's definition only:
"parity unpinned" for ``var_logE``.
"""
from __future__ import annotations

import math

import numpy as np


# --------------------------------------------------------------------------
# A1: covariance, eigen-decomposition, Jacobian      (MCEvidence.py:851-882)
# --------------------------------------------------------------------------
def covariance_stats(s):
    """``np.cov`` (unweighted, ddof=1, no weights) + ``eig`` + ``sqrt(det)``.

    Restates ``get_covariance`` (MCEvidence.py:865-877).  A negative eigenvalue
    makes the reference crash later in ``math.sqrt`` (``:847``); here it is
    reported through ``posdef=False``.
    """
    cov = np.atleast_2d(np.cov(np.asarray(s, dtype=np.float64).T))
    e_val, e_vec = np.linalg.eig(cov)
    posdef = not (e_val < 0).any()
    jac = math.sqrt(np.linalg.det(cov)) if posdef else 1.0
    return {'cov': cov, 'eVal': e_val, 'eVec': e_vec, 'J': jac, 'posdef': posdef}


# --------------------------------------------------------------------------
# A2: whitening                                       (MCEvidence.py:842-849)
# --------------------------------------------------------------------------
def whiten(s, e_vec, e_val):
    """``s @ eVec`` then divide column i by ``sqrt(eVal[i])``; no centring."""
    y = np.dot(np.asarray(s, dtype=np.float64), e_vec)
    for i in range(y.shape[1]):
        y[:, i] = y[:, i] / math.sqrt(e_val[i])
    return y


# --------------------------------------------------------------------------
# A3/A4: neighbour search                            (MCEvidence.py:1093-1104)
# --------------------------------------------------------------------------
def knn_sklearn(ref, query, n_neighbors, n_jobs=-1):
    """The reference's own call pair: fit on ``ref``, query with ``query``.

    Same constructor arguments as MCEvidence.py:1093-1094 / 1100-1101.  In auto
    mode the caller passes ``ref is query`` and column 0 is the point itself.
    """
    from sklearn.neighbors import NearestNeighbors
    nbrs = NearestNeighbors(n_neighbors=n_neighbors, metric='euclidean', leaf_size=20,
                            algorithm='auto', n_jobs=n_jobs).fit(ref)
    return nbrs.kneighbors(query)


def sq_dist_sequential(q, r):
    """Squared Euclidean distance summed dimension by dimension in float64.

    This is the arithmetic of sklearn's kd-tree metric
    (``metrics/_dist_metrics.pxd.tp:39-49``: ``d += tmp * tmp`` in a loop) and
    the exact order the CUDA re-rank kernel uses, so that on identical whitened
    inputs the two agree bit for bit.  ``q`` is (nq, d), ``r`` is (nr, d);
    returns (nq, nr).
    """
    acc = np.zeros((q.shape[0], r.shape[0]), dtype=np.float64)
    for c in range(q.shape[1]):
        t = q[:, c][:, None] - r[:, c][None, :]
        acc += t * t
    return acc


def knn_brute_fp64(ref, query, k, exclude_self=False, self_offset=0, block=512):
    """Exact k nearest neighbours by exhaustive float64 search.

    Order is (distance, index) ascending, the order the CUDA path defines
    (sklearn's order among exact ties is unspecified).
    With ``exclude_self`` row ``i`` of ``query`` is reference row
    ``self_offset + i`` and is skipped; duplicates of it are kept.
    Returns ``(dist (nq, k) float64, idx (nq, k) int64)``.
    """
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    query = np.ascontiguousarray(query, dtype=np.float64)
    nq = query.shape[0]
    dist = np.empty((nq, k), dtype=np.float64)
    idx = np.empty((nq, k), dtype=np.int64)
    ar = np.arange(ref.shape[0], dtype=np.int64)
    for lo in range(0, nq, block):
        hi = min(nq, lo + block)
        d2 = sq_dist_sequential(query[lo:hi], ref)
        if exclude_self:
            rows = np.arange(hi - lo)
            d2[rows, self_offset + lo + rows] = np.inf
        kth = np.partition(d2, k - 1, axis=1)[:, k - 1]
        for i in range(hi - lo):
            # everything tied with the k-th competes; order is lexicographic (d2, index)
            cand = ar[d2[i] <= kth[i]]
            order = np.lexsort((cand, d2[i, cand]))[:k]
            idx[lo + i] = cand[order]
            dist[lo + i] = np.sqrt(d2[i, cand[order]])
    return dist, idx


# --------------------------------------------------------------------------
# A5/A6: volumes and the evidence estimator          (MCEvidence.py:1107-1131)
# --------------------------------------------------------------------------
def log_unit_ball_volume(ndim):
    """``log(pi^{d/2} / Gamma(1 + d/2))`` (MCEvidence.py:1110, true division)."""
    return 0.5 * ndim * math.log(math.pi) - math.lgamma(1.0 + 0.5 * ndim)


def evidence_from_dknn_linear(dknn, logl, weight, ndim, kmax, k0, sum_w, jac, log_pv,
                              scalar_loop=False):
    """Line-by-line restatement of the reference's linear-space reduction.

    ``dknn`` has ``kmax + 1`` columns as returned by ``kneighbors``.  Returns the
    full ``MLE`` row of length ``kmax`` (entry 0 is only filled when ``k0 == 0``),
    before the ``[1:]`` slice of MCEvidence.py:1155-1157.  ``scalar_loop`` runs the
    reference's per-element ``math.pow`` loop (``:1108-1110``) instead of the
    vectorised equivalent; only the timing differs.
    """
    import scipy.special as sp
    s_n = dknn.shape[0]
    logl_max = np.amax(logl)                                   # :1063
    fs = logl - logl_max                                       # :1064
    mle = np.zeros(kmax)
    cd = math.pow(math.pi, ndim / 2) / sp.gamma(1 + ndim / 2)
    volume = np.zeros(s_n)
    for k in range(k0, kmax):                                  # :1107
        if scalar_loop:
            for j in range(s_n):                               # :1108-1110
                volume[j] = math.pow(math.pi, ndim / 2) * math.pow(dknn[j, k], ndim) / sp.gamma(1 + ndim / 2)
        else:
            volume = cd * np.power(dknn[:, k], ndim)
        dotp = np.dot(volume / weight, np.exp(fs))             # :1117
        k_nn = k + 1 if k0 == 0 else k                         # :1120-1122
        amax = dotp / (s_n * k_nn + 1.0)                       # :1123
        mle[k] = math.log(sum_w * amax * jac) + logl_max - log_pv   # :1131
    return mle


def logsumexp_terms(dknn_col, logl, weight, ndim):
    """Per-sample log terms ``t_j = d log r_j + logL_j - log w_j`` (no constants)."""
    with np.errstate(divide='ignore'):
        return ndim * np.log(dknn_col) + logl - np.log(weight)


def evidence_from_dknn_log(dknn, logl, weight, ndim, kmax, k0, sum_w, log_jac, log_pv):
    """Log-space closed form of the same estimator.

    ``logE_k = log SumW + log J - log PV - log(S k_nn + 1) + log c_d + logsumexp_j t_jk``.
    Also returns the delta-method variance ``sum t~^2 / (sum t~)^2 - 1/S``
    (new output, no reference counterpart: SURVEY.md F6).
    Returns ``(mle (kmax,), var (kmax,))`` with the reference's column convention.
    """
    s_n = dknn.shape[0]
    mle = np.zeros(kmax)
    var = np.zeros(kmax)
    logc = log_unit_ball_volume(ndim)
    for k in range(k0, kmax):
        t = logsumexp_terms(dknn[:, k], logl, weight, ndim)
        m = np.max(t)
        e = np.exp(t - m)
        s1 = e.sum()
        s2 = (e * e).sum()
        k_nn = k + 1 if k0 == 0 else k
        mle[k] = (math.log(sum_w) + log_jac - log_pv - math.log(s_n * k_nn + 1.0)
                  + logc + m + math.log(s1))
        var[k] = s2 / (s1 * s1) - 1.0 / s_n
    return mle, var


# --------------------------------------------------------------------------
# whole path                                          (MCEvidence.py:1033-1157)
# --------------------------------------------------------------------------
def evidence(s1, logl1, w1, ndim, kmax, s2=None, priorvolume=1.0, adjusted_weights=None,
             knn='sklearn', n_jobs=-1, scalar_loop=False, return_parts=False,
             covtype='all', cov_from=None, sum_w=None):
    """Restatement of ``MCEvidence.evidence()`` for one batch.

    ``covtype='all'`` (the default of ``evidence()``): one covariance of s1 U s2, or of ``cov_from`` when
    given -- in a batch-size sweep the covariance is that of ALL samples while s1 is a prefix
    (MCEvidence.py:1034-1045, 860-861), and ``sum_w`` is then the weight sum of all of s1 (``:1126``).
    ``covtype='single'``: s1 whitened with its own eigen-system, s2 (cross mode) with its own, J of s1
    (``:1052-1054, 1080-1086``).

    ``s1`` (S, >=ndim) samples of partition s1, ``logl1 = -loglikes`` (the sign
    flip of ``arrays()``, MCEvidence.py:399), ``w1`` weights; ``s2`` the second
    partition for cross mode (``split=True``) or ``None`` for auto mode.
    ``knn`` selects the neighbour engine: ``'sklearn'`` (what the reference runs)
    or ``'brute'`` (exact float64).  Returns ``MLE[1:]`` (length ``max(2,kmax)-1``).
    """
    kmax = max(2, int(kmax))                                    # :694
    s1 = np.asarray(s1, dtype=np.float64)[:, :ndim]
    cross = s2 is not None
    if cross:
        s2 = np.asarray(s2, dtype=np.float64)[:, :ndim]
        allsamp = np.concatenate((s1, s2))                      # :413, :860-861
    else:
        allsamp = s1
    if covtype == 'single':
        cs = covariance_stats(s1)                               # :1052-1054
        cs_ref = covariance_stats(s2) if cross else cs          # :1080-1081 (J stays that of s1)
    else:
        cs = covariance_stats(allsamp if cov_from is None else np.asarray(cov_from, dtype=np.float64)[:, :ndim])
        cs_ref = cs                                             # :1034-1037
    y1 = whiten(s1, cs['eVec'], cs['eVal'])                     # :1057
    if cross:
        y2 = whiten(s2, cs_ref['eVec'], cs_ref['eVal'])         # :1086
        ref, k0 = y2, 0                                         # :1093-1096
    else:
        ref, k0 = y1, 1                                         # :1099-1101
    if knn == 'sklearn':
        dknn, idx = knn_sklearn(ref, y1, kmax + 1, n_jobs=n_jobs)   # :1104
    else:
        if cross:
            dknn, idx = knn_brute_fp64(ref, y1, kmax + 1)
        else:
            d, i = knn_brute_fp64(ref, y1, kmax, exclude_self=True)
            dknn = np.concatenate((np.zeros((len(y1), 1)), d), axis=1)
            idx = np.concatenate((np.arange(len(y1))[:, None], i), axis=1)
    aw = w1 if adjusted_weights is None else adjusted_weights
    if sum_w is None:
        sum_w = np.sum(aw)                                      # :1126
    mle = evidence_from_dknn_linear(dknn, logl1, w1, ndim, kmax, k0, sum_w, cs['J'],
                                    math.log(priorvolume), scalar_loop=scalar_loop)
    if return_parts:
        return mle[1:], {'dknn': dknn, 'idx': idx, 'cov': cs, 'y1': y1,
                         'y2': ref if cross else None, 'k0': k0, 'sum_w': sum_w}
    return mle[1:]                                              # :1155-1157
