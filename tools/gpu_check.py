"""Development harness run on the GPU box through gpurun: correctness of the CUDA path against the
oracle, the FP32 FMA microbenchmark and a sweep of filter configurations.  Results go to
gpurun_out/ (scratch).  Not part of the product; tests/ and bench.py are the judged artefacts.

    python tools/gpu_check.py [--skip-sweep] [--n 1000000]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mcevidence_b200 import _lib as L          # noqa: E402
from mcevidence_b200.synthetic import gaussian_chain  # noqa: E402
from oracle import mce_oracle as O             # noqa: E402

OUT = os.path.join(ROOT, 'gpurun_out')
os.makedirs(OUT, exist_ok=True)
log_lines = []


def log(*a):
    s = ' '.join(str(x) for x in a)
    print(s, flush=True)
    log_lines.append(s)


def check_knn(ctx, n, d, k, cross, seed, exact_only=False, tag=''):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, d)) * (1 + rng.random(d)) + rng.normal(size=d)
    ref = rng.normal(size=(n + 37, d)) * (1 + rng.random(d)) if cross else None
    t = time.time()
    dk, ix, st = ctx.knn(x, k, ref=ref, identity=True, exact_only=exact_only)
    tg = time.time() - t
    if cross:
        od, oi = O.knn_brute_fp64(ref, x, k)
    else:
        od, oi = O.knn_brute_fp64(x, x, k - 1, exclude_self=True)
        od = np.concatenate((np.zeros((n, 1)), od), axis=1)
        oi = np.concatenate((np.arange(n)[:, None], oi), axis=1)
    nbad_i = int((ix != oi).sum())
    maxd = float(np.abs(dk - od).max())
    ok = nbad_i == 0 and maxd == 0.0
    log('knn %-10s n=%-6d d=%-2d k=%-2d cross=%d exact_only=%d: idx_mismatch=%d max|dd|=%.3e rescanned=%d '
        'cfg=(tq%d,tr%d,pk%d,chunks%d,L%d) %.3fs %s' % (
            tag, n, d, k, cross, exact_only, nbad_i, maxd, st['rows_rescanned'], st['filter_tq'], st['filter_tr'],
            st['filter_packed'], st['ref_chunks'], st['list_len'], tg, 'OK' if ok else 'FAIL'))
    return ok


def check_evidence(ctx, n, d, kmax, cross, seed):
    chain, info = gaussian_chain(n, d, seed=seed, nextra=2, mean_shift=1.3)
    s = chain[:, 2:]
    logl = -chain[:, 1]
    w = chain[:, 0]
    if cross:
        h = n // 2
        s1, s2 = s[:h], s[h:]
        l1, w1 = logl[:h], w[:h]
    else:
        s1, s2, l1, w1 = s, None, logl, w
    pv = 3.5
    want = O.evidence(s1, l1, w1, d, kmax, s2=s2, priorvolume=pv, knn='brute')
    got = ctx.evidence(s1, l1, w1, kmax, ref=s2, ndim=d, log_sum_w=math.log(w1.sum()),
                       log_prior_volume=math.log(pv))
    err = float(np.abs(got['logE'] - want).max())
    ok = err < 1e-9
    log('evidence n=%-6d d=%-2d kmax=%d cross=%d: logE=%s max|diff|=%.3e var=%s rescanned=%d analytic=%.4f %s' % (
        n, d, kmax, cross, np.array2string(got['logE'], precision=8), err,
        np.array2string(got['var_logE'], precision=3), got['stats']['rows_rescanned'],
        info['lnZ'] - math.log(pv), 'OK' if ok else 'FAIL'))
    return ok


def sweep(ctx, n, d, kmax, cross, configs, results):
    chain, info = gaussian_chain(n * (2 if cross else 1), d, seed=0)
    s = np.ascontiguousarray(chain[:, 2:])
    logl = -chain[:, 1]
    w = chain[:, 0]
    if cross:
        s1, s2, l1, w1 = s[:n], s[n:], logl[:n], w[:n]
    else:
        s1, s2, l1, w1 = s, None, logl, w
    base = None
    for (tq, tr, pk, chunks) in configs:
        ctx.set_option('filter_tq', tq)
        ctx.set_option('filter_tr', tr)
        ctx.set_option('filter_packed', pk)
        ctx.set_option('ref_chunks', chunks)
        best = None
        for rep in range(2):
            got = ctx.evidence(s1, l1, w1, kmax, ref=s2, ndim=d, log_sum_w=math.log(w1.sum()))
            st = got['stats']
            if best is None or st['ms_filter'] < best['ms_filter']:
                best = st
        if base is None:
            base = got['logE']
        pairs = best['pairs']
        rate = pairs / (best['ms_filter'] * 1e-3)
        rec = {'n': n, 'd': d, 'kmax': kmax, 'cross': cross, 'tq': best['filter_tq'], 'tr': best['filter_tr'],
               'packed': best['filter_packed'], 'chunks': best['ref_chunks'], 'ctas_per_sm': best['filter_ctas_per_sm'],
               'smem': best['filter_smem'], 'grid': best['filter_grid'],
               'ms_filter': best['ms_filter'], 'ms_total': best['ms_total'], 'ms_rerank': best['ms_rerank'],
               'ms_exact': best['ms_exact'], 'ms_whiten': best['ms_whiten'], 'ms_cov': best['ms_cov'],
               'ms_reduce': best['ms_reduce'], 'ms_h2d': best['ms_h2d'],
               'rescanned': best['rows_rescanned'], 'pairs_per_s': rate, 'fma_tflops': 2 * d * rate / 1e12,
               'algo_tflops_3d': 3 * d * rate / 1e12, 'logE_diff_vs_first': float(np.abs(got['logE'] - base).max()),
               'analytic_diff': float(np.abs(got['logE'] - info['lnZ']).max())}
        results.append(rec)
        log('sweep n=%d d=%d cross=%d tq=%d tr=%d pk=%d chunks=%d ctas/sm=%d: filter %.2f ms (%.3e pairs/s, '
            '%.1f TFLOP/s FMA, %.1f algo) total %.2f ms rerank %.2f exact %.2f rescanned=%d dlogE=%.1e' % (
                n, d, cross, rec['tq'], rec['tr'], rec['packed'], rec['chunks'], rec['ctas_per_sm'], rec['ms_filter'],
                rate, rec['fma_tflops'], rec['algo_tflops_3d'], rec['ms_total'], rec['ms_rerank'], rec['ms_exact'],
                rec['rescanned'], rec['logE_diff_vs_first']))
    for name in ('filter_tq', 'filter_tr', 'filter_packed', 'ref_chunks'):
        ctx.set_option(name, -1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--skip-sweep', action='store_true')
    ap.add_argument('--skip-checks', action='store_true')
    ap.add_argument('--n', type=int, default=1000000)
    args = ap.parse_args()
    ctx = L.Context(0)
    allok = True
    results = {'checks_ok': None, 'peak': {}, 'sweep': []}
    try:
        for variant in (0, 1):
            best = 0.0
            for rep in range(3):
                tf, ms = ctx.fma_peak(variant, 8192)
                best = max(best, tf)
            results['peak']['ffma2' if variant else 'ffma'] = best
            log('fma_peak variant=%d: %.2f TFLOP/s (best of 3, last %.2f ms)' % (variant, best, ms))
        if not args.skip_checks:
            allok &= check_knn(ctx, 300, 3, 4, False, 1, exact_only=True, tag='exact')
            allok &= check_knn(ctx, 300, 3, 4, True, 2, exact_only=True, tag='exact')
            allok &= check_knn(ctx, 5, 2, 3, False, 3, tag='tiny')
            allok &= check_knn(ctx, 129, 1, 2, False, 4, tag='d1')
            for (n, d, k, cross) in [(2000, 5, 5, False), (3000, 10, 5, False), (3000, 10, 9, True),
                                     (2500, 17, 4, False), (2000, 27, 5, True), (1000, 40, 3, False),
                                     (20000, 8, 3, False), (1500, 64, 2, False), (4097, 6, 17, False)]:
                allok &= check_knn(ctx, n, d, k, cross, 10 + d)
            # duplicates / ties: every point repeated 4 times
            rng = np.random.default_rng(5)
            base = rng.normal(size=(500, 4))
            x = np.repeat(base, 4, axis=0)
            dk, ix, st = ctx.knn(x, 6, identity=True)
            od, oi = O.knn_brute_fp64(x, x, 5, exclude_self=True)
            okd = np.array_equal(dk[:, 1:], od) and np.array_equal(ix[:, 1:], oi)
            log('knn duplicates x4: rescanned=%d %s' % (st['rows_rescanned'], 'OK' if okd else 'FAIL'))
            allok &= okd
            for (n, d, kmax, cross) in [(4000, 5, 4, False), (6000, 10, 4, False), (6000, 10, 4, True),
                                        (5000, 20, 8, False), (4000, 27, 4, True), (3000, 3, 2, False)]:
                allok &= check_evidence(ctx, n, d, kmax, cross, 100 + d)
            results['checks_ok'] = bool(allok)
            log('ALL CHECKS', 'OK' if allok else 'FAILED')
        if not args.skip_sweep:
            n = args.n
            sweep(ctx, n, 10, 4, False,
                  [(4, 8, 1, -1), (4, 4, 1, -1), (8, 4, 1, -1), (6, 4, 1, -1), (8, 8, 1, -1), (2, 8, 1, -1),
                   (8, 4, 0, -1), (4, 8, 0, -1), (4, 8, 1, 1), (4, 8, 1, 4), (4, 8, 1, 16)], results['sweep'])
            sweep(ctx, n // 2, 20, 8, False,
                  [(4, 4, 1, -1), (2, 8, 1, -1), (4, 8, 1, -1), (3, 8, 1, -1), (4, 4, 0, -1), (2, 8, 0, -1)],
                  results['sweep'])
            sweep(ctx, n // 2, 27, 4, True,
                  [(4, 4, 1, -1), (2, 8, 1, -1), (3, 4, 1, -1), (2, 4, 1, -1), (3, 8, 1, -1), (2, 8, 0, -1)],
                  results['sweep'])
            sweep(ctx, 15000, 7, 2, False, [(4, 8, 1, -1), (4, 4, 1, -1), (8, 4, 1, -1)], results['sweep'])
    except Exception:
        log('EXCEPTION\n' + traceback.format_exc())
        allok = False
    with open(os.path.join(OUT, 'gpu_check.json'), 'w') as fh:
        json.dump(results, fh, indent=1)
    with open(os.path.join(OUT, 'gpu_check.log'), 'w') as fh:
        fh.write('\n'.join(log_lines) + '\n')
    return 0 if allok else 1


if __name__ == '__main__':
    sys.exit(main())
