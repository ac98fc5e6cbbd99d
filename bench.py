#!/usr/bin/env python
"""Benchmark of the kNN-evidence hot path (BASELINE.json metric) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (covariance -> whitening -> exact k-th nearest neighbours ->
log-volume reduction -> per-k logE) over one synthetic chain.  Default workload: BASELINE cfg2,
synthetic correlated Gaussian chain N=1e6, d=10, kmax=4, auto mode.

* ``value``  pair-distances/s with the chain already resident in HBM, timed with CUDA events on the
  launching stream, max over ranks (query rows are sharded over ranks: strong scaling).
* ``e2e``    the same metric through the public host API (``Context.evidence`` /
  ``dist.evidence_sharded``): pinned host arrays in, logE out, copies inside the timed region.
* ``roofline``  the dominant kernel (K2 filter) against the FP32 FMA pipe, whose peak is measured in
  this run with an FMA microbenchmark (MEASURED_PEAKS.json has HBM and tensor peaks only).
* ``cpu_baseline``  the oracle port of the reference path (sklearn, all host cores) on a bounded
  sample; ``--impl reference`` times the same port as its own arm.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, d, kmax, cross, description)
    'cfg2': (1000000, 10, 4, False, 'cfg2: synthetic correlated Gaussian chain N=1e6 d=10 kmax=4 auto'),
    'cfg3': (1000000, 27, 4, True, 'cfg3: cross evidence d=27, 2e6 rows split 1e6/1e6, kmax=4'),
    'cfg5': (16000000, 20, 8, False, 'cfg5: N=1.6e7 d=20 kmax=8 auto (scaling run)'),
    'cfg5s': (2000000, 20, 8, False, 'cfg5 shape at N=2e6 d=20 kmax=8 auto'),
    'small': (100000, 10, 4, False, 'N=1e5 d=10 kmax=4 auto'),
    # BASELINE cfg1: one CosmoMC root of the Planck grid -- 4 chain files of ~6.7k rows, 30 % burn-in, the 6 cosmological
    # parameters of the base model out of 10 sampled columns, integer weights, kmax=2 (`MCEvidence.py <root> -k 2 -b 0.3`)
    'cfg1': (4 * 4690, 6, 2, False, 'cfg1: Planck-shaped base-model root, 4 chains x 6.7k rows after 30% burn-in, ndim=6 of 10 columns, kmax=2'),
}
# BASELINE cfg4: every (model, data) cell of planck_fullgrid_R2/AllChains as one batch of independent problems
GRID_WORKLOADS = {'cfg4': ('AllChains', 'cfg4: 175 (model, data) cells of planck_fullgrid_R2/AllChains, synthetic chains of the '
                                        'tabulated (ndim, N), kmax=2, one mce_evidence_batch per step'),
                  'cfg4s': ('SingleChains', 'cfg4 per-chain variant: 700 chains of planck_fullgrid_R2/SingleChains, kmax=2')}
FP32_PEAK_THEORETICAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.4: SMs x lanes x 2 FLOP x max clock
# dram__bytes_read.sum + dram__bytes_write.sum of one K2 launch from the ncu --set full capture of this command
# (profiles/r01_filter_ncu_raw_final.csv: 148.7 MB read + 31.2 MB written; algorithmic: 44 MB tiles + 80 MB
# queries + ~24 MB candidate lists -- the reference set is streamed from L2, not from HBM)
# tcgen05 engine (profiles/r02_tc_filter_ncu_raw_cfg2_final.csv): 738.5 MB read + 43.5 MB written; algorithmic: 64 MB image +
# 80 MB queries + ~25 MB candidate lists -- the image is re-read from HBM when it falls out of L2 between query blocks
NCU_DRAM_BYTES_PER_LAUNCH = {('cfg2', 1, False): 179.9e6, ('cfg2', 1, True): 781.9e6}   # (workload, gpus, tcgen05 engine?)
METRIC = 'knn_pair_dists_per_s'
UNIT = 'pair-dists/s'


def make_workload(name, seed=0):
    from mcevidence_b200.synthetic import gaussian_chain
    n, d, kmax, cross, desc = WORKLOADS[name]
    if name == 'cfg1':
        chain, info = gaussian_chain(n, d, seed=seed, nextra=4, integer_weights=True, inflate=1.0, mean_shift=0.3)
        return dict(s1=chain[:, 2:], s2=None, logl=-chain[:, 1], w=np.ascontiguousarray(chain[:, 0]), d=d, kmax=kmax,
                    cross=False, desc=desc, lnZ=info['lnZ'])
    chain, info = gaussian_chain(n * (2 if cross else 1), d, seed=seed)
    s = np.ascontiguousarray(chain[:, 2:])
    logl, w = -chain[:, 1], np.ascontiguousarray(chain[:, 0])
    if cross:
        return dict(s1=s[:n], s2=s[n:], logl=logl[:n], w=w[:n], d=d, kmax=kmax, cross=True, desc=desc, lnZ=info['lnZ'])
    return dict(s1=s, s2=None, logl=logl, w=w, d=d, kmax=kmax, cross=False, desc=desc, lnZ=info['lnZ'])


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix='mce_clocks_', suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, smax, power, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                t = [x.strip() for x in line.split(',')]
                if len(t) < 9:
                    continue
                try:
                    sm.append(float(t[1])); smax.append(float(t[2])); power.append(float(t[3]))
                except ValueError:
                    continue
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), t[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            pw = np.array(power)
            load = np.array(sm)[pw >= 0.5 * pw.max()] if pw.max() > 0 else np.array(sm)
            out.update({'sm_mhz': float(np.median(load)), 'sm_max_mhz': float(max(smax)), 'reasons': sorted(reasons),
                        'samples': len(sm), 'power_w_max': float(pw.max())})
        return out


def host_info():
    model = ''
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                model = line.split(':', 1)[1].strip()
                break
    except Exception:
        pass
    import sklearn
    import scipy
    return {'cores': os.cpu_count(), 'cpu_model': model, 'sklearn': sklearn.__version__, 'numpy': np.__version__,
            'scipy': scipy.__version__}


def config_for(name, world):
    """The `config` object of a workload: identical in both arms (ours / --impl reference)."""
    n, d, kmax, cross, desc = WORKLOADS[name]
    return {'workload': desc, 'n': n, 'n_ref': n, 'ndim': d, 'kmax': kmax, 'mode': 'cross' if cross else 'auto',
            'parallelism': 'our arm: query rows sharded over %d GPU(s), reference set replicated; reference arm: '
                           'sklearn n_jobs=-1 on the host cores of rank 0' % world,
            'l2': 'our arm: 512 MiB memset between timed steps (not timed); working set > L2'}


def cpu_reference_once(wl, n_s):
    """One run of the reference's own CPU path on the first n_s rows of the workload: the UNMODIFIED
    MCEvidence(...).evidence(nproc=-1) staged under oracle/_ref (oracle/ref_loader.py) when it travelled to this
    box, else the oracle port (numpy + sklearn called as MCEvidence.py:1033-1131 does).  Returns (seconds, logE, kind)."""
    from oracle import ref_loader
    d, kmax = wl['d'], wl['kmax']
    if ref_loader.available():
        import logging
        R = ref_loader.load()
        parts = [np.column_stack((wl['w'][:n_s], -wl['logl'][:n_s], wl['s1'][:n_s]))]
        if wl['cross']:      # the reference splits one chain at random into s1 / s2 (MCEvidence.py:221-249)
            parts.append(np.column_stack((wl['w'][:n_s], -wl['logl'][:n_s], wl['s2'][:n_s])))
        chain = np.concatenate(parts)
        logging.disable(logging.CRITICAL)
        try:
            np.random.seed(0)
            t = time.perf_counter()
            mle = R.MCEvidence([chain], kmax=kmax, ndim=d, split=wl['cross'], verbose=0).evidence(nproc=-1)
            dt = time.perf_counter() - t
        finally:
            logging.disable(logging.NOTSET)
        return dt, np.asarray(mle), 'reference'
    from oracle import mce_oracle as O
    s2 = wl['s2'][:n_s] if wl['cross'] else None
    t = time.perf_counter()
    mle = O.evidence(wl['s1'][:n_s], wl['logl'][:n_s], wl['w'][:n_s], d, kmax, s2=s2, knn='sklearn', n_jobs=-1,
                     scalar_loop=True)
    return time.perf_counter() - t, mle, 'port'


def cpu_plan(wl, target_s, n_full):
    """Calibrate on a small sample and size the per-step sample for ~target_s seconds of CPU work (full size when
    target_s <= 0).  Returns (n_sample, calibration seconds)."""
    if target_s <= 0:
        return n_full, 0.0
    n0 = min(20000, n_full)
    cpu_reference_once(wl, min(4000, n_full))                 # thread pools, imports
    t0, _, _ = cpu_reference_once(wl, n0)
    guess = 1.3 if wl['d'] <= 15 else 2.0                     # only to size the sample; the reported exponent is fitted
    n_s = int(n0 * (target_s / max(t0, 1e-3)) ** (1.0 / guess))
    return int(max(n0, min(n_full, n_s))), t0


def cpu_extrapolate(wl, n_s, t_s, n_full):
    """Seconds per logE at the full size from the measured sample: t_full = t_s (n_full / n_s)^p with p fitted IN
    THIS RUN from a second timing at n_s / 2 (kd_tree is ~N^1.3, brute N^2)."""
    if n_s >= n_full:
        return t_s, None, None
    n_h = max(2000, n_s // 2)
    t_h = min(cpu_reference_once(wl, n_h)[0] for _ in range(2))
    p_fit = math.log(t_s / t_h) / math.log(n_s / n_h)
    # The local slope at small N overstates the cost of the kd-tree at large N, which would flatter the GPU/CPU ratio:
    # the one full-size run of the unmodified reference (cfg2, N = 1e6: 243.6 s, profiles/r02_reference_arm_full_size_cfg2.json)
    # against its 139k-row samples on the same box (27.0 s) gives N^1.115 from there to 1e6, less than the N^1.3 that
    # SURVEY section 6 measured between 25k and 400k rows.  Use the smaller of the fitted slope and that measured law
    # (N^2 for the brute-force branch, ndim > 15).  Both are reported.
    ceiling = 1.12 if wl['d'] <= 15 else 2.0
    p = float(min(ceiling, max(1.0, p_fit)))
    return t_s * (n_full / n_s) ** p, p, (n_h, t_h, float(p_fit))


def cpu_baseline(wl, target_s=20.0):
    n_full = len(wl['s1'])
    n_s, _ = cpu_plan(wl, target_s, n_full)
    t, _, kind = cpu_reference_once(wl, n_s)
    t_full, p, half = cpu_extrapolate(wl, n_s, t, n_full)
    hi = host_info()
    algo = 'kd_tree' if wl['d'] <= 15 else 'brute'
    n_ref = len(wl['s2']) if wl['cross'] else n_full
    sample = ('%s MCEvidence(...).evidence(nproc=-1), sklearn %s (%s), %d cores: first %d of %d rows of the same chain in '
              '%.2f s' % ('unmodified reference (oracle/_ref)' if kind == 'reference' else 'oracle port of', hi['sklearn'],
                          algo, hi['cores'], n_s, n_full, t))
    if p is not None:
        sample += ('; %d rows in %.2f s -> slope N^%.2f fitted in this run, N^%.2f used (capped at the large-N law); value = '
                   'N_full^2 / t_full with t_full = %.1f s EXTRAPOLATED' % (half[0], half[1], half[2], p, t_full))
    return {'value': float(n_full) * float(n_ref) / t_full, 'unit': UNIT, 'cores': hi['cores'], 'kind': kind,
            'sample': sample, 'seconds': t, 'n_sample': n_s, 'cpu_model': hi['cpu_model'], 'extrapolated': p is not None,
            'fit_exponent': half[2] if half else None, 'used_exponent': p, 's_per_logE_full': t_full}


# ---------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU path on the host cores of rank 0 -- the unmodified
    MCEvidence.evidence() staged under oracle/_ref (kind "reference"), else the oracle port.  Each step is a bounded
    sample of the workload; `value` is the like-for-like figure for the FULL configuration, pair-distances of the
    full problem / seconds per logE at full size, extrapolated with an exponent fitted in this run
    (--cpu-seconds 0 runs the full size instead: minutes per step)."""
    if rank != 0:
        return 0
    wl = make_workload(args.workload)
    n_full = len(wl['s1'])
    n_ref = len(wl['s2']) if wl['cross'] else n_full
    budget = 200.0 / max(1, args.steps + args.warmup + 1)
    target = 0.0 if args.cpu_seconds <= 0 else min(25.0, budget, args.cpu_seconds)
    n_s, _ = cpu_plan(wl, target, n_full)
    for _ in range(args.warmup if n_s < n_full else 0):
        cpu_reference_once(wl, n_s)
    times, kind = [], 'port'
    for _ in range(args.steps):
        t, mle, kind = cpu_reference_once(wl, n_s)
        times.append(t)
    t_step = float(np.mean(times))
    t_full, p, half = cpu_extrapolate(wl, n_s, t_step, n_full)
    hi = host_info()
    val = float(n_full) * float(n_ref) / t_full
    algo = 'kd_tree' if wl['d'] <= 15 else 'brute'
    sample = ('%s MCEvidence(...).evidence(nproc=-1), sklearn %s (%s), %d cores; each step = first %d of %d rows' % (
        'unmodified reference (oracle/_ref)' if kind == 'reference' else 'oracle port of', hi['sklearn'], algo, hi['cores'],
        n_s, n_full))
    if p is not None:
        sample += ('; %.2f s per step, %d rows in %.2f s -> slope N^%.2f fitted in this run, N^%.2f used (capped at the large-N '
                   'law); value is for the full configuration: N_full^2 / t_full, t_full = %.1f s EXTRAPOLATED' % (
                       t_step, half[0], half[1], half[2], p, t_full))
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * t_step, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': config_for(args.workload, max(1, args.gpus)),
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': hi['cores'], 'kind': kind, 'sample': sample,
                             'cpu_model': hi['cpu_model']},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'extrapolated': p is not None, 'fit_exponent': half[2] if half else None, 'used_exponent': p, 'sample_rows': n_s, 's_per_logE': t_full,
            's_per_logE_sample': t_step, 'logE_sample': [float(x) for x in np.ravel(mle)], 'gpu_launches': 0}
    print(json.dumps(line), flush=True)
    return 0



# ---------------------------------------------------------------------------------------------
def run_grid_workload(args, rank, world, local_rank):
    """cfg4: the whole Planck grid as ONE batch of independent auto-mode problems per step (planck_mcevidence.py:306-359).
    `value` = sum of the cells' pair distances / device time of mce_evidence_batch with every chain resident in HBM;
    `e2e` = the same batch from host arrays (grid.evidence_many); with N > 1 ranks the cells are dealt largest first
    over the GPUs and no data-path collective is involved.  --impl reference: the unmodified reference cell by cell on
    the host cores (a bounded sample of cells, scaled by pair count)."""
    kind, desc = GRID_WORKLOADS[args.workload]
    shapes = json.load(open(os.path.join(ROOT, 'mcevidence_b200', 'planck_grid_shapes.json')))[kind]
    cells = [(nd, n) for (_, _, nd, ns) in shapes for n in ns]
    from mcevidence_b200.synthetic import gaussian_chain
    pairs_total = float(sum(float(n) ** 2 for _, n in cells))
    config = {'workload': desc, 'cells': len(cells), 'rows': int(sum(n for _, n in cells)), 'pairs': pairs_total,
              'ndim': '6-9', 'kmax': 2, 'mode': 'auto',
              'parallelism': 'our arm: cells dealt largest-first over %d GPU(s), up to 8 problems in flight per GPU; reference arm: '
                             'cell by cell, sklearn n_jobs=-1 on the host cores of rank 0' % max(1, world),
              'l2': 'our arm: 512 MiB memset between timed steps (not timed)'}

    def make(i):
        nd, n = cells[i]
        return gaussian_chain(n, nd, seed=i, inflate=1.0, integer_weights=True, nextra=3)[0]

    def cpu_cells(target_s):
        """The unmodified reference on a spread of cells until ~target_s seconds are spent; returns (seconds, pairs, n)."""
        from oracle import ref_loader
        import logging
        order = np.argsort([-n for _, n in cells])
        pick = [int(order[i]) for i in np.linspace(0, len(cells) - 1, 12).astype(int)]
        R = ref_loader.load() if ref_loader.available() else None
        t_sum, p_sum, done = 0.0, 0.0, 0
        for i in pick:
            c = make(i)
            t0 = time.perf_counter()
            if R is not None:
                logging.disable(logging.CRITICAL)
                try:
                    R.MCEvidence([c], kmax=2, ndim=cells[i][0], verbose=0).evidence(nproc=-1)
                finally:
                    logging.disable(logging.NOTSET)
            else:
                from oracle import mce_oracle as O
                O.evidence(c[:, 2:], -c[:, 1], c[:, 0], cells[i][0], 2, knn='sklearn', scalar_loop=True)
            t_sum += time.perf_counter() - t0
            p_sum += float(cells[i][1]) ** 2
            done += 1
            if t_sum > target_s:
                break
        return t_sum, p_sum, done, 'reference' if R is not None else 'port'

    if args.impl == 'reference':
        if rank != 0:
            return 0
        hi = host_info()
        ts = []
        for _ in range(max(1, args.steps)):
            t_sum, p_sum, done, kind_ = cpu_cells(min(25.0, 200.0 / max(1, args.steps + 1)))
            ts.append(t_sum)
        # the kd-tree cost per cell is ~N^1.3, not N^2: scale the sample by sum(N^1.3), which favours the reference
        w_all = float(sum(float(n) ** 1.3 for _, n in cells))
        order = np.argsort([-n for _, n in cells])
        pick = [int(order[i]) for i in np.linspace(0, len(cells) - 1, 12).astype(int)][:done]
        w_s = float(sum(float(cells[i][1]) ** 1.3 for i in pick))
        t_full = float(np.mean(ts)) * w_all / w_s
        val = pairs_total / t_full
        sample = ('%s MCEvidence([chain], kmax=2, ndim).evidence(nproc=-1) on %d of the %d cells (a spread over sizes), %.2f s; '
                  'scaled to the whole grid by sum N^1.3 (kd_tree law): %.1f s per grid EXTRAPOLATED' % (
                      'unmodified reference (oracle/_ref)' if kind_ == 'reference' else 'oracle port of', done, len(cells),
                      float(np.mean(ts)), t_full))
        print(json.dumps({'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
                          'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * float(np.mean(ts)),
                          'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                          'config': config, 'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': hi['cores'], 'kind': kind_,
                                                             'sample': sample, 'cpu_model': hi['cpu_model']},
                          'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                          'extrapolated': True, 's_per_grid': t_full, 'gpu_launches': 0}), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    from mcevidence_b200 import _lib
    from mcevidence_b200.grid import assign_largest_first, evidence_many
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=device)
    ctx = _lib.default_context(local_rank)
    ctx.set_option('batch_lanes', 8)
    chains = [make(i) for i in range(len(cells))]
    ndims = [nd for nd, _ in cells]
    mine = assign_largest_first([n for _, n in cells], world)[rank]
    # resident copies of this rank's cells: parameter columns contiguous, logL and weights as vectors
    res, probs = [], []
    for i in mine:
        c = chains[i]
        s = torch.from_numpy(np.ascontiguousarray(c[:, 2:2 + ndims[i]])).to(device)
        ll = torch.from_numpy(np.ascontiguousarray(-c[:, 1])).to(device)
        w = torch.from_numpy(np.ascontiguousarray(c[:, 0])).to(device)
        res.append((s, ll, w))
        probs.append({'query_ptr': s.data_ptr(), 'n_query': len(c), 'ld_query': ndims[i], 'logl_ptr': ll.data_ptr(),
                      'weights_ptr': w.data_ptr(), 'ndim': ndims[i], 'kmax': 2, 'log_sum_w': math.log(c[:, 0].sum())})
    torch.cuda.synchronize(device)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)
    stream = torch.cuda.current_stream(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(args.warmup):
        outs = ctx.evidence_batch_device(probs) if probs else []
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t_ms, launches = 0.0, 0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)        # the batch runs on the library's own lanes: bracket it with a device-wide wait
        t0 = time.perf_counter()
        outs = ctx.evidence_batch_device(probs) if probs else []
        torch.cuda.synchronize(device)
        t_ms += 1e3 * (time.perf_counter() - t0)
        launches += sum(o['stats']['launches'] for o in outs)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([t_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    t_ms = float(tmax.item())
    value = args.steps * pairs_total / (t_ms * 1e-3)
    # e2e: host arrays in, logE out
    for _ in range(2):
        evidence_many(chains, kmax=2, ndim=ndims, device=local_rank)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loge, var, _ms = evidence_many(chains, kmax=2, ndim=ndims, device=local_rank)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    te = float(te.item())
    if rank == 0:
        flop = float(sum(3.0 * nd * float(n) ** 2 for nd, n in cells))
        h2d = int(sum(n * (nd + 3) + 2 * n for nd, n in cells)) * 8
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': t_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                'dtype': 'f32 filter, f64 re-rank/reduce', 'data': 'synthetic', 'config': config,
                'e2e': {'value': args.steps * pairs_total / te, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': len(cells) * 8 * 8, 's_per_grid': te / args.steps},
                'gpu_launches': launches, 'clocks': clocks,
                'roofline': {'bound': 'fp32_fma', 'achieved': args.steps * flop / (t_ms * 1e-3) / 1e12,
                             'peak': FP32_PEAK_THEORETICAL_TFLOPS, 'unit': 'TFLOP/s',
                             'frac': args.steps * flop / (t_ms * 1e-3) / 1e12 / FP32_PEAK_THEORETICAL_TFLOPS, 'traffic': None,
                             'kernel': 'whole batch (covariance, whitening, filter, re-rank, reduction of every cell; the filter '
                                       'kernels of up to 8 cells overlap, so no single launch can be singled out)',
                             'note': 'algorithmic 3 d FLOP per pair summed over the cells / wall time of the resident batch, '
                                     'against the theoretical FP32 FMA peak'},
                's_per_grid': t_ms * 1e-3 / args.steps, 'logE_first_cells': [float(x[0]) for x in loge[:4]]}
        if world == 1 and not args.no_cpu_baseline:
            t_sum, p_sum, done, kind_ = cpu_cells(args.cpu_seconds)
            hi = host_info()
            w_all = float(sum(float(n) ** 1.3 for _, n in cells))
            order = np.argsort([-n for _, n in cells])
            pick = [int(order[i]) for i in np.linspace(0, len(cells) - 1, 12).astype(int)][:done]
            t_full = t_sum * w_all / float(sum(float(cells[i][1]) ** 1.3 for i in pick))
            line['cpu_baseline'] = {'value': pairs_total / t_full, 'unit': UNIT, 'cores': hi['cores'], 'kind': kind_,
                                    'sample': '%d of %d cells in %.2f s, scaled by sum N^1.3 to %.1f s per grid EXTRAPOLATED' % (
                                        done, len(cells), t_sum, t_full), 'cpu_model': hi['cpu_model'], 's_per_grid': t_full}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------
def _measured_tf32_peak():
    """TF32 dense peak for the roofline of the tcgen05 engine: half of the measured bf16 figure of
    MEASURED_PEAKS.json (tcgen05 issues K = 8 TF32 against K = 16 bf16 per instruction), else half of the
    profiling guide's nominal 2250 TFLOP/s."""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            bf16 = float(json.load(fh)['bf16_tflops'])
        return bf16 / 2.0, 'MEASURED_PEAKS.json bf16_tflops / 2 (TF32 MMA has half the K per instruction)'
    except Exception:
        return 1125.0, 'nominal 2250 TFLOP/s bf16 / 2 (MEASURED_PEAKS.json not found)'


def _alu_floor(ctx, pairs, t_k2):
    """Time the FMNMX3 max tree of the tcgen05 epilogue would need alone at the measured FMNMX3 rate."""
    try:
        rate = max(ctx.unit_peak(0, 4096)[0] for _ in range(3))          # lane-operations per second
        floor_s = pairs * 17.0 / 32.0 / rate
        return {'fmnmx3_lane_ops_per_s': rate, 'lane_ops_per_pair': 17.0 / 32.0, 'floor_ms': 1e3 * floor_s,
                'frac': floor_s / t_k2}
    except Exception as exc:                                               # diagnostics only
        return {'error': str(exc)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS) + sorted(GRID_WORKLOADS))
    ap.add_argument('--filter-engine', default=None, choices=['ffma', 'tc', 'auto'],
                    help="K2 engine: 'ffma' (FP32 FMA pipe), 'tc' (tcgen05 tensor-core filter) or 'auto' (chosen per problem, "
                         "the library default); not given = the library default or MCE_FILTER_ENGINE")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=20.0,
                    help='target seconds of CPU work for the cpu_baseline sample / per step of the reference arm; '
                         '0 = the full configuration (minutes per step at cfg2)')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-API leg (long scaling runs only)')
    ap.add_argument('--allow-short-warmup', action='store_true',
                    help='do not raise --warmup to 3 (minutes-long steps of the cfg5 scaling run only)')
    args = ap.parse_args()
    if args.impl == 'ours' and not args.allow_short_warmup:
        args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.workload in GRID_WORKLOADS:
        return run_grid_workload(args, rank, world, local_rank)
    if args.impl == 'reference':
        return run_reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    from mcevidence_b200 import _lib
    from mcevidence_b200.dist import evidence_sharded, shard_range

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'      # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group('nccl', device_id=device)
    ctx = _lib.default_context(local_rank)
    stream = torch.cuda.current_stream(device)
    ctx.set_stream(stream.cuda_stream)
    if args.filter_engine is not None:
        ctx.set_option('filter_engine', {'ffma': 0, 'tc': 1, 'auto': 2}[args.filter_engine])

    # e2e at N > 1: every rank holds the (same, seeded) chain on its host, as the ranks of the reference's MPI driver each
    # read their files; a rank uploads its own block of rows and the blocks are all-gathered (dist.evidence_sharded,
    # host_on_every_rank).  The resident leg still starts from rank 0's copy, broadcast once before the timed region.
    every_rank_host = world > 1 and not args.no_e2e
    wl = make_workload(args.workload) if (rank == 0 or every_rank_host) else None
    meta = [None]
    if rank == 0:
        meta = [{k: wl[k] for k in ('d', 'kmax', 'cross', 'desc', 'lnZ')}]
        meta[0]['n'] = len(wl['s1'])
        meta[0]['n_ref'] = len(wl['s2']) if wl['cross'] else len(wl['s1'])
        meta[0]['sum_w'] = float(wl['w'].sum())
    if world > 1:
        dist.broadcast_object_list(meta, src=0)
    m = meta[0]
    n, n_ref, d, kmax, cross, sum_w = m['n'], m['n_ref'], m['d'], m['kmax'], m['cross'], m['sum_w']

    # ---- pinned host copies (rank 0) and resident device copies (every rank) ------------------
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t
    host = {}
    if rank == 0 or every_rank_host:
        host = {'samples': pinned(np.ascontiguousarray(wl['s1'][:, :wl['d']])), 'logl': pinned(wl['logl']), 'weights': pinned(wl['w'])}
        if cross:
            host['ref'] = pinned(wl['s2'])
    res = {}
    for key, shape in (('samples', (n, d)), ('logl', (n,)), ('weights', (n,))) + ((('ref', (n_ref, d)),) if cross else ()):
        t = host[key].to(device, non_blocking=True) if rank == 0 else torch.empty(shape, dtype=torch.float64, device=device)
        if world > 1:
            dist.broadcast(t, src=0)
        res[key] = t
    torch.cuda.synchronize(device)

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2

    def step_resident():
        if world == 1:
            r = res.get('ref')
            return ctx.evidence_device(res['samples'].data_ptr(), n, d, res['logl'].data_ptr(), res['weights'].data_ptr(),
                                       d, kmax, ref_ptr=r.data_ptr() if cross else 0, n_ref=n_ref, ld_ref=d,
                                       log_sum_w=math.log(sum_w))
        return evidence_sharded(None, None, None, kmax, resident=res, sum_w=sum_w, ctx=ctx, device=device)

    def step_e2e():
        if world == 1:
            return ctx.evidence(host['samples'].numpy(), host['logl'].numpy(), host['weights'].numpy(), kmax,
                                ref=host['ref'].numpy() if cross else None, ndim=d, log_sum_w=math.log(sum_w))
        return evidence_sharded(host['samples'].numpy(), host['logl'].numpy(), host['weights'].numpy(), kmax,
                                ref=host['ref'].numpy() if cross else None, ndim=d, sum_w=sum_w, ctx=ctx, device=device,
                                host_on_every_rank=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- FP32 FMA peak, measured live ----------------------------------------------------------
    peak = {}
    for variant, name in ((0, 'ffma'), (1, 'ffma2')):
        peak[name] = max(ctx.fma_peak(variant, 8192)[0] for _ in range(3))
    fma_peak = max(peak.values())

    # ---- value: inputs resident, CUDA events on the launching stream --------------------------
    for _ in range(args.warmup):
        out = step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches, ms_filter, rescanned, stats = 0, [], 0, None
    for i in range(args.steps):
        flush.zero_()                                  # L2 flush between timed iterations (not timed)
        ev[i][0].record(stream)
        out = step_resident()
        ev[i][1].record(stream)
        stats = out['stats']
        launches += stats['launches'] + (2 if world > 1 else 0)
        ms_filter.append(stats['ms_filter'])
        rescanned += stats['rows_rescanned']
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    tmax = torch.tensor([t_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    t_ms = float(tmax.item())
    loge = out['logE']
    pairs_total = float(n) * float(n_ref)
    value = args.steps * pairs_total / (t_ms * 1e-3)

    # ---- e2e: public host API, pinned host arrays in, logE out --------------------------------
    h2d = (n * d + 2 * n + (n_ref * d if cross else 0)) * 8
    d2h = (kmax - 1) * 3 * 8 + 160
    if args.no_e2e:
        te, e2e_value = float('nan'), None
    else:
        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        te = float(te.item())
        e2e_value = args.steps * pairs_total / te

    if rank == 0:
        lo, hi = shard_range(n, 0, world)
        pairs_launch = float(hi - lo) * float(n_ref)                    # units one K2 launch processes on rank 0
        t_k2 = float(np.mean(ms_filter)) * 1e-3
        achieved = 3.0 * d * pairs_launch / t_k2 / 1e12                 # algorithmic: 3d FLOP per pair (sub, mul, add)
        exec_fma = 2.0 * d * pairs_launch / t_k2 / 1e12               # executed: d FMA per pair (score form)
        roofline = {'bound': 'fp32_fma', 'achieved': achieved, 'peak': FP32_PEAK_THEORETICAL_TFLOPS, 'unit': 'TFLOP/s',
                    'frac': achieved / FP32_PEAK_THEORETICAL_TFLOPS,
                    'traffic': NCU_DRAM_BYTES_PER_LAUNCH.get((args.workload, world, stats['filter_mode'] == 16)),
                    'kernel': 'knn_filter_kernel<D=%d,TQ=%d,TR=%d,packed=%d,mode=%d>' % (
                        d, stats['filter_tq'], stats['filter_tr'], stats['filter_packed'], stats['filter_mode']),
                    'ms_per_launch': 1e3 * t_k2, 'share_of_step': float(np.sum(ms_filter)) / t_ms if world == 1 else None,
                    'peak_source': 'theoretical FP32 FMA peak 148 SMs x 128 lanes x 2 FLOP x 1.965 GHz = %.1f TFLOP/s '
                                   '(MEASURED_PEAKS.json has no FP32 entry; the FMA microbenchmark of this run, mce_fma_peak, '
                                   'reads ffma %.1f / ffma2 %.1f)' % (FP32_PEAK_THEORETICAL_TFLOPS, peak['ffma'], peak['ffma2']),
                    'algorithmic_tflops': achieved, 'executed_fma_tflops': exec_fma,
                    'fma_pipe_frac': exec_fma / FP32_PEAK_THEORETICAL_TFLOPS,
                    'note': 'achieved = algorithmic 3d FLOP per pair (sub, mul, add; SURVEY 8d M4); the filter executes d FMA '
                            '= 2d FLOP per pair, so frac can exceed 1; fma_pipe_frac = executed FMA rate / theoretical peak'}
        if stats['filter_mode'] == 16:
            # tcgen05 engine: the d FMAs per pair run on the tensor cores as TF32 MMAs of depth K = 8 ceil((d+2)/8)
            kpad = 8 * ((d + 2 + 7) // 8)
            tf32_peak = _measured_tf32_peak()
            mma = 2.0 * kpad * pairs_launch / t_k2 / 1e12
            roofline.update({'bound': 'tensor', 'achieved': mma, 'peak': tf32_peak[0], 'frac': mma / tf32_peak[0],
                             'kernel': 'knn_filter_tc_kernel<KSTEPS=%d>' % (kpad // 8), 'peak_source': tf32_peak[1],
                             'fma_pipe_frac': None, 'executed_fma_tflops': None,
                             'algorithmic_frac_of_fp32_fma_peak': achieved / FP32_PEAK_THEORETICAL_TFLOPS,
                             # the unit that actually binds this kernel is the ALU pipe of the epilogue: 17 FMNMX3 per 32
                             # scores; its peak is measured live with the kernel's own instruction (mce_unit_peak 0)
                             'epilogue_alu': _alu_floor(ctx, pairs_launch, t_k2),
                             'note': 'achieved = executed TF32 MMA FLOP (2 K per pair, K = %d for ndim %d) against half the '
                                     'measured bf16 peak; algorithmic_tflops = 3d FLOP per pair (SURVEY 8d M4), beside it as '
                                     'a fraction of the theoretical FP32 FMA peak the CUDA-core filter is bounded by' % (kpad, d)})
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': t_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                'dtype': 'f32 filter, f64 re-rank/reduce', 'data': 'synthetic',
                'config': config_for(args.workload, world),
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'upload': ('one H2D copy of the pinned arrays' if world == 1 else
                                   'each rank copies its own 1/%d of the rows from its pinned host arrays, NCCL all-gather' % world),
                        's_per_logE': None if args.no_e2e else te / args.steps},
                'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline,
                's_per_logE': t_ms * 1e-3 / args.steps, 'logE': [float(x) for x in loge],
                'logE_minus_analytic': [float(x - m['lnZ']) for x in loge], 'rows_rescanned_per_step': rescanned / args.steps,
                'stage_ms': {k: stats[k] for k in ('ms_cov', 'ms_whiten', 'ms_filter', 'ms_rerank', 'ms_exact', 'ms_reduce')},
                'filter': {k: stats[k] for k in ('filter_tq', 'filter_tr', 'filter_packed', 'filter_mode', 'ref_chunks', 'list_len',
                                                 'filter_grid', 'filter_smem', 'filter_ctas_per_sm')}}
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(wl, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
