"""BASELINE config 4 shape: every (model, data) cell of the Planck fullgrid_R2 analysis as an independent
auto-mode evidence problem (kmax=2), batched chain-per-GPU.  Chains are synthetic Gaussian chains with the
(ndim, N) of the reference's result tables (mcevidence_b200/planck_grid_shapes.json); the reference needs
"about ~30 mins" on one MPI process for the real grid, dominated by text parsing (planck_mcevidence.py:21).

    python tools/grid_bench.py [--kind AllChains|SingleChains] [--cpu-cells 8]
    torchrun --nproc-per-node N tools/grid_bench.py        # chains dealt largest-first over N GPUs
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcevidence_b200.grid import evidence_many         # noqa: E402
from mcevidence_b200.synthetic import gaussian_chain   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--kind', default='AllChains')
ap.add_argument('--cpu-cells', type=int, default=6)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--nstreams', type=int, default=8)
a = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
shapes = json.load(open(os.path.join(ROOT, 'mcevidence_b200', 'planck_grid_shapes.json')))[a.kind]
cells = [(nd, n) for (_, _, nd, ns) in shapes for n in ns]
t0 = time.perf_counter()
chains = [gaussian_chain(n, nd, seed=i, inflate=1.0, integer_weights=True, nextra=3)[0] for i, (nd, n) in enumerate(cells)]
t_gen = time.perf_counter() - t0
ndims = [nd for nd, _ in cells]
evidence_many(chains[:32], kmax=2, ndim=ndims[:32], device=local, nstreams=a.nstreams)   # warm-up (module load, workspaces)
times = []
for _ in range(a.reps):
    t0 = time.perf_counter()
    loge, var, ms = evidence_many(chains, kmax=2, ndim=ndims, device=local, nstreams=a.nstreams)
    times.append(time.perf_counter() - t0)
out = {'kind': a.kind, 'n_gpus': world, 'nstreams': a.nstreams, 'cells': len(cells), 'rows': int(sum(n for _, n in cells)),
       'pairs': float(sum(float(n) ** 2 for _, n in cells)), 'wall_s_best': min(times), 'wall_s_all': times,
       'device_ms_sum': float(np.sum(ms)), 'synthetic_generation_s': t_gen,
       'cells_per_s': len(cells) / min(times)}
if rank == 0 and a.cpu_cells > 0:
    from oracle import mce_oracle as O
    order = np.argsort([-n for _, n in cells])
    pick = [int(order[i]) for i in np.linspace(0, len(cells) - 1, a.cpu_cells).astype(int)]
    t_cpu, err = 0.0, 0.0
    for i in pick:
        c = chains[i]
        t0 = time.perf_counter()
        want = O.evidence(c[:, 2:], -c[:, 1], c[:, 0], ndims[i], 2, knn='sklearn', scalar_loop=True)
        t_cpu += time.perf_counter() - t0
        err = max(err, float(np.abs(want - loge[i]).max()))
    out.update({'cpu_port_cells': len(pick), 'cpu_port_s': t_cpu, 'cpu_port_rows': int(sum(cells[i][1] for i in pick)),
                'max_abs_logE_diff_vs_cpu_port': err, 'cpu_cores': os.cpu_count()})
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
