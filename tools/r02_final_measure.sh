#!/bin/bash
# Final measurement pass of round 2 on one B200 (outputs under gpurun_out/, copied to profiles/ by hand):
# GPU suite, default bench run (with cpu_baseline), the other single-GPU workloads, the ncu launch list of the bench
# command and one full ncu capture of the dominant kernel.
T=${1:-r02k}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_gpu_suite.log
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg3.json 2>/dev/null
python bench.py --workload cfg5s --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg5s.json 2>/dev/null
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg4.json 2>/dev/null
python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --allow-short-warmup > gpurun_out/${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_filter_tc -c 1 -f -o gpurun_out/${T}_tc_cfg2 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --allow-short-warmup > gpurun_out/${T}_ncu_full.log 2>&1
ncu -i gpurun_out/${T}_tc_cfg2.ncu-rep --page raw --csv > gpurun_out/${T}_tc_cfg2_raw.csv 2>/dev/null
for f in default cfg3 cfg5s cfg4 cfg1; do tail -c 1500 gpurun_out/${T}_bench_$f.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d.get('e2e'), d['roofline'].get('ms_per_launch'), d['roofline'].get('frac'), d.get('cpu_baseline'))"; done
