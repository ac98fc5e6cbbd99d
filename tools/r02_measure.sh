#!/bin/bash
# Round-2 measurement pass on one B200: bench lines for every single-GPU workload, the ncu launch list of the bench
# command, and one full ncu capture of the dominant kernel (outputs under gpurun_out/, copied to profiles/ by hand).
set -x
mkdir -p gpurun_out
T=${1:-r02}
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
python bench.py --workload cfg5s --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg5s.json 2> gpurun_out/${T}_bench_cfg5s.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --allow-short-warmup > gpurun_out/${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_filter_tc -c 1 -f -o gpurun_out/${T}_tc_cfg2 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --allow-short-warmup > gpurun_out/${T}_ncu_full.log 2>&1
ncu -i gpurun_out/${T}_tc_cfg2.ncu-rep --page raw --csv > gpurun_out/${T}_tc_cfg2_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_tc_cfg2.ncu-rep --page source --csv > gpurun_out/${T}_tc_cfg2_source.csv 2>/dev/null
tail -3 gpurun_out/${T}_bench_cfg2.json gpurun_out/${T}_bench_cfg3.json gpurun_out/${T}_bench_cfg5s.json
