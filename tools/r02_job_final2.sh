#!/bin/bash
# cfg5 on one GPU, smoke(), sanitizer passes over the tcgen05 engine including a long unit (bootstrap + slabs).
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02k_smoke.log 2>&1; tail -4 gpurun_out/r02k_smoke.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py --engine 1 --long > gpurun_out/r02k_memcheck_tc.log 2>&1; tail -3 gpurun_out/r02k_memcheck_tc.log
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_case.py --engine 1 --long > gpurun_out/r02k_synccheck_tc.log 2>&1; tail -3 gpurun_out/r02k_synccheck_tc.log
timeout 900 python bench.py --workload cfg5 --steps 2 --warmup 1 --allow-short-warmup --no-cpu-baseline --no-e2e > gpurun_out/r02k_bench_cfg5_1gpu.json 2> gpurun_out/r02k_bench_cfg5_1gpu.err
tail -c 900 gpurun_out/r02k_bench_cfg5_1gpu.json
