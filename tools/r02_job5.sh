#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02_gpu_suite_c.log 2>&1; tail -8 gpurun_out/r02_gpu_suite_c.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -6 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_cfg2.json 2> gpurun_out/r02b_bench_cfg2.err; tail -c 600 gpurun_out/r02b_bench_cfg2.json
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 3 > gpurun_out/r02b_bench_cfg1.json 2> gpurun_out/r02b_bench_cfg1.err; tail -c 1500 gpurun_out/r02b_bench_cfg1.json
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/r02b_bench_cfg4.json 2> gpurun_out/r02b_bench_cfg4.err; tail -c 2500 gpurun_out/r02b_bench_cfg4.json; tail -3 gpurun_out/r02b_bench_cfg4.err
timeout 600 python bench.py --workload cfg4 --impl reference --steps 1 --warmup 0 > gpurun_out/r02b_bench_cfg4_reference.json 2> gpurun_out/r02b_bench_cfg4_reference.err; tail -c 800 gpurun_out/r02b_bench_cfg4_reference.json
for S in 1 2 3; do
  timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py --sync $S > gpurun_out/r02_racecheck_ffma_sync$S.log 2>&1; tail -3 gpurun_out/r02_racecheck_ffma_sync$S.log
done
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py --engine 1 > gpurun_out/r02_racecheck_tc.log 2>&1; tail -3 gpurun_out/r02_racecheck_tc.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py --engine 1 > gpurun_out/r02_memcheck_tc.log 2>&1; tail -3 gpurun_out/r02_memcheck_tc.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/r02_memcheck_ffma.log 2>&1; tail -3 gpurun_out/r02_memcheck_ffma.log
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_case.py --engine 1 > gpurun_out/r02_synccheck_tc.log 2>&1; tail -3 gpurun_out/r02_synccheck_tc.log
