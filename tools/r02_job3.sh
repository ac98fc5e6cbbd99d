#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_var_pin.py -x -q -m gpu > gpurun_out/r02_job3_tests.log 2>&1
tail -15 gpurun_out/r02_job3_tests.log
for L in 1 4 8 16; do
  timeout 300 python tools/grid_bench.py --nstreams $L --cpu-cells 0 > gpurun_out/r02_grid_all_lanes$L.json 2> gpurun_out/r02_grid_all_lanes$L.err
  cat gpurun_out/r02_grid_all_lanes$L.json
done
timeout 300 python tools/grid_bench.py --kind SingleChains --nstreams 8 --cpu-cells 0 > gpurun_out/r02_grid_single_lanes8.json 2> gpurun_out/r02_grid_single.err
cat gpurun_out/r02_grid_single_lanes8.json
timeout 600 python tools/sweep_bench.py --out gpurun_out/r02_sweep_1e6_d10.json 2> gpurun_out/r02_sweep.err | cut -c1-1500
timeout 600 python tools/sweep_bench.py --n 200000 --d 6 --scale logpower --nbatch 8 --out gpurun_out/r02_sweep_2e5_d6_log.json 2>> gpurun_out/r02_sweep.err | cut -c1-1500
