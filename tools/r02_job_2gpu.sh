#!/bin/bash
# 2-GPU check of the final library: the distributed tests and the cfg2 bench line at N=2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02k_gpu_dist_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_cfg2_2gpu.json 2> gpurun_out/r02k_bench_cfg2_2gpu.err
tail -c 1200 gpurun_out/r02k_bench_cfg2_2gpu.json
