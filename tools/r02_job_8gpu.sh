#!/bin/bash
# 8-GPU check of the final library: NCCL check of both upload paths and the cfg2 bench line at N=8.
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py 2>&1 | grep -v "^\[W\|^W0\|^\*\*\*" | tail -6 | tee gpurun_out/r02m_dist_check_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench_cfg2_8gpu.json 2> gpurun_out/r02m_bench_cfg2_8gpu.err
tail -c 700 gpurun_out/r02m_bench_cfg2_8gpu.json
