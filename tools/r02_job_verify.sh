#!/bin/bash
# Final verification of the committed library on one B200: GPU suite, smoke(), default bench line, small-chain workloads.
T=${1:-r02m}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_gpu_suite.log
python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg4.json 2>/dev/null
python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg1.json 2>/dev/null
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_reference_default.json 2> gpurun_out/${T}_bench_reference_default.err
for f in default cfg4 cfg1 reference_default; do python -c "
import json,sys
d=json.loads(open('gpurun_out/${T}_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d.get('ms_per_step'), d.get('value'), d.get('e2e'), (d.get('roofline') or {}).get('frac'), d.get('clocks'))"; done
