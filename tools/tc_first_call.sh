#!/bin/bash
# First hardware session of the tcgen05 filter engine, as one gpurun call:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/tc_first_call.sh'
# Every step runs under its own timeout; later steps are skipped when bring-up stage 1 fails.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
set -x
timeout 900 python tools/tc_bringup.py > gpurun_out/tc_bringup.log 2>&1
tail -n 40 gpurun_out/tc_bringup.log
if ! python - <<'PY'
import json, sys
recs = json.load(open('gpurun_out/tc_bringup.json'))
ok = any(r.get('stage') == '1' and r.get('idx_equal') and r.get('dist_equal') for r in recs)
sys.exit(0 if ok else 1)
PY
then
    echo "bring-up stage 1 failed: see gpurun_out/tc_bringup.json and gpurun_out/tc_dump_stage*.npz"
    exit 0
fi
MCE_TEST_NEW_GOLDEN=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k golden_reference_outputs 2>&1 | tail -n 5 | tee gpurun_out/new_golden.log
MCE_TEST_TC=1 timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -n 15 | tee gpurun_out/tc_tests.log
# the whole GPU suite on the tensor-core engine
MCE_FILTER_ENGINE=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 | tee gpurun_out/tc_suite.log
timeout 600 python bench.py --filter-engine tc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
tail -c 3000 gpurun_out/bench_tc.json
# one wave of units (148 CTAs x 256 queries) against 1e6 references: launch list and one full capture
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tc_launches.csv \
    python tools/ncu_variants.py --n 1000000 --d 10 --qn 37888 --cfg tc > gpurun_out/tc_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_filter_tc -c 1 -f -o gpurun_out/tc_full \
    python tools/ncu_variants.py --n 1000000 --d 10 --qn 37888 --cfg tc > gpurun_out/tc_full.log 2>&1
ls -la gpurun_out | tail -n 20
