#!/bin/bash
# ncu --set full of the tcgen05 filter at the cfg3 (K = 32) and cfg5s (K = 24, kmax = 8) shapes, raw pages as csv.
mkdir -p gpurun_out
for W in cfg3 cfg5s; do
  ncu --set full --clock-control none --import-source on -k regex:knn_filter_tc -c 1 -f -o gpurun_out/r02m_tc_$W \
      python bench.py --workload $W --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --allow-short-warmup > gpurun_out/r02m_ncu_$W.log 2>&1
  ncu -i gpurun_out/r02m_tc_$W.ncu-rep --page raw --csv > gpurun_out/r02m_tc_${W}_raw.csv 2>/dev/null
  rm -f gpurun_out/r02m_tc_$W.ncu-rep
  tail -2 gpurun_out/r02m_ncu_$W.log | cut -c1-300
done
