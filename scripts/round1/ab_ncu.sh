#!/bin/bash
# kernel-time A/B of library builds on the C3 workload (ncu launch list, numerator kernels only)
for L in "$@"; do
  echo "== $L"
  CHB_LIB=$PWD/$L ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:numerator_f32 -s 2 -c 2 --csv \
    --log-file gpurun_out/ab.csv python scripts/sweep.py CHB_SPLIT 1 > gpurun_out/ab.log 2>&1
  tail -1 gpurun_out/ab.log
  grep -E "gpu__time_duration|inst_executed" gpurun_out/ab.csv | cut -d, -f5,15-
done
