#!/bin/bash
# racecheck after the fix of the rows zeroing, over every kernel family (1-D fused 64/128/256 threads with windows, direct
# sums, bins, Epanechnikov blocks; 'marginalized'; 'full'; split kernels)
TAG=r03j
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 40 \
  python -m pytest tests -m gpu -q -k "golden_fp32 or epanechnikov_unbinned or fast_path_variants or kde_options or marginalized_binned or full_3d_windows" 2>&1 | grep -v "^=========$" | tail -60 | tee gpurun_out/racecheck_$TAG.log
echo "racecheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/racecheck_$TAG.log
