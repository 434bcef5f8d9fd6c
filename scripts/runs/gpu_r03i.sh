#!/bin/bash
# compute-sanitizer: memcheck over the model matrix / 'full' / 'marginalized' tests, racecheck (shared-memory hazards: the
# kernels alias several shared arrays across phases) over the small golden cases of every kind
TAG=r03i
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 \
  python -m pytest tests -m gpu -q -x -k "model_matrix or full_3d_windows or marginalized_binned" 2>&1 | tail -8 | tee gpurun_out/memcheck2_$TAG.log
echo "memcheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/memcheck2_$TAG.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 8 \
  python -m pytest tests -m gpu -q -x -k "golden_fp32 or epanechnikov_unbinned" 2>&1 | tail -30 | tee gpurun_out/racecheck_$TAG.log
echo "racecheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/racecheck_$TAG.log
