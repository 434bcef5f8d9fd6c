#!/bin/bash
# compute-sanitizer memcheck over the small parity tests of every kernel variant (golden cases, KDE options, windows,
# 64/128/256-thread instantiations, 3-D windows, selection)
TAG=r03h
mkdir -p gpurun_out
timeout 560 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 \
  python -m pytest tests -m gpu -q -x -k "golden_fp32 or kde_options or epanechnikov_unbinned or selection_bpl or fast_path_variants" 2>&1 | tail -25 | tee gpurun_out/memcheck_$TAG.log
echo "rc=${PIPESTATUS[0]}" | tee -a gpurun_out/memcheck_$TAG.log
