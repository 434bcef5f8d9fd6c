#!/bin/bash
# racecheck + memcheck over the fp64-mode kernels, the table / model kernels, the selection kernels and the set-up kernels
TAG=r03l
mkdir -p gpurun_out
timeout 330 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 \
  python -m pytest tests -m gpu -q -k "golden_fp64 or cosmology_functions or mass_functions or selection_bpl or neff_gate or model_matrix_fp32_split" 2>&1 | grep -v "^=========$" | tail -40 | tee gpurun_out/racecheck_fp64_$TAG.log
echo "racecheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/racecheck_fp64_$TAG.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 \
  python -m pytest tests/test_gpu_setup.py tests/test_gpu_parity.py -m gpu -q -k "setup or pixel or healpix or p_cat or golden_fp64 or z_grids" 2>&1 | tail -8 | tee gpurun_out/memcheck_setup_$TAG.log
echo "memcheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/memcheck_setup_$TAG.log
