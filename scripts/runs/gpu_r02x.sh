#!/bin/bash
TAG=r02x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "full or c4 or C4 or golden or model_matrix" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
for O in "" "kde_win=0"; do
echo "== C4 options [$O]"
timeout 400 python bench.py --config C4 --sub none --no-cpu-baseline --steps 3 --warmup 2 --options "$O" 2>> gpurun_out/ab_$TAG.err \
  | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']))" | tee -a gpurun_out/ab_$TAG.log
done
