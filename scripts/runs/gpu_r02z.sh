#!/bin/bash
TAG=r02z
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "full or c4 or C4 or golden or model_matrix" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
for L in chimera_b200/libchimera_b200.so chimera_b200/ab/full256.so; do
echo "== C4 $L"
CHB_LIB=$PWD/$L timeout 400 python bench.py --config C4 --sub none --no-cpu-baseline --steps 3 --warmup 2 2>> gpurun_out/ab_$TAG.err \
  | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))" | tee -a gpurun_out/ab_$TAG.log
done
