#!/bin/bash
TAG=r02u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "paths or windowed or model_matrix or golden" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
for L in chimera_b200/libchimera_b200.so chimera_b200/ab/base_u.so; do
echo "== C5 (2000 events) $L"
CHB_LIB=$PWD/$L timeout 600 python bench.py --config C5 --nev 2000 --ninj 100000 --sub none --no-cpu-baseline --steps 3 --warmup 2 2>> gpurun_out/ab_$TAG.err \
  | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))" | tee -a gpurun_out/ab_$TAG.log
echo "== C3 $L"
CHB_LIB=$PWD/$L timeout 300 python bench.py --sub none --no-cpu-baseline --steps 5 --warmup 3 --ninj 100000 2>> gpurun_out/ab_$TAG.err \
  | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))" | tee -a gpurun_out/ab_$TAG.log
done
