#!/bin/bash
TAG=r03d
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
for L in chimera_b200/libchimera_b200.so chimera_b200/ab/base_d.so; do
for C in C3 C2 C1; do
echo "== $C $L"
CHB_LIB=$PWD/$L timeout 300 python bench.py --config $C --sub none --no-cpu-baseline --steps 10 --warmup 3 2>> gpurun_out/ab_$TAG.err \
  | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))" | tee -a gpurun_out/ab_$TAG.log
done
done
