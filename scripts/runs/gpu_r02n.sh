#!/bin/bash
# A/B: double-buffered constants + deferred unit end + 2-in-1 tail rounds (main) vs HEAD~ (base_l); block-moment epan KDE
TAG=r02n
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "paths or model_matrix or golden or c3 or windowed" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$TAG.log
bash scripts/ab_fused.sh $TAG
for KD in epan-binned epan; do
  for L in chimera_b200/libchimera_b200.so chimera_b200/ab/base_l.so; do
    echo "== C3 --kde $KD $L"
    CHB_LIB=$PWD/$L timeout 300 python bench.py --kde $KD --sub none --no-cpu-baseline --steps 5 --warmup 3 --ninj 100000 2>> gpurun_out/ab_$TAG.err \
      | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))"
  done
done 2>&1 | tee -a gpurun_out/ab_$TAG.log
for L in chimera_b200/libchimera_b200.so chimera_b200/ab/base_l.so; do
  echo "== C1 $L"
  CHB_LIB=$PWD/$L timeout 300 python bench.py --config C1 --sub none --no-cpu-baseline --steps 10 --warmup 3 2>> gpurun_out/ab_$TAG.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))"
done 2>&1 | tee -a gpurun_out/ab_$TAG.log
