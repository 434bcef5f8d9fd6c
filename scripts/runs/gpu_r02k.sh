#!/bin/bash
# A/B: bucket-row z lookup, tail prefetch; parity tests; C4 with the packed 3-D pair loop; selection tiles
TAG=r02k
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$TAG.log
bash scripts/ab_fused.sh $TAG
for C in C1 C4; do
  for L in chimera_b200/libchimera_b200.so; do
    echo "== $C $L"
    CHB_LIB=$PWD/$L timeout 300 python bench.py --config $C --sub none --no-cpu-baseline --steps 5 --warmup 3 2>> gpurun_out/ab_$TAG.err \
      | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f sel %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['kernel_ms']['selection_ms'], d['parity_check']['max_err_vs_oracle']))"
  done
done 2>&1 | tee -a gpurun_out/ab_$TAG.log
timeout 300 python bench.py --sub none --no-cpu-baseline --steps 10 --warmup 3 2>> gpurun_out/ab_$TAG.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3 full: ms/step %.3f' % d['ms_per_step'], d['kernel_ms'])" | tee -a gpurun_out/ab_$TAG.log
