#!/bin/bash
# A/B of the cp.async sample staging + parity tests of the paths that use fu_reweight + fresh ncu capture of the fused kernel
TAG=r02j
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "paths or model_matrix or golden or c3" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$TAG.log
bash scripts/ab_fused.sh $TAG
for C in C1 C2; do
  for L in chimera_b200/libchimera_b200.so chimera_b200/ab/cpasync0.so; do
    echo "== $C $L"
    CHB_LIB=$PWD/$L timeout 300 python bench.py --config $C --sub none --no-cpu-baseline --steps 10 --warmup 3 2>> gpurun_out/ab_$TAG.err \
      | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))"
  done
done 2>&1 | tee -a gpurun_out/ab_$TAG.log
B="python bench.py --sub none --no-cpu-baseline"
N="ncu --set full --metrics smsp__inst_executed_pipe_xu.sum --clock-control none --import-source on -f"
timeout 900 $N -k regex:numerator_fused -s 3 -c 1 -o gpurun_out/fused_c3_$TAG $B --steps 1 --warmup 3 > gpurun_out/ncu_fused_c3_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
