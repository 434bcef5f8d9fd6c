#!/bin/bash
TAG=r02y
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
for O in "" "kde_win_t2=30"; do
echo "== C4 options [$O]"
timeout 400 python bench.py --config C4 --sub none --no-cpu-baseline --steps 3 --warmup 2 --options "$O" 2>> gpurun_out/ab_$TAG.err \
  | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))" | tee -a gpurun_out/ab_$TAG.log
done
N="ncu --set full --metrics smsp__inst_executed_pipe_xu.sum --clock-control none --import-source on -f"
timeout 600 $N -k regex:numerator_f32 -s 5 -c 2 -o gpurun_out/full_c4_$TAG python bench.py --config C4 --sub none --no-cpu-baseline --steps 1 --warmup 2 > gpurun_out/ncu_full_c4_$TAG.log 2>&1
