#!/bin/bash
# zgrid_terms geometry, 'full' pair loop with registers + balanced items: parity + timings + ncu of the C4 KDE kernel
TAG=r02l
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "full or c4 or C4 or model_matrix" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
for C in C3 C4; do
  echo "== $C"
  timeout 300 python bench.py --config $C --sub none --no-cpu-baseline --steps 5 --warmup 3 2>> gpurun_out/ab_$TAG.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f parity %s' % (d['ms_per_step'], d['parity_check']['max_err_vs_oracle']), d['kernel_ms'])"
done 2>&1 | tee gpurun_out/ab_$TAG.log
N="ncu --set full --metrics smsp__inst_executed_pipe_xu.sum --clock-control none --import-source on -f"
timeout 600 $N -k regex:numerator_f32 -s 5 -c 2 -o gpurun_out/full_c4_$TAG python bench.py --config C4 --sub none --no-cpu-baseline --steps 1 --warmup 2 > gpurun_out/ncu_full_c4_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
