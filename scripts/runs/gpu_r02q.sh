#!/bin/bash
TAG=r02q
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "paths or windowed or model_matrix" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
echo "== C5 at N=1"
timeout 600 python bench.py --config C5 --sub none --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/c5_n1_$TAG.json 2> gpurun_out/c5_n1_$TAG.err
python -c "import sys,json; d=json.loads(open('gpurun_out/c5_n1_$TAG.json').read().strip().splitlines()[-1]); print('C5 ms/step %.3f value %.4g parity %s launches %s' % (d['ms_per_step'], d['value'], d.get('parity_check',{}).get('max_err_vs_oracle'), d['gpu_launches']), d['kernel_ms'])" | tee -a gpurun_out/ab_$TAG.log; tail -3 gpurun_out/c5_n1_$TAG.err
N="ncu --set full --metrics smsp__inst_executed_pipe_xu.sum --clock-control none --import-source on -f"
timeout 900 $N -k regex:numerator_fused -s 1 -c 1 -o gpurun_out/fused_c5_$TAG python bench.py --config C5 --nev 1000 --sub none --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_fused_c5_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
