#!/bin/bash
TAG=r02s
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fast_path_variants" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
for O in "fused_nt=64" "fused_nt=128"; do
echo "== C5 at N=1 options [$O]"
timeout 600 python bench.py --config C5 --sub none --no-cpu-baseline --steps 2 --warmup 1 --options "$O" > gpurun_out/c5_n1_$TAG.json 2> gpurun_out/c5_n1_$TAG.err
python -c "import sys,json; d=json.loads(open('gpurun_out/c5_n1_$TAG.json').read().strip().splitlines()[-1]); print('C5 ms/step %.3f value %.4g parity %s launches %s' % (d['ms_per_step'], d['value'], d.get('parity_check',{}).get('max_err_vs_oracle'), d['gpu_launches']), d['kernel_ms'])" | tee -a gpurun_out/ab_$TAG.log; tail -3 gpurun_out/c5_n1_$TAG.err
done
