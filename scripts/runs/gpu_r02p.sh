#!/bin/bash
TAG=r02p
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
for C in C2 C1; do
echo "== $C"
timeout 300 python bench.py --config $C --sub none --no-cpu-baseline --steps 10 --warmup 3 2>> gpurun_out/ab_$TAG.err \
  | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))"
done | tee gpurun_out/ab_$TAG.log
echo "== C5 at N=1"
timeout 600 python bench.py --config C5 --sub none --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/c5_n1_$TAG.json 2> gpurun_out/c5_n1_$TAG.err
python -c "import sys,json; d=json.loads(open('gpurun_out/c5_n1_$TAG.json').read().strip().splitlines()[-1]); print('C5 ms/step %.3f value %.4g parity %s launches %s' % (d['ms_per_step'], d['value'], d.get('parity_check',{}).get('max_err_vs_oracle'), d['gpu_launches']), d['kernel_ms'])" | tee -a gpurun_out/ab_$TAG.log; tail -3 gpurun_out/c5_n1_$TAG.err
