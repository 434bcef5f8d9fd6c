#!/bin/bash
TAG=r02o
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "epan or kde_options or golden" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
echo "== C3 --kde epan"
timeout 300 python bench.py --kde epan --sub none --no-cpu-baseline --steps 5 --warmup 3 --ninj 100000 2>> gpurun_out/ab_$TAG.err \
  | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))" | tee gpurun_out/ab_$TAG.log
echo "== C5 at N=1"
timeout 600 python bench.py --config C5 --sub none --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/c5_n1_$TAG.json 2> gpurun_out/c5_n1_$TAG.err
python -c "import sys,json; d=json.loads(open('gpurun_out/c5_n1_$TAG.json').read().strip().splitlines()[-1]); print('C5 ms/step %.3f value %.4g parity %s' % (d['ms_per_step'], d['value'], d.get('parity_check')), d['kernel_ms'])" | tee -a gpurun_out/ab_$TAG.log; tail -3 gpurun_out/c5_n1_$TAG.err
B="python bench.py --sub none --no-cpu-baseline"
N="ncu --set full --metrics smsp__inst_executed_pipe_xu.sum --clock-control none --import-source on -f"
timeout 900 $N -k regex:numerator_marg -s 3 -c 1 -o gpurun_out/marg_c2_$TAG $B --config C2 --steps 1 --warmup 3 > gpurun_out/ncu_marg_c2_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
