#!/bin/bash
TAG=r03a
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "selection or golden or model_matrix or rate or parity" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
echo "== C3"
timeout 300 python bench.py --sub none --no-cpu-baseline --steps 10 --warmup 3 2>> gpurun_out/ab_$TAG.err \
  | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f parity %s cold %.2f' % (d['ms_per_step'], d['parity_check']['max_err_vs_oracle'], d['e2e_cold']['seconds']), d['kernel_ms'])" | tee -a gpurun_out/ab_$TAG.log
