#!/bin/bash
# A/B of build variants of the fused kernel on the C3 workload (100k injections: the numerator is what is compared).
# usage (under gpurun): bash scripts/ab_fused.sh <tag>    -- the variant libraries are built beforehand by scripts/ab_build.sh
TAG=${1:-ab}
mkdir -p gpurun_out
for L in chimera_b200/libchimera_b200.so chimera_b200/ab/*.so; do
  [ -f "$L" ] || continue
  echo "== $L"
  CHB_LIB=$PWD/$L timeout 300 python bench.py --sub none --no-cpu-baseline --steps 5 --warmup 3 --ninj 100000 2> gpurun_out/ab_$TAG.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f  zgrid %.3f  sel %.3f  parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['kernel_ms']['zgrid_terms_ms'], d['kernel_ms']['selection_ms'], d['parity_check']['max_err_vs_oracle']))"
done 2>&1 | tee gpurun_out/ab_$TAG.log
