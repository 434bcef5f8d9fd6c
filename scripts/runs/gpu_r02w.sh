#!/bin/bash
TAG=r02w
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "paths or model_matrix or golden or kde_options" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
for L in chimera_b200/libchimera_b200.so chimera_b200/ab/base_w.so; do
echo "== C3 refdefault $L"
CHB_LIB=$PWD/$L timeout 300 python bench.py --kde epan-binned --sub none --no-cpu-baseline --steps 10 --warmup 3 --ninj 100000 2>> gpurun_out/ab_$TAG.err \
  | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))" | tee -a gpurun_out/ab_$TAG.log
echo "== C1 $L"
CHB_LIB=$PWD/$L timeout 300 python bench.py --config C1 --sub none --no-cpu-baseline --steps 10 --warmup 3 2>> gpurun_out/ab_$TAG.err \
  | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  numerator_kernels %.3f parity %s' % (d['ms_per_step'], d['kernel_ms']['numerator_kernels_ms'], d['parity_check']['max_err_vs_oracle']))" | tee -a gpurun_out/ab_$TAG.log
done
B="python bench.py --sub none --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'build_tables|zgrid_terms|numerator|selection|reduce_kernel|catalog_collapse' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  $B --steps 2 --warmup 1 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
tail -3 gpurun_out/launches_$TAG.csv | cut -c1-300
