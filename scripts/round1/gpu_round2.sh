#!/bin/bash
# Round-2 GPU visit: parity tests, bench line (fused kernel) + A/B against the round-1 split kernels, ncu launch list,
# full capture of the fused kernel.   usage (under gpurun): bash scripts/gpu_round2.sh <tag> [skip_tests]
TAG=${1:-r02}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_$TAG.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --options fused=0 > gpurun_out/bench_split_$TAG.json 2> gpurun_out/bench_split_$TAG.err
tail -c 1500 gpurun_out/bench_split_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:numerator_fused -s 3 -c 1 -f -o gpurun_out/fused_$TAG \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused_$TAG.log 2>&1
ls -la gpurun_out | tail -12
