#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list of the bench command, full captures of the hot kernels.
# usage (under gpurun): bash scripts/gpu_round.sh <tag> [skip_tests]
TAG=${1:-r01}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$TAG.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3500 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -c 1200 gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
# reweighting (MODE 1) and KDE/z-integral (MODE 2) kernels of one evaluation at the bench size: traffic + pipe utilisation
timeout 900 ncu --set full --clock-control none --import-source on -k regex:numerator_f32 -s 4 -c 2 -f -o gpurun_out/numf32_bench_$TAG \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_numf32_bench_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:selection_f32 -s 1 -c 1 -f -o gpurun_out/self32_$TAG \
  python scripts/profile_run.py 148 4 fp32 > gpurun_out/ncu_self32_$TAG.log 2>&1
timeout 300 python scripts/profile_run.py 592 4 fp32 2>&1 | tail -3 | tee gpurun_out/phase_$TAG.log
