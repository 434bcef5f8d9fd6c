#!/bin/bash
# last visit: all GPU tests, the bench line, ncu of the marg kernel (the only kernel changed since scripts/gpu_round2_final.sh r03c)
TAG=r03e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 300 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
N="ncu --set full --metrics smsp__inst_executed_pipe_xu.sum --clock-control none --import-source on -f"
timeout 900 $N -k regex:numerator_marg -s 3 -c 1 -o gpurun_out/marg_c2_$TAG python bench.py --sub none --no-cpu-baseline --config C2 --steps 1 --warmup 3 > gpurun_out/ncu_marg_c2_$TAG.log 2>&1
