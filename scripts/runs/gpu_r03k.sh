#!/bin/bash
# after the racecheck fix: all GPU tests + the bench line
TAG=r03k
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 200 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
