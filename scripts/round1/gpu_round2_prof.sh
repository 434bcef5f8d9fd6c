#!/bin/bash
# Round-2 profiling visit: ncu launch list of the bench command + ncu --set full of the dominant kernel of each config.
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --sub none --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  $B --steps 2 --warmup 1 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
N="ncu --set full --metrics smsp__inst_executed_pipe_xu.sum --clock-control none --import-source on -f"
timeout 900 $N -k regex:numerator_fused -s 3 -c 1 -o gpurun_out/fused_c3_$TAG $B --steps 1 --warmup 3 > gpurun_out/ncu_fused_c3_$TAG.log 2>&1
timeout 900 $N -k regex:selection_f32 -s 3 -c 1 -o gpurun_out/sel_c3_$TAG $B --steps 1 --warmup 3 > gpurun_out/ncu_sel_c3_$TAG.log 2>&1
timeout 900 $N -k regex:numerator_fused -s 3 -c 1 -o gpurun_out/fused_c1_$TAG $B --config C1 --steps 1 --warmup 3 > gpurun_out/ncu_fused_c1_$TAG.log 2>&1
timeout 900 $N -k regex:numerator_f32 -s 6 -c 2 -o gpurun_out/marg_c2_$TAG $B --config C2 --steps 1 --warmup 3 > gpurun_out/ncu_marg_c2_$TAG.log 2>&1
timeout 900 $N -k regex:numerator_kernel -s 2 -c 1 -o gpurun_out/fp64_c3_$TAG $B --fp-mode fp64 --ninj 100000 --steps 1 --warmup 2 > gpurun_out/ncu_fp64_c3_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
