#!/bin/bash
# Round-2 final GPU visit: all GPU tests, the bench line (all sub-records), the reference arm, the ncu launch list of the
# bench command and `ncu --set full` captures of the dominant kernel of each configuration.
#   usage (under gpurun): bash scripts/gpu_round2_final.sh <tag> [skip_tests]
TAG=${1:-r02z}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_$TAG.log
fi
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 400 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -c 300 gpurun_out/bench_ref_$TAG.json
B="python bench.py --sub none --no-cpu-baseline"
# (the workload's own set-up launches ~400 HEALPix kernels first: the list is restricted to the kernels of a step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'build_tables|zgrid_terms|numerator|selection|reduce_kernel|catalog_collapse' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  $B --steps 2 --warmup 1 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
N="ncu --set full --metrics smsp__inst_executed_pipe_xu.sum --clock-control none --import-source on -f"
timeout 900 $N -k regex:numerator_fused -s 3 -c 1 -o gpurun_out/fused_c3_$TAG $B --steps 1 --warmup 3 > gpurun_out/ncu_fused_c3_$TAG.log 2>&1
timeout 900 $N -k regex:selection_f32 -s 3 -c 1 -o gpurun_out/sel_c3_$TAG $B --steps 1 --warmup 3 > gpurun_out/ncu_sel_c3_$TAG.log 2>&1
timeout 900 $N -k regex:numerator_fused -s 3 -c 1 -o gpurun_out/fused_c3_refdefault_$TAG $B --kde epan-binned --steps 1 --warmup 3 > gpurun_out/ncu_fused_c3_refdefault_$TAG.log 2>&1
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; tail -3 gpurun_out/smoke_$TAG.log
timeout 900 $N -k regex:numerator_marg -s 3 -c 1 -o gpurun_out/marg_c2_$TAG $B --config C2 --steps 1 --warmup 3 > gpurun_out/ncu_marg_c2_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
