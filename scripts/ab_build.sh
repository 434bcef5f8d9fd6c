#!/bin/bash
# Build variant libraries of the fused kernel into chimera_b200/ab/ (they travel to the GPU box with the snapshot).
# usage: bash scripts/ab_build.sh name1 "defs1" name2 "defs2" ...
mkdir -p chimera_b200/ab
while [ $# -ge 2 ]; do
  N=$1; D=$2; shift 2
  CHB_BUILD_OUT=$PWD/chimera_b200/ab/$N.so CHB_BUILD_DEFS="$D" python -m chimera_b200.build --force > /dev/null 2>&1 && echo "built $N ($D)" || echo "FAILED $N"
done
