#!/bin/bash
# A/B of library builds on the C3 workload: bash scripts/ab_libs.sh lib1.so lib2.so ...
for L in "$@"; do
  echo "== $L"
  CHB_LIB=$PWD/$L python scripts/sweep.py CHB_SPLIT 1 2>&1 | tail -1
done
