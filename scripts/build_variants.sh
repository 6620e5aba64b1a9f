#!/bin/bash
# Development helper: builds variants of the library with different -D flags for rollout.cu into build/variants/lib_<name>.so
# usage: scripts/build_variants.sh name1 "-DNODE_R=2 -DNODE_WARPS=8" name2 "..." ...   (then STRIVE_LIB=build/variants/lib_<name>.so python scripts/prof_step.py)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  ( nvcc $FLAGS $defs -c strive_b200/csrc/rollout.cu -o build/variants/rollout_$name.o &&
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/lib_$name.so build/api.o build/mapenc.o build/variants/rollout_$name.o build/loss.o build/tc_selftest.o build/mapenc_tc.o build/metrics.o &&
    echo built $name ) &
done
wait
