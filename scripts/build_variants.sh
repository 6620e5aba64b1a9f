#!/bin/bash
# Development helper: builds variants of the library with different -D flags for rollout.cu (or, with SRC=mapenc_tc, for
# mapenc_tc.cu) into build/variants/lib_<name>.so
# usage: scripts/build_variants.sh name1 "-DNODE_R=2 -DNODE_WARPS=8" name2 "..." ...   (then STRIVE_LIB=build/variants/lib_<name>.so python scripts/prof_step.py)
#        SRC=mapenc_tc scripts/build_variants.sh dbg "-DSTRIVE_TC_DEBUG=1"           (then STRIVE_LIB=... python scripts/mapenc_dbg_sweep.py)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include -I strive_b200/csrc"
SRC=${SRC:-rollout}
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  objs=""
  for o in api mapenc rollout loss tc_selftest mapenc_tc metrics; do
    if [ $o = $SRC ]; then objs="$objs build/variants/${SRC}_$name.o"; else objs="$objs build/$o.o"; fi
  done
  ( nvcc $FLAGS $defs -c strive_b200/csrc/$SRC.cu -o build/variants/${SRC}_$name.o &&
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/lib_$name.so $objs &&
    echo built $name ) &
done
wait
