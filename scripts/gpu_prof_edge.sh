#!/bin/bash
# Edge-kernel profiling visit: ncu --set full of the tcgen05 edge forward kernel (and the mma.sync backward) on the bench batch.
mkdir -p gpurun_out
FT=3 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-edge_fwd_tc}" -s ${SKIP:-3} -c ${COUNT:-1} -f -o gpurun_out/prof_edge_tc \
   python scripts/run_decode.py > gpurun_out/prof_edge_tc.log 2>&1
tail -3 gpurun_out/prof_edge_tc.log
ncu -i gpurun_out/prof_edge_tc.ncu-rep --page raw --csv > gpurun_out/edge_tc_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_edge_tc.ncu-rep --page source --csv > gpurun_out/edge_tc_src.csv 2>/dev/null
ls -la gpurun_out/prof_edge_tc.ncu-rep gpurun_out/edge_tc_raw.csv gpurun_out/edge_tc_src.csv
