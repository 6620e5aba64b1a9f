#!/bin/bash
# Map-encoder profiling visit: pipeline trace at 2048 crops + one ncu --set full capture of each encoder kernel.
mkdir -p gpurun_out
N=2048 REPS=3 timeout 300 python scripts/run_mapenc.py > gpurun_out/mapenc_trace.txt 2>&1
cat gpurun_out/mapenc_trace.txt
N=2048 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_conv|crop_pack|tc_gemm" -s 8 -c 8 -f -o gpurun_out/prof_enc \
   python scripts/run_mapenc.py > gpurun_out/prof_enc.log 2>&1
tail -3 gpurun_out/prof_enc.log
ls -la gpurun_out/*.ncu-rep
