#!/bin/bash
# Round-2 profile visit: ncu --set full of the dominant encoder kernel (conv2) and of the tcgen05 edge kernels, raw pages -> gpurun_out/
mkdir -p gpurun_out
N=2048 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_conv_kernel|tc_conv3|tc_conv1" -s 3 -c 4 -f -o gpurun_out/r02_enc \
   python scripts/run_mapenc.py > gpurun_out/r02_enc.log 2>&1
tail -2 gpurun_out/r02_enc.log
ncu -i gpurun_out/r02_enc.ncu-rep --page raw --csv > gpurun_out/r02_enc_raw.csv 2>/dev/null
FT=3 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"edge_fwd_tc|edge_bwd_tc" -s 4 -c 2 -f -o gpurun_out/r02_edge \
   python scripts/run_decode.py > gpurun_out/r02_edge.log 2>&1
tail -2 gpurun_out/r02_edge.log
ncu -i gpurun_out/r02_edge.ncu-rep --page raw --csv > gpurun_out/r02_edge_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_enc_raw.csv > gpurun_out/r02_enc_summary.txt
python scripts/ncu_summary.py gpurun_out/r02_edge_raw.csv > gpurun_out/r02_edge_summary.txt
cat gpurun_out/r02_edge_summary.txt | head -60
rm -f gpurun_out/r02_enc.ncu-rep gpurun_out/r02_edge.ncu-rep
