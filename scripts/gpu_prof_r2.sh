#!/bin/bash
# Round-2 profile visit: ncu --set full of the dominant encoder kernel (conv2) and of the tcgen05 edge kernels, raw pages -> gpurun_out/
mkdir -p gpurun_out
N=2048 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_conv_kernel|tc_conv3|tc_conv1" -s 3 -c 4 -f -o gpurun_out/r02_enc \
   python scripts/run_mapenc.py > gpurun_out/r02_enc.log 2>&1
tail -2 gpurun_out/r02_enc.log
ncu -i gpurun_out/r02_enc.ncu-rep --page raw --csv > gpurun_out/r02_enc_raw.csv 2>/dev/null
for K in edge_fwd_tc edge_bwd_tc; do
  FT=3 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 3 -c 1 -f -o gpurun_out/r02_$K \
     python scripts/run_decode.py > gpurun_out/r02_$K.log 2>&1
  tail -1 gpurun_out/r02_$K.log
  ncu -i gpurun_out/r02_$K.ncu-rep --page raw --csv > gpurun_out/r02_${K}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_$K.ncu-rep --page source --csv > gpurun_out/r02_${K}_src.csv 2>/dev/null
  rm -f gpurun_out/r02_$K.ncu-rep
done
python scripts/ncu_summary.py gpurun_out/r02_enc_raw.csv > gpurun_out/r02_enc_summary.txt
(python scripts/ncu_summary.py gpurun_out/r02_edge_fwd_tc_raw.csv; python scripts/ncu_summary.py gpurun_out/r02_edge_bwd_tc_raw.csv) > gpurun_out/r02_edge_summary.txt
for K in edge_fwd_tc edge_bwd_tc; do echo "# stall samples $K"; python scripts/ncu_stalls.py gpurun_out/r02_${K}_src.csv 0.03 | head -12; done >> gpurun_out/r02_edge_summary.txt
rm -f gpurun_out/r02_enc.ncu-rep
