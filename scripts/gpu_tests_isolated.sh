#!/bin/bash
# Runs every GPU test in its own process (a sticky CUDA error must not poison the rest); logs to gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/diag_gpu.txt
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
ids=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep '::')
: > gpurun_out/pytest_gpu.log
for t in $ids; do
  echo "=== $t" >> gpurun_out/pytest_gpu.log
  timeout 600 python -m pytest "$t" -x -q -m gpu 2>&1 | tail -40 >> gpurun_out/pytest_gpu.log
done
grep -E "^=== |passed|failed|error" gpurun_out/pytest_gpu.log | tail -60
echo "---- diag"; cat gpurun_out/diag_gpu.txt 2>/dev/null | tail -80
