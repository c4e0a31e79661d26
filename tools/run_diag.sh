#!/bin/bash
# Run every diagnostic family in its own process with a hard timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/diag_gpu.txt 2>&1
for fam in "$@"; do
  echo "##### $fam"
  timeout 300 python tools/gpu_diag.py $fam > gpurun_out/diag_$fam.log 2>&1
  echo "exit $?" >> gpurun_out/diag_$fam.log
  tail -n 40 gpurun_out/diag_$fam.log
done
