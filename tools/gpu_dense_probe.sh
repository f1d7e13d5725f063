#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/dense_probe.log
for cfg in "64 200 21" "16 260 160" "128 300 144" "100 520 5" "30 200 21" "37 300 144"; do
  echo "=== $cfg" >> gpurun_out/dense_probe.log
  CUDA_LAUNCH_BLOCKING=1 timeout 120 python tools/dense_probe.py $cfg >> gpurun_out/dense_probe.log 2>&1
  echo "rc=$?" >> gpurun_out/dense_probe.log
done
echo done
