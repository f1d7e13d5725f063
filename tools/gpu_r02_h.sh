#!/bin/bash
# round 2: does the dense mode's whole-fit deviation fall with shorter TMEM accumulation runs of its forward?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for run in 18 9; do
  for seed in 0 1 2 3; do
    VS_RRR_RUN_DENSE=$run SEED=$seed MODES=dense timeout 600 python tools/parity_probe.py > gpurun_out/r02h_parity_run${run}_seed$seed.log 2>&1
    echo "run $run seed $seed: $(grep -E '^dense' gpurun_out/r02h_parity_run${run}_seed$seed.log | cut -c1-200)"
  done
done
