#!/bin/bash
# round 2: does the default mode's deviation on its two worst sessions (seeds 4, 7) fall with shorter accumulation runs?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for run in 36 18; do
  for seed in 4 7 1; do
    VS_RRR_RUN_EXACT=$run SEED=$seed MODES=exact timeout 600 python tools/parity_probe.py > gpurun_out/r02q_parity_run${run}_seed$seed.log 2>&1
    echo "run $run seed $seed: $(grep -E '^exact' gpurun_out/r02q_parity_run${run}_seed$seed.log | cut -c1-230)"
  done
done
