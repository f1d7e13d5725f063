#!/bin/bash
# round 2: whole-fit parity of the default mode on four more sessions (seeds 4..7)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for seed in 4 5 6 7; do
  SEED=$seed MODES=exact timeout 600 python tools/parity_probe.py > gpurun_out/r02p_parity_seed$seed.log 2>&1; grep -E "^fp64|^exact" gpurun_out/r02p_parity_seed$seed.log | cut -c1-200
done
