#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
A=exact B=dense SEED=1 timeout 600 python tools/mode_diff_probe.py > gpurun_out/r02f_modediff_seed1.log 2>&1; grep -E "^eval|vs float64" gpurun_out/r02f_modediff_seed1.log | cut -c1-250
