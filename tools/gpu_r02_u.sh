#!/bin/bash
# round 2: source-level ncu of the dense forward kernel (one launch), converted to CSV on the box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --workload rrr --mode dense --steps 1 --warmup 1 --dropin-e2e 0 --no-cpu-baseline --no-parity"
timeout 600 $CMD > gpurun_out/r02u_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rrr_fwd_dense_pair_kernel' -s 2 -c 1 -o /tmp/r02u_prof -f $CMD > gpurun_out/r02u_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02u_prof.ncu-rep --page source --csv > gpurun_out/r02u_fwd_dense_source.csv 2>/dev/null
ncu -i /tmp/r02u_prof.ncu-rep --page raw --csv > gpurun_out/r02u_fwd_dense_raw.csv 2>/dev/null
ls -la gpurun_out/r02u_*
