#!/bin/bash
# round 2: ncu --set full of the rewritten kernels (epi_fx, epi_bx, prep_u<3>, pack_fused) + smoke + the e2e phase profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02j_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02j_smoke.log
CMD="python bench.py --workload rrr --steps 1 --warmup 3 --dropin-e2e 0 --no-cpu-baseline --no-parity"
timeout 600 $CMD > gpurun_out/r02j_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none \
    -k regex:'pack_fused_kernel|epi_fx_kernel|epi_bx_kernel|prep_u_kernel|update_compact_kernel|finalize_kernel' -c 10 \
    -o /tmp/r02j_prof -f $CMD > gpurun_out/r02j_ncu_full.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02j_prof.ncu-rep --page raw --csv > gpurun_out/r02j_ncu_full_epilogues.csv 2>/dev/null
MODE=exact timeout 600 python tools/profile_e2e.py > gpurun_out/r02j_e2e_phases_exact.txt 2>&1; grep -E "^[0-9] |^api" gpurun_out/r02j_e2e_phases_exact.txt | cut -c1-300 | tail -8
