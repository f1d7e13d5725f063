#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/plain_dense.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rrr_bwd_dense_kernel|gemm_tn_pair_kernel' -s 20 -c 4 -o gpurun_out/prof_rrr_dense -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_dense.log 2>&1
echo done
