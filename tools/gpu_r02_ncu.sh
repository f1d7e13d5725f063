#!/bin/bash
# round 2 ncu evidence (ONE GPU): launch list of one default-mode fit + --set full captures of the dominant kernels.
# Each ncu run follows a plain run of the same command line that exited 0 (B200_PROFILING.md).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --workload rrr --steps 1 --warmup 3 --dropin-e2e 0 --no-cpu-baseline --no-parity"
timeout 600 $CMD > gpurun_out/r02n_plain_exact.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02n_launches_exact.csv $CMD > gpurun_out/r02n_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 600 $CMD > gpurun_out/r02n_plain_exact2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'gemm_tn_pair_kernel|rrr_bwd_dense_pair_kernel|pack_fused_kernel|epi_f_kernel|epi_b_kernel|prep_u_kernel' -c 22 \
    -o gpurun_out/r02n_prof_exact -f $CMD > gpurun_out/r02n_ncu_full_exact.log 2>&1
echo "full exact rc=$?"
CMD2="python bench.py --workload rrr --mode dense --steps 1 --warmup 3 --dropin-e2e 0 --no-cpu-baseline --no-parity"
timeout 600 $CMD2 > gpurun_out/r02n_plain_dense.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'rrr_fwd_dense_pair_kernel|rrr_bwd_dense_pair_kernel|epi_d_kernel|dv_reduce_kernel' -s 4 -c 8 \
    -o gpurun_out/r02n_prof_dense -f $CMD2 > gpurun_out/r02n_ncu_full_dense.log 2>&1
echo "full dense rc=$?"
CMD3="python bench.py --workload linear --steps 3 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD3 > gpurun_out/r02n_plain_linear.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02n_launches_linear.csv $CMD3 > gpurun_out/r02n_ncu_list_lin.log 2>&1
echo "launch list linear rc=$?"
ls -la gpurun_out/ | grep r02n
python tools/summarize_launches.py gpurun_out/r02n_launches_exact.csv 2>/dev/null | head -40
