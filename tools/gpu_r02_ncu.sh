#!/bin/bash
# round 2 ncu evidence (ONE GPU): launch list of one default-mode fit + --set full captures of the dominant kernels.
# Each ncu run follows a plain run of the same command line that exited 0 (B200_PROFILING.md).  The .ncu-rep files are
# converted to CSV on the box and removed (gpurun_out/ is capped at 64 MiB).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --workload rrr --steps 1 --warmup 3 --dropin-e2e 0 --no-cpu-baseline --no-parity"
if [ "$1" != "nolist" ]; then
timeout 600 $CMD > gpurun_out/r02n_plain_exact.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02n_launches_exact.csv $CMD > gpurun_out/r02n_ncu_list.log 2>&1
echo "launch list rc=$?"
fi
timeout 600 $CMD > gpurun_out/r02n_plain_exact2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none \
    -k regex:'gemm_tn_pair_kernel|rrr_bwd_dense_pair_kernel|pack_fused_kernel|epi_f_kernel|epi_b_kernel|prep_u_kernel|dots_kernel|direction_kernel' -c 12 \
    -o /tmp/r02n_prof_exact -f $CMD > gpurun_out/r02n_ncu_full_exact.log 2>&1
echo "full exact rc=$?"
ncu -i /tmp/r02n_prof_exact.ncu-rep --page raw --csv > gpurun_out/r02n_ncu_full_exact.csv 2>/dev/null; ls -la /tmp/r02n_prof_exact.ncu-rep
CMD2="python bench.py --workload rrr --mode dense --steps 1 --warmup 3 --dropin-e2e 0 --no-cpu-baseline --no-parity"
timeout 600 $CMD2 > gpurun_out/r02n_plain_dense.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none \
    -k regex:'rrr_fwd_dense_pair_kernel|rrr_bwd_dense_pair_kernel|epi_d_kernel|dv_reduce_kernel' -s 4 -c 6 \
    -o /tmp/r02n_prof_dense -f $CMD2 > gpurun_out/r02n_ncu_full_dense.log 2>&1
echo "full dense rc=$?"
ncu -i /tmp/r02n_prof_dense.ncu-rep --page raw --csv > gpurun_out/r02n_ncu_full_dense.csv 2>/dev/null
ls -la gpurun_out/ | grep r02n
# per-evaluation probe of both modes against the float64 truth along one exact-driven fit (current accumulation-run defaults)
for seed in 1 3; do
  A=exact B=dense SEED=$seed timeout 600 python tools/mode_diff_probe.py > gpurun_out/r02n_mode_diff_seed$seed.txt 2>&1; echo "mode diff seed $seed rc=$?"
done
grep -E "vs float64" gpurun_out/r02n_mode_diff_seed3.txt | cut -c1-200 | head -22
