#!/bin/bash
# final-state evidence: bench lines, ncu launch list of the RRR bench, --set full capture of the two dominant kernels
mkdir -p gpurun_out
python bench.py --workload rrr --steps 5 > gpurun_out/bench_rrr.json 2> gpurun_out/bench_rrr.err
python bench.py --workload linear --steps 30 > gpurun_out/bench_linear.json 2> gpurun_out/bench_linear.err
python bench.py --workload rrr --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/plain_rrr.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_rrr.csv \
    python bench.py --workload rrr --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_rrr.log 2>&1
python bench.py --workload rrr --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/plain_rrr2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_tn_pair_kernel|rrr_bwd_dense_pair_kernel' -s 20 -c 4 -o gpurun_out/prof_rrr_gemm2 -f \
    python bench.py --workload rrr --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_full_rrr.log 2>&1
echo done
