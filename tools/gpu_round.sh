#!/bin/bash
# One GPU round trip: parity tests, both bench workloads, ncu launch lists (run under gpurun).
#   tools/gpu_round.sh [ncu] [full]
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --workload rrr --steps 3 > gpurun_out/bench_rrr.json 2> gpurun_out/bench_rrr.err
python bench.py --workload linear --steps 30 > gpurun_out/bench_linear.json 2> gpurun_out/bench_linear.err
if [[ " $* " == *" ncu "* ]]; then
  python bench.py --workload rrr --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/plain_rrr.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_rrr.csv \
      python bench.py --workload rrr --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_rrr.log 2>&1
  python bench.py --workload linear --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_lin.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_linear.csv \
      python bench.py --workload linear --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lin.log 2>&1
fi
if [[ " $* " == *" full "* ]]; then
  python bench.py --workload rrr --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/plain_rrr2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 20 -c 2 -o gpurun_out/prof_rrr_gemm -f \
      python bench.py --workload rrr --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_full_rrr.log 2>&1
  python bench.py --workload linear --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_lin2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:dw_adamw_kernel -s 3 -c 1 -o gpurun_out/prof_lin_dw -f \
      python bench.py --workload linear --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_lin.log 2>&1
fi
echo done
