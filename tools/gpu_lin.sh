#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --workload linear --steps 40 --no-cpu-baseline > gpurun_out/bench_linear.json 2> gpurun_out/bench_linear.err; echo "rc=$?" >> gpurun_out/bench_linear.err
python bench.py --workload linear --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_lin.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_linear.csv \
    python bench.py --workload linear --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lin.log 2>&1
python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_rrr.json 2> gpurun_out/bench_rrr.err
echo done
