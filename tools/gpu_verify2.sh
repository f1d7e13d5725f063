#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_drivers.py -m gpu -x -q -k "rrr" > gpurun_out/pytest_rrr.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rrr.log
python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/bench_rrr.json 2> gpurun_out/bench_rrr.err; echo "rc=$?" >> gpurun_out/bench_rrr.err
python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-parity > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pack|colstats|smooth' -c 40 --csv --log-file gpurun_out/launches_pack.csv \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-parity > /dev/null 2>&1
echo done
