#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/bench_rrr.json 2> gpurun_out/bench_rrr.err; echo "rc=$?" >> gpurun_out/bench_rrr.err
echo done
