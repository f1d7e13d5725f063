#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_drivers.py tests/test_gpu_lbfgs.py -m gpu -x -q > gpurun_out/pytest_rrr.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rrr.log
python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_rrr.json 2> gpurun_out/bench_rrr.err; echo "rc=$?" >> gpurun_out/bench_rrr.err
echo done
