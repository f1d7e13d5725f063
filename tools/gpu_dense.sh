#!/bin/bash
# dense RRR backward: parity tests first, then the fit with the dense route off / on
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "rrr" > gpurun_out/pytest_dense.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dense.log
if grep -q "^rc=0" gpurun_out/pytest_dense.log; then
  VS_RRR_DENSE=0 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/dense_off.json 2> gpurun_out/dense_off.err
  timeout 300 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/dense_on.json 2> gpurun_out/dense_on.err
  VS_RRR_DENSE=0 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/dense_off2.json 2> gpurun_out/dense_off2.err
  timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/dense_on2.json 2> gpurun_out/dense_on2.err
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
fi
echo done
