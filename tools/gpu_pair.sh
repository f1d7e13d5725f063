#!/bin/bash
# CTA-pair GEMM kernel: correctness (kernel + model tests), then the RRR fit with the pair kernel on and off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > gpurun_out/pytest_pair_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_pair_gemm.log
if grep -q "rc=0" gpurun_out/pytest_pair_gemm.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  VS_GEMM_PAIR=0 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/pair_off.json 2> gpurun_out/pair_off.err
  timeout 300 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/pair_on.json 2> gpurun_out/pair_on.err
  VS_GEMM_PAIR=0 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/pair_off2.json 2> gpurun_out/pair_off2.err
  timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/pair_on2.json 2> gpurun_out/pair_on2.err
fi
echo done
