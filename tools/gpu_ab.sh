#!/bin/bash
# quick A/B of the RRR fit: VS_RRR_DENSE=0 (factorised backward) vs default
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -m gpu -x -q -k "rrr or gemm" > gpurun_out/pytest_ab.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ab.log
if grep -q "^rc=0" gpurun_out/pytest_ab.log; then
  for i in 1 2; do
    VS_RRR_DENSE=0 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/ab_off$i.json 2> gpurun_out/ab_off$i.err
    timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/ab_on$i.json 2> gpurun_out/ab_on$i.err
  done
fi
echo done
