#!/bin/bash
# full GPU test suite, then the RRR bench for each backward route
mkdir -p gpurun_out; rm -f gpurun_out/ab_*.json gpurun_out/ab_*.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; }
if grep -q "pytest rc=0" gpurun_out/pytest_gpu.log; then
run pair_a X=1
run half_a VS_RRR_DENSE=1
run fact_a VS_RRR_DENSE=0
run pair_b X=1
run half_b VS_RRR_DENSE=1
run fact_b VS_RRR_DENSE=0
fi
echo done
