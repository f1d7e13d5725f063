#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab_*.json gpurun_out/ab_*.err
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "rrr" > gpurun_out/pytest_ab.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ab.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; }
if grep -q "^rc=0" gpurun_out/pytest_ab.log; then
run pair_a X=1
run fact_a VS_RRR_DENSE=0
run pair_b X=1
run fact_b VS_RRR_DENSE=0
fi
echo done
