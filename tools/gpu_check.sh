#!/bin/bash
# full GPU test suite, then the RRR bench twice (default route) and once with the dense backward
mkdir -p gpurun_out; rm -f gpurun_out/ab_*.json gpurun_out/ab_*.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-parity > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; }
run def_a X=1
run dense_a VS_RRR_DENSE=1
run def_b X=1
run dense_spin_b VS_RRR_DENSE=1 VS_DENSE_FLAGS=1
run def_c X=1
echo done
