#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab_*.json
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-parity > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; }
run off_a VS_RRR_DENSE=0
run f0_a VS_DENSE_FLAGS=0
run f1_a VS_DENSE_FLAGS=1
run f2_a VS_DENSE_FLAGS=2
run f2p8_a VS_DENSE_FLAGS=2 VS_DENSE_PREFETCH=8
run f2p64_a VS_DENSE_FLAGS=2 VS_DENSE_PREFETCH=64
run f3_a VS_DENSE_FLAGS=3
run off_b VS_RRR_DENSE=0
run f0_b VS_DENSE_FLAGS=0
run f2_b VS_DENSE_FLAGS=2
echo done
