#!/bin/bash
# round 2, first exact-mode check: GPU tests, parity probe of every mode, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python tools/parity_probe.py > gpurun_out/r02a_parity.log 2>&1; echo "parity rc=$?"
cat gpurun_out/r02a_parity.log | tail -12
SEED=1 MODES=exact,exact32,p1 timeout 600 python tools/parity_probe.py > gpurun_out/r02a_parity_seed1.log 2>&1
tail -4 gpurun_out/r02a_parity_seed1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02a_bench.err
head -c 3000 gpurun_out/r02a_bench.json
