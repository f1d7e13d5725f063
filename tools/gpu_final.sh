#!/bin/bash
# what the driver runs at round end: smoke, gpu tests, default bench, reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?" >> gpurun_out/bench_default.err
python bench.py --workload linear > gpurun_out/bench_linear.json 2> gpurun_out/bench_linear.err; echo "rc=$?" >> gpurun_out/bench_linear.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?" >> gpurun_out/bench_ref.err
echo done
