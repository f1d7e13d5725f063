#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q -k "gemm or linear or mlp" > gpurun_out/pytest_lin.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_lin.log
for g in 1 2 4; do
  VS_GEMM_KBGROUP=$g python bench.py --workload linear --steps 40 --no-cpu-baseline > gpurun_out/lin_kbg$g.json 2> gpurun_out/lin_kbg$g.err
  VS_GEMM_KBGROUP=$g python bench.py --workload linear --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 &&
  VS_GEMM_KBGROUP=$g ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 -k regex:gemm_tn_kernel --csv --log-file gpurun_out/lin_kbg${g}_launches.csv \
    python bench.py --workload linear --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
done
echo done
