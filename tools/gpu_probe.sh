#!/bin/bash
# ad-hoc probe round: tests, e2e phase breakdown, full ncu capture of the secondary RRR kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/profile_e2e.py > gpurun_out/e2e_phases.log 2>&1
python bench.py --workload rrr --steps 3 > gpurun_out/bench_rrr.json 2> gpurun_out/bench_rrr.err
python bench.py --workload rrr --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_rrr3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'epi_f_kernel|epi_b_kernel|prep_u_kernel|dots_kernel|direction_kernel|small_mats' -s 60 -c 12 -o gpurun_out/prof_rrr_misc -f \
    python bench.py --workload rrr --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_misc.log 2>&1
echo done
