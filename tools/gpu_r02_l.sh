#!/bin/bash
# round 2: source-level ncu of the fused loader (one launch, source page converted to CSV on the box) + regression tests of the guards
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02l_pytest.log | head
CMD="python bench.py --workload rrr --steps 1 --warmup 1 --dropin-e2e 0 --no-cpu-baseline --no-parity"
timeout 600 $CMD > gpurun_out/r02l_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pack_fused_kernel' -c 1 -o /tmp/r02l_prof -f $CMD > gpurun_out/r02l_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02l_prof.ncu-rep --page source --csv > gpurun_out/r02l_pack_fused_source.csv 2>/dev/null
ncu -i /tmp/r02l_prof.ncu-rep --page raw --csv > gpurun_out/r02l_pack_fused_raw.csv 2>/dev/null
ls -la gpurun_out/r02l_*
