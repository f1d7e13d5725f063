#!/bin/bash
# round 2: the loader's overflow-flag test (fused and two-kernel) + smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "train_constant or fused_loader" > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02o_pytest.log | head
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02o_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02o_smoke.log
