#!/bin/bash
# round 2: last full -m gpu suite + smoke after the dense-forward change
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02v_pytest.log | head
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.log 2>&1; echo "smoke rc=$?"
