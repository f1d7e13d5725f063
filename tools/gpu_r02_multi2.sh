#!/bin/bash
# round 2: re-validation of the trial-sharded path after the init broadcast (2 GPUs)
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02m2_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
grep -E "passed|failed|skipped|^E  |^FAILED" gpurun_out/r02m2_pytest_multi.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --workload rrr --strong --dropin-e2e 0 \
      > gpurun_out/r02m2_strong_$N.json 2> gpurun_out/r02m2_strong_$N.err; echo "strong rc=$?"
tail -c 300 gpurun_out/r02m2_strong_$N.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02m2_strong_$N.json") if l.startswith("{")][-1])
print("strong n_gpus", d["n_gpus"], "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"]["ms_each_rank0"], "parity", (d.get("parity") or {}).get("fit_rel_diff"))
PY
