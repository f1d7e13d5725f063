#!/bin/bash
# round 2, multi-GPU: pytest multi-GPU checks + the bench lines the driver's scaling run produces (joint shared-V model, N ranks)
# usage: bash tools/gpu_r02_multi.sh N
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02m_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
grep -E "passed|failed|skipped|^E  |^FAILED" gpurun_out/r02m_pytest_multi.log | head -20
run() {  # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 $2 \
      > gpurun_out/r02m_$1_$N.json 2> gpurun_out/r02m_$1_$N.err; echo "$1 rc=$?"
  tail -c 400 gpurun_out/r02m_$1_$N.err
}
run joint "--workload rrr --dropin-e2e 0"
run indep "--workload rrr --independent --dropin-e2e 0"
run both ""
run strong "--workload rrr --strong --dropin-e2e 0"
python - <<PY
import json
for nm in ("joint", "indep", "both", "strong"):
    try:
        d = json.loads([l for l in open("gpurun_out/r02m_%s_$N.json" % nm) if l.startswith("{")][-1])
        print(nm, "n_gpus", d["n_gpus"], "ms", round(d["ms_per_step"], 2), "value", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 1), d["config"]["parallelism"][:60], "scaling", d["scaling"], "parity", (d.get("parity") or {}).get("fit_rel_diff"),
              "| linear", (d.get("linear") or {}).get("ms_per_step"))
    except Exception as e:
        print(nm, "unreadable", e)
PY
