#!/bin/bash
# round 2: exact mode for sessions wider than 160 neurons (neuron groups): test + a full-size N = 436 fit against float64
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "wide_session or exact_mode or default_mode" > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02w_pytest.log | head
timeout 900 python bench.py --workload rrr --neurons 436 --steps 3 --warmup 2 --dropin-e2e 0 --no-cpu-baseline > gpurun_out/r02w_bench_n436.json 2> gpurun_out/r02w_bench_n436.err; echo "bench N=436 rc=$?"
tail -3 gpurun_out/r02w_bench_n436.err | cut -c1-300
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02w_bench_n436.json") if l.startswith("{")][-1]
print("N=436: ms", round(d["ms_per_step"], 2), "mode", d["config"]["operand_mode"], "e2e", round(d["e2e"]["ms_per_step"], 1), "parity", (d.get("parity") or {}))
PY
