#!/bin/bash
# round 2, last check: full -m gpu suite, smoke, default bench line (as the driver runs them)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02s_pytest.log | head
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02s_bench.json") if l.startswith("{")][-1]
r = d["roofline"]
print("ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 1), d["e2e"]["ms_each_rank0"], "parity", d["parity"]["fit_rel_diff"])
print("roofline frac", r["frac"], "executed_frac", r.get("executed_frac"), "traffic", r["traffic"], "algorithmic bytes", r.get("algorithmic_bytes_per_launch"))
print("config lbfgs:", d["config"]["lbfgs"])
print("linear", d["linear"]["ms_per_step"], d["linear"]["e2e"]["ms_per_step"])
PY
