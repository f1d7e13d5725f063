#!/bin/bash
# round 2: default bench line after the last host-side change (neuron groups)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02x_bench.json") if l.startswith("{")][-1]
print("ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 1), d["e2e"]["ms_each_rank0"], "parity", d["parity"]["fit_rel_diff"], "launches", d["gpu_launches"], "steps", d["steps"], d["warmup"])
PY
