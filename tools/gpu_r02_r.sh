#!/bin/bash
# round 2: the reference arm exactly as the driver launches it (default budget), timed
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02r_ref.json 2> gpurun_out/r02r_ref.err; echo "reference arm rc=$? wall ${SECONDS}s"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02r_ref.json") if l.startswith("{")][-1])
print("value", d["value"], "ms_per_step", d["ms_per_step"], "steps", d["steps"], "neurons", d["config"]["neurons"], d["config"]["extrapolation"])
PY
