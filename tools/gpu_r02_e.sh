#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
A=exact B=dense SEED=1 timeout 600 python tools/mode_diff_probe.py > gpurun_out/r02e_modediff_seed1.log 2>&1; grep "^eval" gpurun_out/r02e_modediff_seed1.log
A=exact B=dense SEED=0 timeout 600 python tools/mode_diff_probe.py > gpurun_out/r02e_modediff_seed0.log 2>&1; grep "^eval" gpurun_out/r02e_modediff_seed0.log | head -30
VS_LBFGS_COMPACT=1 timeout 600 python bench.py --workload rrr --mode dense --steps 5 --warmup 3 --dropin-e2e 0 --no-cpu-baseline --no-parity > gpurun_out/r02e_bench_dense.json 2> gpurun_out/r02e_bench_dense.err
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02e_bench_dense.json") if l.startswith("{")][-1]
print("dense ms", round(d["ms_per_step"], 2)); r = d["roofline"]
for b in [r] + r.get("other_kernels", []):
    print("   ", b["kernel"][:50], "avg ms", round(b["avg_launch_ms"], 4), "n", b["launches"], "share", round(b["share_of_step"], 3), "frac", round(b["frac"], 3))
PY
