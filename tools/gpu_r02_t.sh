#!/bin/bash
# round 2: dense forward with item-ahead prefetch of the U records: dense-mode tests + dense bench line (forward kernel time under tag 3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "dense or exact_mode" > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02t_pytest.log | head
timeout 600 python bench.py --workload rrr --mode dense --steps 5 --warmup 3 --dropin-e2e 0 --no-cpu-baseline > gpurun_out/r02t_bench_dense.json 2> gpurun_out/r02t_bench_dense.err; echo "bench dense rc=$?"
python - <<'PY'
import json
d = [json.loads(l) for l in open("gpurun_out/r02t_bench_dense.json") if l.startswith("{")][-1]
print("ms", round(d["ms_per_step"], 2), "parity", (d.get("parity") or {}).get("fit_rel_diff"))
r = d["roofline"]
for b in [r] + r.get("other_kernels", []):
    print("   ", b["kernel"][:50], "avg ms", round(b["avg_launch_ms"], 4), "n", b["launches"], "share", round(b["share_of_step"], 3), "frac", round(b["frac"], 3))
PY
