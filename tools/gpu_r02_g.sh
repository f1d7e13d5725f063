#!/bin/bash
# round 2: full GPU suite + parity over 4 seeds (exact: 4 runs, dense: 8 runs) + default bench + e2e phases
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED" gpurun_out/r02g_pytest.log | head -20
for seed in 0 1 2 3; do
  SEED=$seed MODES=exact,dense timeout 600 python tools/parity_probe.py > gpurun_out/r02g_parity_seed$seed.log 2>&1; grep -E "^dense|^exact" gpurun_out/r02g_parity_seed$seed.log | cut -c1-330
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload rrr --mode dense --steps 5 --warmup 3 --dropin-e2e 0 --no-cpu-baseline > gpurun_out/r02g_bench_dense.json 2> gpurun_out/r02g_bench_dense.err; echo "bench dense rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02g_bench.json", "gpurun_out/r02g_bench_dense.json"):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, "ms", round(d["ms_per_step"], 2), "e2e mean", round(d["e2e"]["ms_per_step"], 1), d["e2e"]["ms_each_rank0"], "parity", (d.get("parity") or {}).get("fit_rel_diff"))
        r = d["roofline"]
        for b in [r] + r.get("other_kernels", []):
            print("   ", b["kernel"][:50], "avg ms", round(b["avg_launch_ms"], 4), "n", b["launches"], "share", round(b["share_of_step"], 3), "frac", round(b["frac"], 3))
        if d.get("linear"): print("   linear ms", d["linear"]["ms_per_step"], "e2e", d["linear"]["e2e"]["ms_per_step"], "frac", d["linear"]["roofline"]["frac"])
        print("   cpu", (d.get("cpu_baseline") or {}).get("value"), "dropin", (d["e2e"].get("dropin_fp64") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY
MODE=exact timeout 600 python tools/profile_e2e.py > gpurun_out/r02g_e2e_phases_exact.txt 2>&1; grep -E "^[0-9] |^api" gpurun_out/r02g_e2e_phases_exact.txt | cut -c1-300 | head -16
