#!/bin/bash
# round 2: after the epilogue rewrites (epi_fx, epi_bx, prep_u fast path) and the loader changes: full GPU suite, parity of
# the default mode over 4 seeds, default bench, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02n2_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02n2_pytest.log | head -20
for seed in 0 1 2 3; do
  SEED=$seed MODES=exact timeout 600 python tools/parity_probe.py > gpurun_out/r02n2_parity_seed$seed.log 2>&1; grep -E "^exact" gpurun_out/r02n2_parity_seed$seed.log | cut -c1-330
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02n2_bench.json 2> gpurun_out/r02n2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02n2_bench.json",):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, "ms", round(d["ms_per_step"], 2), "e2e mean", round(d["e2e"]["ms_per_step"], 1), d["e2e"]["ms_each_rank0"], "parity", (d.get("parity") or {}).get("fit_rel_diff"))
        r = d["roofline"]
        for b in [r] + r.get("other_kernels", []):
            print("   ", b["kernel"][:50], "avg ms", round(b["avg_launch_ms"], 4), "n", b["launches"], "share", round(b["share_of_step"], 3), "frac", round(b["frac"], 3))
        if d.get("linear"): print("   linear ms", d["linear"]["ms_per_step"], "e2e", d["linear"]["e2e"]["ms_per_step"], "frac", d["linear"]["roofline"]["frac"])
        print("   cpu", (d.get("cpu_baseline") or {}), "dropin", (d["e2e"].get("dropin_fp64") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY
CMD="python bench.py --workload rrr --steps 1 --warmup 3 --dropin-e2e 0 --no-cpu-baseline --no-parity"
timeout 600 $CMD > gpurun_out/r02n2_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02n2_launches_exact.csv $CMD > gpurun_out/r02n2_ncu_list.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r02n2_launches_exact.csv 2>/dev/null | head -16
