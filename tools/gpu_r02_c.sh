#!/bin/bash
# round 2: exact + dense modes with bounded accumulation runs: tests, parity on 4 seeds, bench lines, ncu launch lists
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -k "exact_mode or fused_loader or default_mode" > gpurun_out/r02c_pytest_exact.log 2>&1; echo "pytest exact rc=$?"
grep -E "passed|failed|^FAILED" gpurun_out/r02c_pytest_exact.log | head -20
for seed in 0 1 2 3; do
  SEED=$seed MODES=dense,exact timeout 600 python tools/parity_probe.py > gpurun_out/r02c_parity_seed$seed.log 2>&1; grep -E "^dense|^exact" gpurun_out/r02c_parity_seed$seed.log
done
for mode in dense exact; do
  timeout 600 python bench.py --workload rrr --mode $mode --steps 5 --warmup 3 --dropin-e2e 0 --no-cpu-baseline > gpurun_out/r02c_bench_$mode.json 2> gpurun_out/r02c_bench_$mode.err; echo "bench $mode rc=$?"
done
python - <<'PY'
import json
for f in ("gpurun_out/r02c_bench_dense.json", "gpurun_out/r02c_bench_exact.json"):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"]["ms_each_rank0"], "parity", (d.get("parity") or {}).get("fit_rel_diff"))
        r = d["roofline"]
        for b in [r] + r.get("other_kernels", []):
            print("   ", b["kernel"][:50], "avg ms", round(b["avg_launch_ms"], 4), "n", b["launches"], "share", round(b["share_of_step"], 3), "frac", round(b["frac"], 3))
    except Exception as e:
        print(f, "unreadable", e)
PY
# launch lists (cold-cache, serialised): one fit per mode
for mode in dense exact; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02c_launches_$mode.csv \
      python bench.py --workload rrr --mode $mode --steps 1 --warmup 3 --dropin-e2e 0 --no-cpu-baseline --no-parity > gpurun_out/r02c_ncu_$mode.log 2>&1; echo "ncu $mode rc=$?"
done
python tools/summarize_launches.py gpurun_out/r02c_launches_dense.csv 2>/dev/null | head -30
python tools/summarize_launches.py gpurun_out/r02c_launches_exact.csv 2>/dev/null | head -30
