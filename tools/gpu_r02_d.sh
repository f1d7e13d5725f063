#!/bin/bash
# round 2: dV accumulation fix + end-of-fit probes, compact-history L-BFGS, fit times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lbfgs.py tests/test_gpu_models.py -m gpu -q -k "compact or exact_mode" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED" gpurun_out/r02d_pytest.log | head -20
for seed in 0 1 2 3; do
  SEED=$seed MODES=dense,exact timeout 600 python tools/parity_probe.py > gpurun_out/r02d_parity_seed$seed.log 2>&1; grep -E "^dense|^exact" gpurun_out/r02d_parity_seed$seed.log
done
echo "--- compact history"
for seed in 1 3; do
  VS_LBFGS_COMPACT=1 SEED=$seed MODES=dense,exact timeout 600 python tools/parity_probe.py > gpurun_out/r02d_parity_compact_seed$seed.log 2>&1; grep -E "^dense|^exact" gpurun_out/r02d_parity_compact_seed$seed.log
done
for cfg in "exact 0" "exact 1" "dense 1"; do
  set -- $cfg
  VS_LBFGS_COMPACT=$2 timeout 600 python bench.py --workload rrr --mode $1 --steps 5 --warmup 3 --dropin-e2e 0 --no-cpu-baseline --no-parity > gpurun_out/r02d_bench_$1_c$2.json 2> gpurun_out/r02d_bench_$1_c$2.err; echo "bench $1 compact=$2 rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02d_bench_*.json")):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"]["ms_each_rank0"])
        r = d["roofline"]
        for b in [r] + r.get("other_kernels", []):
            print("   ", b["kernel"][:50], "avg ms", round(b["avg_launch_ms"], 4), "n", b["launches"], "share", round(b["share_of_step"], 3), "frac", round(b["frac"], 3))
    except Exception as e:
        print(f, "unreadable", e)
PY
