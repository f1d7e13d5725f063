#!/bin/bash
# round 2, first run of the dense-forward mode + fused loader on hardware (bounded: every step under its own timeout)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -k "exact_mode or fused_loader" > gpurun_out/r02b_pytest_exact.log 2>&1; echo "pytest exact rc=$?"
grep -E "passed|failed|^E  |^FAILED" gpurun_out/r02b_pytest_exact.log | head -40
MODES=dense,exact timeout 600 python tools/parity_probe.py > gpurun_out/r02b_parity.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/r02b_parity.log
SEED=1 MODES=dense,dense32 timeout 600 python tools/parity_probe.py > gpurun_out/r02b_parity_seed1.log 2>&1; tail -3 gpurun_out/r02b_parity_seed1.log
SEED=2 MODES=dense,exact timeout 600 python tools/parity_probe.py > gpurun_out/r02b_parity_seed2.log 2>&1; tail -3 gpurun_out/r02b_parity_seed2.log
timeout 600 python bench.py --workload rrr --mode dense --steps 5 --warmup 3 --dropin-e2e 0 > gpurun_out/r02b_bench_dense.json 2> gpurun_out/r02b_bench_dense.err; echo "bench dense rc=$?"
tail -c 600 gpurun_out/r02b_bench_dense.err; head -c 2500 gpurun_out/r02b_bench_dense.json; echo
timeout 600 python bench.py --workload rrr --mode exact --steps 5 --warmup 3 --dropin-e2e 0 --no-parity --no-cpu-baseline > gpurun_out/r02b_bench_exact.json 2> gpurun_out/r02b_bench_exact.err; echo "bench exact rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02b_bench_dense.json", "gpurun_out/r02b_bench_exact.json"):
    try:
        d = json.load(open(f)); print(f, "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 1), d["e2e"]["ms_each_rank0"], "parity", (d.get("parity") or {}).get("fit_rel_diff"))
    except Exception as e:
        print(f, "unreadable", e)
PY
