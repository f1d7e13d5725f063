#!/bin/bash
# round 2: loader after batching its staging loads: bit-exactness tests + its duration and DRAM bytes under ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "fused_loader or exact_mode or default_mode or pack" > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02m_pytest.log | head
CMD="python bench.py --workload rrr --steps 1 --warmup 1 --dropin-e2e 0 --no-cpu-baseline --no-parity"
timeout 600 $CMD > gpurun_out/r02m_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pack_fused_kernel' -c 2 -o /tmp/r02m_prof -f $CMD > gpurun_out/r02m_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02m_prof.ncu-rep --page source --csv > gpurun_out/r02m_pack_fused_source.csv 2>/dev/null
ncu -i /tmp/r02m_prof.ncu-rep --page raw --csv > gpurun_out/r02m_pack_fused_raw.csv 2>/dev/null
python - <<'PY'
import csv
r = list(csv.reader(open("gpurun_out/r02m_pack_fused_raw.csv"))); h = r[0]; u = r[1]
for row in r[2:]:
    print(" | ".join(row[h.index(w)][:40] + " " + u[h.index(w)] for w in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread")))
PY
