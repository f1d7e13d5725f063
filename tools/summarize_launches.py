"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and count per kernel."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.DictReader(lines)
tot = defaultdict(float); cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)   # -> microseconds
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':90s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:90]:90s} {cnt[k]:8d} {v:12.1f} {v / cnt[k]:10.1f} {100 * v / total:6.1f}%")
print(f"{'TOTAL':90s} {sum(cnt.values()):8d} {total:12.1f}")
