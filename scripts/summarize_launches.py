#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time, share."""
import csv, re, sys
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
tot = {}
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    e = tot.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += v
allus = sum(v[1] for v in tot.values())
print("kernel,launches,total_us,share")
for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('"%s",%d,%.1f,%.4f' % (k, c, us, us / allus if allus else 0))
print('"TOTAL",%d,%.1f,1.0' % (sum(v[0] for v in tot.values()), allus))
