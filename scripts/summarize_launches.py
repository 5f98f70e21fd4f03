#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((int(r["ID"]), r["Kernel Name"], ns))
rows = [r for r in rows if r[0] >= skip]
agg = defaultdict(lambda: [0, 0.0])
for _, name, ns in rows:
    short = re.sub(r"\(.*", "", name)
    agg[short][0] += 1
    agg[short][1] += ns
total = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, total {total / 1e3:.1f} us (cold-cache, serialised: compare shares)")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ns / total * 100:6.2f}%  {ns / 1e3:10.1f} us  n={n:5d}  avg {ns / n / 1e3:8.2f} us  {name}")
