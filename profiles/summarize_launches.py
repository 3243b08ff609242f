#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: summarize_launches.py <launches.csv> "<comment>" > profiles/<name>_summary.csv"""
import csv, re, sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    name = re.sub(r"^void\s+", "", r["Kernel Name"])
    name = re.sub(r"\(.*$", "", name)
    rows.append((name, v))
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, v in rows:
    agg[n][0] += 1
    agg[n][1] += v
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print(f"# {len(rows)} launches captured, total {tot/1e3:.2f} ms (cold-cache, serialised under ncu: compare shares)")
print("kernel,launches,total_us,share,avg_us")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},{c},{v:.1f},{v/tot:.4f},{v/c:.2f}")
