#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
per kernel (time share, and DRAM bytes when captured).
usage: summarize_launches.py <launches.csv> "<comment>" > profiles/<name>_summary.csv"""
import csv, re, sys
from collections import defaultdict, OrderedDict

with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
L = OrderedDict()
for r in csv.DictReader(lines):
    i = int(r["ID"])
    name = re.sub(r"\(.*$", "", re.sub(r"^void\s+", "", r["Kernel Name"]))
    e = L.setdefault(i, {"name": name, "us": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "")
    m = r.get("Metric Name")
    if m == "gpu__time_duration.sum":
        e["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        e["rd" if m.endswith("read.sum") else "wr"] = v
rows = list(L.values())
tot = sum(e["us"] for e in rows)
rd = sum(e["rd"] for e in rows) ; wr = sum(e["wr"] for e in rows)
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for e in rows:
    a = agg[e["name"]] ; a[0] += 1 ; a[1] += e["us"] ; a[2] += e["rd"] ; a[3] += e["wr"]
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print(f"# {len(rows)} launches captured, total {tot/1e3:.2f} ms (cold-cache, serialised under ncu: compare shares)")
if rd + wr > 0:
    print(f"# DRAM traffic of the captured launches: {rd/1e9:.3f} GB read + {wr/1e9:.3f} GB written = {(rd+wr)/1e9:.3f} GB")
print("kernel,launches,total_us,share,avg_us,dram_read_MB,dram_write_MB")
for n, (c, v, r_, w_) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},{c},{v:.1f},{v/tot:.4f},{v/c:.2f},{r_/1e6:.2f},{w_/1e6:.2f}")
