#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the
count, mean and min duration in microseconds, in order of first appearance; with --sequence, the
launches in order (name, grid, block, us)."""
import csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("ggs::<unnamed>::", "")
    rows.append((name, r["Grid Size"], r["Block Size"], float(r["Metric Value"].replace(",", "")) / 1e3))
if "--sequence" in sys.argv:
    for n, g, b, us in rows:
        if not n.startswith(("at::", "native::", "cuda::")) or "--all" in sys.argv:
            print(f"{n:40s} grid {g:>14s} block {b:>12s} {us:9.2f} us")
else:
    seen = {}
    for n, g, b, us in rows:
        seen.setdefault(n, []).append(us)
    for n, v in seen.items():
        print(f"{n:60s} x{len(v):4d}  mean {sum(v) / len(v):9.2f} us  min {min(v):9.2f} us")
