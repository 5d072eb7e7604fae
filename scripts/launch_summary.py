#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel as a markdown table.
  python scripts/launch_summary.py launches.csv "title" "command" [--grid] > profiles/rN_launches_x.md"""
import collections
import csv
import re
import sys

path, title, command = sys.argv[1], sys.argv[2], sys.argv[3]
by_grid = "--grid" in sys.argv
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
h = rows[0]
ki, vi, gi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Metric Unit")
agg = collections.OrderedDict()
total = 0.0
for r in rows[1:]:
    name = re.sub(r"^void ", "", r[ki])
    name = re.sub(r"\(.*$", "", name)[:48]
    key = (name, r[gi]) if by_grid else (name,)
    us = float(r[vi].replace(",", "")) / (1000.0 if r[ui].startswith("ns") else 1.0)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
print(f"# {title}\ncommand: {command}\n(cold-cache, serialised per-launch times: compare SHARES, not absolutes)\n")
print(f"total kernel time {total / 1000:.3f} ms over {len(rows) - 1} launches\n")
print("| kernel |" + (" grid |" if by_grid else "") + " launches | total us | avg us | share |")
print("|---|" + ("---|" if by_grid else "") + "---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k[0]} |" + (f" {k[1]} |" if by_grid else "") + f" {n} | {t:.1f} | {t / n:.1f} | {100 * t / total:.1f}% |")
