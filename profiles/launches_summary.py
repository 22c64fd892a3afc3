#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python profiles/launches_summary.py file.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
d = defaultdict(list)
for r in rows[1:]:
    try:
        d[r[ki]].append(float(r[vi].replace(',', '')))
    except ValueError:
        pass
tot = sum(sum(v) for v in d.values())
print(f"{'kernel':78s} {'n':>5s} {'avg us':>9s} {'share':>7s}")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:78]:78s} {len(v):5d} {sum(v) / len(v) / 1e3:9.2f} {sum(v) / tot:7.3f}")
