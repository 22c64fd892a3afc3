#!/usr/bin/env python
"""usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python profiles/source_hotspots.py src.csv [N]
Aggregates warp-stall samples / executed instructions per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[2]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
data = []
for r in rows[3:]:
    if len(r) < 40 or r[0] == '':
        continue
    try:
        s = int(r[ix['# Samples']]); inst = int(r[ix['Instructions Executed']]); tin = int(r[ix['Thread Instructions Executed']])
    except ValueError:
        continue
    data.append((int(r[0]), s, inst, tin, r[1].strip()[:100], r))
tot_s = sum(d[1] for d in data); tot_i = sum(d[2] for d in data)
print('total samples', tot_s, 'warp instructions', tot_i)
for d in sorted(data, key=lambda d: -d[1])[:top]:
    r = d[5]
    print(f"{d[0]:5d} smp {100*d[1]/tot_s:5.1f}% inst {100*d[2]/tot_i:5.1f}% thr/inst {d[3]/max(d[2],1):5.1f} "
          f"long_sb {r[ix['stall_long_sb']]:>5} bar {r[ix['stall_barrier']]:>4} short_sb {r[ix['stall_short_sb']]:>4} wait {r[ix['stall_wait']]:>4} | {d[4]}")
