#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the product library: python profiles/sass_histogram.py [lib.so] > profiles/rNN_sass_opcodes.txt
Shows the architecture of the cubin (sm_100a), registers / spills per kernel (cuobjdump -res-usage) and the opcodes that prove
the Blackwell-specific paths: UBLKCP (cp.async.bulk, the TMA bulk-copy engine), SYNCS (mbarrier), LDG.E.*.256 (32-byte loads),
ACQBULK / griddepcontrol (programmatic dependent launch), REDG / ATOMG, DADD/DMUL/DFMA (fp64 path)."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "elimaloc_b200", "libelimaloc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in line:
        usage[cur] = line.strip()
        cur = None
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print("library:", os.path.basename(lib), "| cubin architectures:", ", ".join(arch))
kern = None
hist = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
demangle = subprocess.run(["cu++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
for (k, h), name in zip(hist.items(), demangle):
    total = sum(h.values())
    base = collections.Counter()
    for op, c in h.items():
        base[op.split(".")[0]] += c
    marks = {t: sum(c for op, c in h.items() if re.search(t, op)) for t in
             (r"^UBLKCP", r"^SYNCS", r"^LDG.*\.256", r"^LDG", r"^LDGSTS", r"^ACQBULK|^PREEXIT|^DEPBAR", r"^ATOM|^RED", r"^D(ADD|MUL|FMA)", r"^F(ADD|MUL|FMA)", r"^SHFL", r"^BAR", r"^STL|^LDL")}
    print(f"\n== {name[:150]}\n   {total} instructions | {usage.get(k, '')}")
    print("   " + "  ".join(f"{t.replace('^', '').replace('\\\\', '')}={c}" for t, c in marks.items()))
    print("   top: " + ", ".join(f"{op} {c}" for op, c in base.most_common(14)))
