#!/usr/bin/env python
"""Reads the JSON lines of profiles/ab_chain.py runs and prints `export` lines for the fastest sound variant of P2P at 131072 points
(sound: no failure, all iterations run, pose bit-identical to the established mode on the same summation grid); keeps the established
variant (async, 128 blocks, default library) unless the best one is at least 3 % faster."""
import json
import sys

rows = []
for path in sys.argv[1:]:
    try:
        for line in open(path):
            line = line.strip()
            if line.startswith("[{"):
                rows += json.loads(line)
    except OSError:
        pass
cand = [r for r in rows if r.get("method") == "p2p" and r.get("n") == 131072 and "failed" not in r and r.get("iterations") == 20
        and r.get("pose_diff_same_grid") in (None, 0.0) and r.get("pose_diff_first", 1.0) < 1e-9]
base = [r for r in cand if r["mode"] == "async" and r["grid"] == 128 and r["lib"] == "default"]
if not cand or not base:
    print("# no sound candidate: keeping the defaults")
    sys.exit(0)
best = max(cand, key=lambda r: r["it_per_s"])
if best["it_per_s"] < 1.03 * base[0]["it_per_s"]:
    best = base[0]
print(f"# winner: {best['lib']} {best['mode']} grid {best['grid']}: {best['it_per_s']:.0f} it/s (established: {base[0]['it_per_s']:.0f})")
print(f"export ELM_WARM_MODE={best['mode']}")
print(f"export ELM_ASYNC_GRID={best['grid']}")
if best["lib"] != "default":
    print(f"export ELIMALOC_B200_LIB={best['lib']}")
