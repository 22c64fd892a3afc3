#!/usr/bin/env python
"""One-off (not part of the test suite: ~2 minutes, ~4 GB): the oracle against the reference-sources build at BASELINE config 2
sizes — 131 072-point Scan-U vs the 10 M-raw-point Map-U of bench.py — map build, all four searches, one linearisation each."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
from elimaloc_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import reference_build as R  # noqa: E402

threads = min(os.cpu_count() or 1, 16)
R.set_threads(threads)
raw = synth.map_u(10_000_000, 100.0)
scan = synth.scan_u(131072, 40.0)
T = synth.se3([50.0, 50.0, 50.0], [0.01, -0.02, 0.3])
t = time.time()
om = O.VoxelHashMap(1.0, 30); om.AddPoints(raw); om.CalVoxelCovAll(); om.CalPointCovAll(0.4)
print(f"oracle map: {om.num_points()} points / {om.num_voxels()} voxels, {time.time() - t:.1f} s", flush=True)
t = time.time()
rm = R.VoxelHashMap(1.0, 30); rm.AddPoints(raw); rm.CalVoxelCovAll(); rm.CalPointCovAll(0.4)
print(f"reference-sources map: {rm.num_points()} points / {rm.num_voxels()} voxels, {time.time() - t:.1f} s ({threads} threads)", flush=True)
eo, er = om.export(), rm.export()
for k in ("keys", "counts", "pxyz", "vmean", "pmean"):
    print(f"  {k}: identical = {np.array_equal(eo[k], er[k])}")
for k in ("vcov", "pcov"):
    print(f"  {k}: max |diff| = {np.abs(eo[k] - er[k]).max():.3g}")
del eo, er
for m, name in enumerate(["P2P", "GICP", "VGICP", "AVGICP"]):
    co, to = O.correspondences(om, scan, T, m, 5.0)
    idx, tgt = R.search_pairs(rm, scan, T, m, 5.0)
    want_idx = np.repeat(np.arange(len(co)), co)
    want_tgt = np.concatenate([to[i, :c] for i, c in enumerate(co) if c])
    same = np.array_equal(idx, want_idx) and np.array_equal(tgt, want_tgt)
    cfg = O.make_config(icp_method=m, max_thread=threads)
    lo, lr = O.Registration().linearize(scan, om, T, cfg), R.Registration().linearize(scan, rm, T, cfg)
    print(f"{name}: {len(idx)} pairs, emission order and targets identical = {same}; JTJ rel diff = "
          f"{np.abs(lo['JTJ'] - lr['JTJ']).max() / np.abs(lr['JTJ']).max():.3g}, JTr rel diff = {np.abs(lo['JTr'] - lr['JTr']).max() / np.abs(lr['JTr']).max():.3g}", flush=True)

# ---- and the PRODUCT's host map builder (no GPU needed: device = -1) directly against the reference-sources build
import elimaloc_b200 as E  # noqa: E402

t = time.time()
pm = E.VoxelHashMap(1.0, 30, device=-1); pm.AddPoints(raw); pm.CalVoxelCovAll(); pm.CalPointCovAll(0.4)
print(f"product host builder (elimaloc_b200/csrc/host_map.cpp, device = -1): {time.time() - t:.1f} s", flush=True)
pe, re_ = pm.export(True, True), rm.export()
for k in ("keys", "counts", "pxyz"):
    print(f"  {k}: identical = {np.array_equal(pe[k], re_[k])}")
for k in ("vmean", "vcov", "pmean", "pcov"):
    print(f"  {k}: max |diff| = {np.abs(pe[k] - re_[k]).max():.3g}")
