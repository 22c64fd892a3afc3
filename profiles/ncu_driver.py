#!/usr/bin/env python
"""Small workload for ncu captures (run under ncu on the GPU box): python profiles/ncu_driver.py --method p2p [--iters 6] [--steps 3]
Builds the bench map (config 2 sizes unless --m-raw / --n-scan), then `steps` RunRegister calls of `iters` forced iterations on
different scans.  Launches per step: 1 (begin) + 2 per iteration (AVGICP: 1 per iteration)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402  (device buffers)
import elimaloc_b200 as E  # noqa: E402
from elimaloc_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--method", default="p2p")
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--n-scan", type=int, default=131072)
ap.add_argument("--m-raw", type=int, default=10_000_000)
ap.add_argument("--box", type=float, default=100.0)
a = ap.parse_args()
method = {"p2p": 0, "gicp": 1, "vgicp": 2, "avgicp": 3}[a.method]
gm = E.VoxelHashMap(1.0, 30, device=0)
gm.AddPoints(synth.map_u(a.m_raw, a.box))
if method >= 2:
    gm.CalVoxelCovAll()
if method == 1:
    gm.CalPointCovAll(0.4)
c = a.box / 2
T0 = synth.se3([c, c, c], np.deg2rad([1.0, -2.0, 30.0]))
reg = E.Registration(device=0)
cfg = E.RegistrationConfig(icp_method=method, max_iteration=a.iters, **synth.timing_knobs())
half = min(40.0, 0.4 * a.box)
for i in range(a.steps):
    d = torch.from_numpy(synth.scan_u(a.n_scan, half, seed=synth.SEED_SCAN + i)).cuda()
    reg.enqueue(d.data_ptr(), a.n_scan, gm, T0, cfg)
    print(i, reg.fetch()[4], flush=True)
