#!/usr/bin/env python
"""A/B harness inside ONE process (one map build): ELM_WARM_MODE (pair | single | async) and ELM_ASYNC_GRID (blocks of the concurrent
refresh kernel) are read when a registration handle is created, so every variant gets its own handle; a compile-time variant of the
library is selected per process with ELIMALOC_B200_LIB (profiles/build_variant.sh).

  python profiles/ab_chain.py [--methods p2p,gicp,vgicp,avgicp] [--sizes 131072,16384] [--grids 80,128] [--modes async,pair] [--steps 30]

(The name is historical: calls Q and R of round 2 used it to measure the "chained" warm iterations, ELM_WARM_MODE=chain, which existed in
commits 802062d..e950929 and were removed — DESIGN.md, table of things that did not pay.  An unknown mode falls back to the default.)
Timing as bench.py's `value`: CUDA events on the launch stream around K enqueued RunRegister calls of 20 forced iterations, a different
scan every step, after 3 warm-up steps, the faster of two timed regions.  One line per variant: iterations/s, us per iteration, the pose
difference to the established mode on the same summation grid (must be 0.0), a checksum of pose + fitness (equal across builds whose sums
are bit-identical) and one JSON line with everything at the end."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import elimaloc_b200 as E  # noqa: E402
from elimaloc_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--methods", default="p2p,gicp")
ap.add_argument("--sizes", default="131072,16384")
ap.add_argument("--grids", default="128,96,80,64,48,32")
ap.add_argument("--modes", default="async,chain")
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--m-raw", type=int, default=10_000_000)
args = ap.parse_args()

METHOD = {"p2p": E.P2P, "gicp": E.GICP, "vgicp": E.VGICP, "avgicp": E.AVGICP}
methods = [m for m in args.methods.split(",") if m]
gm = E.VoxelHashMap(1.0, 30, device=0)
gm.AddPoints(synth.map_u(args.m_raw, 100.0))
if "gicp" in methods:
    gm.CalPointCovAll(0.4)
if "vgicp" in methods or "avgicp" in methods:
    gm.CalVoxelCovAll()
T0 = synth.se3([50, 50, 50], np.deg2rad([1.0, -2.0, 30.0]))
stream = torch.cuda.Stream(device=0)
torch.cuda.set_stream(stream)
out = []
for mname in methods:
    for n in [int(v) for v in args.sizes.split(",")]:
        d_scans = [torch.from_numpy(synth.scan_u(n, 40.0, seed=synth.SEED_SCAN + i)).cuda() for i in range(4)]
        cfg = E.RegistrationConfig(icp_method=METHOD[mname], max_iteration=args.iters, **synth.timing_knobs())
        ref_pose = {}
        broken = set()
        for mode in args.modes.split(","):  # (the established mode first: a broken experimental mode cannot cost the other numbers)
            for gidx, grid in enumerate(int(v) for v in args.grids.split(",")):
                if mode in broken:
                    continue
                if mname in ("vgicp", "avgicp") and (mode != args.modes.split(",")[0] or gidx > 0):
                    continue  # (no warm iterations: one line per size)
                os.environ["ELM_WARM_MODE"] = mode
                os.environ["ELM_ASYNC_GRID"] = str(grid)
                reg = E.Registration(device=0, stream=stream.cuda_stream)
                try:
                    reg.enqueue(d_scans[0].data_ptr(), n, gm, T0, cfg)  # one call alone first: a timeout inside the kernels surfaces here
                    reg.fetch()
                    for i in range(1, 3):
                        reg.enqueue(d_scans[i % 4].data_ptr(), n, gm, T0, cfg)
                    reg.fetch()
                    best = None
                    for rep in range(2):  # two timed regions, the faster one counts (the first one still warms the clocks)
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        torch.cuda.synchronize()
                        e0.record(stream)
                        for i in range(args.steps):
                            reg.enqueue(d_scans[i % 4].data_ptr(), n, gm, T0, cfg)
                        e1.record(stream)
                        torch.cuda.synchronize()
                        T, ok, fit, cov, it = reg.fetch()
                        ms = e0.elapsed_time(e1)
                        best = ms if best is None else min(best, ms)
                except Exception as exc:  # noqa: BLE001
                    print(f"{mname:5s} n={n:7d} {mode:6s} grid={grid:4d}  FAILED: {exc}", flush=True)
                    out.append(dict(method=mname, n=n, mode=mode, grid=grid, failed=str(exc), lib=os.environ.get("ELIMALOC_B200_LIB", "default")))
                    broken.add(mode)
                    del reg
                    continue
                us_it = 1e3 * best / (args.steps * args.iters)
                d_same = None
                if grid in ref_pose:
                    d_same = float(np.abs(T - ref_pose[grid]).max())
                else:
                    ref_pose[grid] = T
                d_first = float(np.abs(T - next(iter(ref_pose.values()))).max())
                row = dict(method=mname, n=n, mode=mode, grid=grid, it_per_s=1e6 / us_it, us_per_iteration=us_it, iterations=it, success=bool(ok),
                           pose_diff_same_grid=d_same, pose_diff_first=d_first, lib=os.environ.get("ELIMALOC_B200_LIB", "default"),
                           pose_sha=hashlib.sha1(np.ascontiguousarray(T).tobytes() + np.float64(fit).tobytes()).hexdigest()[:12])
                out.append(row)
                print(f"{mname:5s} n={n:7d} {mode:6s} grid={grid:4d}  {row['it_per_s']:9.0f} it/s  {us_it:7.2f} us/it  iters={it} ok={ok} "
                      f"dpose(same grid)={d_same} dpose(first)={d_first:.2e} sha={row['pose_sha']}", flush=True)
                del reg
print(json.dumps(out))
