#!/usr/bin/env python
"""One-off, BASELINE config 5 at its full size (not part of the test suite): 131 072-point raw scans at 10 Hz with a 100 Hz IMU
against a 10 M-raw-point surface map in a 100 m box — deskew + AVGICP + time compensation + 27-state EKF in closed loop, the GPU
arm and the oracle arm of tests/pipeline_harness.py on identical streams.  Prints the pose-trajectory difference and the time per
scan of both arms.

    python tests/full_size_pipeline.py [n_scans=100] [arms=gpu,oracle]        (needs a B200 for the gpu arm)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pipeline_harness as H  # noqa: E402
from elimaloc_b200 import synth  # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 100
arms = (sys.argv[2] if len(sys.argv) > 2 else "gpu,oracle").split(",")
BOX, M_RAW, N_PTS = 100.0, 10_000_000, 131072


class BigWorld(H.World):
    """the harness' arc, with the radius grown to the map (40 m box -> 8 m radius; 100 m box -> 25 m radius, 6 m/s)"""

    def __init__(self):
        super().__init__(BOX, N_PTS, seed=7, radius=25.0, omega=0.25)


t = time.time()
raw = synth.map_s(M_RAW, BOX)
print(f"Map-S: {len(raw)} raw points in {time.time() - t:.1f} s", flush=True)
res = {}
for name in arms:
    t = time.time()
    arm = H.GpuArm(raw, {}) if name == "gpu" else H.OracleArm(raw, {})
    t_build = time.time() - t
    t = time.time()
    res[name] = H.run(arm, BigWorld(), n_scans)
    dt = time.time() - t
    w = BigWorld()
    err = [np.linalg.norm(T[:3, 3] - w.pose(ts)[:3, 3]) for ts, T in zip(res[name]["t"], res[name]["icp"])]
    print(f"{name}: map build {t_build:.1f} s, {n_scans} scans in {dt:.1f} s ({1e3 * dt / n_scans:.0f} ms per scan incl. the host glue), "
          f"success {int(res[name]['ok'].sum())}/{n_scans}, error to the true trajectory max {max(err):.3f} m, last {err[-1]:.3f} m", flush=True)
    del arm
if "gpu" in res and "oracle" in res:
    g, o = res["gpu"], res["oracle"]
    scale = np.abs(o["icp"][:, :3, 3]).max()
    print(f"GPU vs oracle over {n_scans} scans: ICP translation max |diff| {np.abs(g['icp'][:, :3, 3] - o['icp'][:, :3, 3]).max():.3g} m "
          f"(relative {np.abs(g['icp'][:, :3, 3] - o['icp'][:, :3, 3]).max() / scale:.3g}), rotation entries {np.abs(g['icp'][:, :3, :3] - o['icp'][:, :3, :3]).max():.3g}, "
          f"filter pose {np.abs(g['ego'] - o['ego']).max():.3g}, success flags equal {np.array_equal(g['ok'], o['ok'])}")
