#!/usr/bin/env python
"""One-off (build container; ~2 minutes): BASELINE config 5 at its full size — 131 072-point scans, 10 M-raw-point Map-S in a
100 m box — on the reference's own two ROS nodes (oracle/_ref/libref_node.so + libref_ekfnode.so) against tests/pipeline_harness.run
with the oracle arm.  Both sides drop scan points that share a 1 mm voxel (the node always down-samples; 0.001 m makes it a no-op
except for exact-duplicate draws of the synthetic scan).

    python tests/full_size_reference_nodes.py [n_scans=8]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pipeline_harness as H  # noqa: E402
import test_reference_build_node as T  # noqa: E402
from elimaloc_b200 import ekf as pekf, synth  # noqa: E402


class BigWorld(H.World):
    def __init__(self):
        super().__init__(100.0, 131072, seed=7, radius=25.0, omega=0.25)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
raw = synth.map_s(10_000_000, 100.0)
t = time.time()
ref = T.run_reference_nodes(raw, BigWorld(), n, pekf.make_ekf_config(), input_voxel_ds_m=0.001)
tr = time.time() - t
t = time.time()
har = H.run(H.OracleArm(raw, {}), BigWorld(), n, input_voxel_ds_m=0.001)
th = time.time() - t
print(f"BASELINE config 5 at full size, {n} scans: the reference's own two ROS nodes ({tr:.0f} s incl. map build) vs the harness with the oracle arm ({th:.0f} s)")
print(f"  success flags equal: {np.array_equal(ref['ok'], har['ok'])} ({int(ref['ok'].sum())}/{n}); ICP pose max |diff| {np.nanmax(np.abs(ref['icp'] - har['icp'])):.3g}; "
      f"filter pose max |diff| {np.abs(ref['ego'] - har['ego']).max():.3g}")
