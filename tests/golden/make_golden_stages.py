#!/usr/bin/env python
"""Generates tests/golden/ref_deskew.npz and ref_ekf_drive_ckf{0,1}.npz FROM THE REFERENCE'S OWN SOURCES (build container only).

    python tests/golden/make_golden_stages.py

ref_deskew.npz       one scan through the reference node's DeskewPointCloud (oracle/_ref/libref_node.so = pcm_matching.cpp compiled
                     unmodified against stand-in ROS / tf / PCL / Eigen headers): the inputs (points, per-point times), the tables
                     the node built from its IMU / odometry queues, and the undistorted cloud it produced.
ref_ekf_drive_*.npz  the arc drive of tests/test_ekf.py::drive through the reference's EkfAlgorithm (oracle/_ref/libref_ekf.so =
                     ekf_algorithm.cpp): GetCurrentState snapshots and the final members.
The oracle (CPU) and the CUDA kernels (GPU) are both checked against these files (tests/test_deskew.py, tests/test_ekf.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from elimaloc_b200 import _capi, ekf as pekf, synth  # noqa: E402
from oracle import reference_build as R  # noqa: E402


def main():
    if not R.sources_present():
        raise SystemExit("the golden vectors are generated from /root/reference; it is not here")
    R.build()
    import test_ekf
    import test_reference_build_node as N
    # ---- deskew
    node = R.PcmMatchingNode(synth.map_u(100, 5.0))
    N.feed(node, (2.0, 1.0, 0.5))
    n = 4096
    rng = np.random.default_rng(21)
    xyz = synth.scan_u(n, 60.0, seed=22)
    rel = np.sort(rng.random(n).astype(np.float32) * np.float32(0.1))
    ok, und, tab = node.deskew(N.T0 + 0.05, xyz, rel)
    assert ok
    k = tab["imu_pointer_cur"]
    for name in ("imu_time", "imu_rot_x", "imu_rot_y", "imu_rot_z"):
        tab[name][k + 1:] = 0.0  # entries beyond the pointer are uninitialised heap in the node
    np.savez_compressed(os.path.join(HERE, "ref_deskew.npz"), xyz=xyz, rel=rel, undistorted=und,
                        **{"tab_" + key: np.asarray(val) for key, val in tab.items()})
    print("deskew: pointer", k, "odom increments", tab["odom_incre"])
    # ---- EKF
    for ckf in (0, 1):
        f = R.EkfAlgorithm(pekf.make_ekf_config(use_complementary_filter=ckf), _capi.EkfState)
        snaps = np.array(test_ekf.drive(f))
        st = pekf.state_to_dict(f.s)
        np.savez_compressed(os.path.join(HERE, f"ref_ekf_drive_ckf{ckf}.npz"), snaps=snaps,
                            **{key: np.asarray(val) for key, val in st.items() if key not in ("reserved", "ego", "ego_prev_timestamp")})
        print("ekf ckf", ckf, "final pos", st["pos"])


if __name__ == "__main__":
    main()
