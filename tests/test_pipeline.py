"""BASELINE config 5: deskew + AVGICP + EKF closed loop, GPU arm vs oracle arm on identical synthetic streams."""
import numpy as np
import pytest

import pipeline_harness as H
from elimaloc_b200 import synth

EKF_KW = dict()


def truth_errors(world, res):
    err = []
    for t, T in zip(res["t"], res["icp"]):
        err.append(np.linalg.norm(T[:3, 3] - world.pose(t)[:3, 3]))
    return np.array(err)


def test_oracle_pipeline_tracks_the_truth():
    """the harness itself: the CPU arm alone must localise.  AVGICP with 1 m voxels is a coarse, biased estimator
    (every scan point is pulled towards up to 7 voxel means), so decimetre-level errors are its normal operating point;
    what the GPU test below checks is that the GPU arm reproduces the CPU arm, not absolute accuracy."""
    raw = synth.map_s(250_000, 30.0)
    arm = H.OracleArm(raw, EKF_KW)
    world = H.World(30.0, 4096, seed=7)
    res = H.run(arm, world, 16)
    assert res["ok"].all()
    e = truth_errors(world, res)
    assert e.max() < 0.6, e
    ekf_err = np.linalg.norm(res["ego"][-1][:3] - world.pose(res["t"][-1] + 0.03)[:3, 3])
    assert ekf_err < 1.5  # the filter free-runs on PCM updates only until 10 updates have passed (ekf_algorithm.cpp:189-194)


GPU_WORLD = dict(m_raw=400_000, box=40.0, n_points=8192, seed=7, n_scans=30)


def test_the_world_of_the_gpu_comparison_is_a_stable_one():
    """The closed loop (AVGICP on 1 m voxels feeding an EKF that starts with zero velocity) is not stable in every synthetic
    world: in some the reference algorithm itself loses track after ~20 scans (the reference's own two ROS nodes do exactly the
    same, tests/test_reference_build_node.py).  That loss of track is deterministic (a rounding-level perturbation of the
    sums stays at 1e-15 m), but a run whose ICP stops succeeding half-way exercises less of the path, and the GPU test below
    requires every scan to succeed.  Its world is one where the loop tracks with a comfortable margin for all 30 scans."""
    raw = synth.map_s(GPU_WORLD["m_raw"], GPU_WORLD["box"])
    world = H.World(GPU_WORLD["box"], GPU_WORLD["n_points"], seed=GPU_WORLD["seed"])
    res = H.run(H.OracleArm(raw, EKF_KW), world, GPU_WORLD["n_scans"])
    assert res["ok"].all()
    e = truth_errors(world, res)
    assert e.max() < 0.6 and e[-10:].max() < 0.35, e


@pytest.mark.gpu
def test_gpu_pipeline_matches_oracle_pipeline():
    """pose-trajectory diff GPU vs CPU reference port over 30 scans (3 s at 10 Hz, 100 Hz IMU).  Tolerance: 1e-4 relative
    on the pose (north star) — the translation is O(20 m), so 2e-3 m absolute; rotation entries 1e-4."""
    raw = synth.map_s(GPU_WORLD["m_raw"], GPU_WORLD["box"])
    ga, oa = H.GpuArm(raw, EKF_KW), H.OracleArm(raw, EKF_KW)
    rg = H.run(ga, H.World(GPU_WORLD["box"], GPU_WORLD["n_points"], seed=GPU_WORLD["seed"]), GPU_WORLD["n_scans"])
    ro = H.run(oa, H.World(GPU_WORLD["box"], GPU_WORLD["n_points"], seed=GPU_WORLD["seed"]), GPU_WORLD["n_scans"])
    assert np.array_equal(rg["ok"], ro["ok"]) and ro["ok"].all()
    scale = np.abs(ro["icp"][:, :3, 3]).max()
    assert np.abs(rg["icp"][:, :3, 3] - ro["icp"][:, :3, 3]).max() <= 1e-4 * scale
    assert np.abs(rg["icp"][:, :3, :3] - ro["icp"][:, :3, :3]).max() <= 1e-4
    assert np.abs(rg["ego"][:, :3] - ro["ego"][:, :3]).max() <= 1e-4 * scale
    assert np.abs(rg["ego"][:, 3:] - ro["ego"][:, 3:]).max() <= 1e-4
    assert np.abs(rg["fit"] - ro["fit"]).max() <= 1e-4
