"""Deskew (pcm_matching.cpp:467-824): oracle known-answer tests on the CPU, GPU-vs-oracle parity on the B200.
float32 path: the GPU evaluates sin/cos in double and rounds, the host libm evaluates in float; tolerance 2 float ulps
of the coordinate magnitude (documented in deskew.cu)."""
import numpy as np
import pytest

from elimaloc_b200 import synth
from oracle import oracle as O


def scene(n=20000, seed=3, imu_hz=200.0, gyro=(0.05, -0.03, 0.6), vel=(8.0, 0.5, 0.1)):
    rng = np.random.default_rng(seed)
    t_end = 1000.0
    span = 0.1
    t_cur = t_end - span
    stamps = np.arange(t_cur - 0.05, t_end + 0.05, 1.0 / imu_hz)
    g = np.tile(np.asarray(gyro), (len(stamps), 1)) + rng.normal(0, 0.002, (len(stamps), 3))
    start_pose = np.array([10.0, 20.0, 1.0, 0.01, -0.02, 0.4])
    end_pose = start_pose + np.array([vel[0] * span, vel[1] * span, vel[2] * span, 0.0, 0.0, gyro[2] * span])
    tab = O.deskew_tables(stamps, g, t_cur, t_end, start_pose, t_cur, end_pose, t_end)
    xyz = ((rng.random((n, 3), dtype=np.float32) * 2 - 1) * np.float32(80.0)).astype(np.float32)
    rel = np.sort(rng.random(n).astype(np.float32) * np.float32(span))
    return tab, xyz, rel


def test_tables_follow_the_reference_rules():
    tab, _, _ = scene()
    assert tab["imu_available"] and tab["odom_available"]
    k = tab["imu_pointer_cur"]
    assert 15 <= k <= 30  # 0.1 s scan +- 10 ms at 200 Hz
    assert tab["imu_rot_x"][0] == 0.0 and tab["imu_time"][0] >= tab["time_scan_cur"] - 0.01  # :538, :558-565
    assert tab["imu_time"][k] <= tab["time_scan_end"] + 0.01                                   # :556
    assert abs(tab["imu_rot_z"][k] - 0.6 * (tab["imu_time"][k] - tab["imu_time"][0])) < 2e-3
    Rb = synth.exp_so3([0, 0, 0.4]) @ synth.exp_so3([0, -0.02, 0]) @ synth.exp_so3([0.01, 0, 0])  # Rz Ry Rx (getTransformation)
    assert np.allclose(tab["odom_incre"], Rb.T @ np.array([0.8, 0.05, 0.01]), atol=1e-5)  # begin^-1 * end, :716


def test_deskew_point_known_answers():
    """identity at the scan end; Q3: the z translation uses the interpolated YAW integral (pcm_matching.cpp:804)"""
    tab, _, _ = scene(gyro=(0.0, 0.0, 0.5), vel=(5.0, 0.0, 0.0))
    k = tab["imu_pointer_cur"]
    t_last = tab["imu_time"][k] - tab["time_scan_cur"]
    p = np.array([[10.0, 0.0, 0.0]], np.float32)
    out = O.deskew_points(tab, p, np.array([t_last + 1.0], np.float32))  # later than every IMU stamp -> rot_cur == rot_end
    ratio = np.float32((np.float32(t_last + 1.0)) / (tab["time_scan_end"] - tab["time_scan_cur"]))
    exp_x = np.float32(10.0) + (ratio * tab["odom_incre"][0] - tab["odom_incre"][0])
    assert abs(out[0, 0] - exp_x) < 1e-5
    exp_z = np.float32(tab["imu_rot_z"][k]) - tab["odom_incre"][2]  # Q3
    assert abs(out[0, 2] - exp_z) < 1e-6
    # at scan start the point is rotated by -(total yaw) about z
    out0 = O.deskew_points(tab, p, np.array([0.0], np.float32))
    yaw = -np.float32(tab["imu_rot_z"][k]) + np.float32(tab["imu_rot_z"][0] if tab["imu_time"][0] > tab["time_scan_cur"] else 0)
    assert abs(np.arctan2(out0[0, 1] - (0 - tab["odom_incre"][1]), out0[0, 0] - (0 - tab["odom_incre"][0])) - np.arctan2(np.sin(yaw), np.cos(yaw))) < 0.02


def test_no_imu_passes_points_through():
    tab, xyz, rel = scene(n=100)
    tab = dict(tab, imu_available=False)
    assert np.array_equal(O.deskew_points(tab, xyz, rel), xyz)  # pcm_matching.cpp:781


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 1000, 131072])
def test_gpu_deskew_matches_oracle(n):
    import elimaloc_b200 as E
    tab, xyz, rel = scene(n=n, seed=11)
    reg = E.Registration(device=0)
    g = reg.DeskewPoints(xyz, rel, tab)
    o = O.deskew_points(tab, xyz, rel)
    tol = 2.0 * np.spacing(np.float32(128.0))
    assert np.abs(g - o).max() <= tol
    assert (g == o).mean() > 0.98  # almost every coordinate is bit-identical
    for variant in (dict(tab, odom_available=False), dict(tab, imu_available=False)):
        g = reg.DeskewPoints(xyz, rel, variant)
        o = O.deskew_points(variant, xyz, rel)
        assert np.abs(g - o).max() <= tol
