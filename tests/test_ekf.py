"""27-state EKF (ekf_algorithm.cpp): oracle known-answer tests on the CPU, GPU-vs-oracle parity on the B200.
fp64 path: tolerance 1e-9 relative on the state and on P (same algorithm, different summation order / libm)."""
import ctypes as C

import numpy as np
import pytest

from elimaloc_b200 import _capi, ekf as pekf, synth
from oracle import oracle as O

N = 27


def quat_wxyz(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    return np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])


def drive(f, n_imu=600, imu_dt=0.01, pcm_every=10, seed=1, init=True):
    """a car on a constant-twist arc: IMU at 100 Hz, a PCM pose measurement at 10 Hz; returns snapshots"""
    rng = np.random.default_rng(seed)
    v, wz = 8.0, 0.25
    t0 = 100.0
    snaps = []
    if init:
        f.RunGnssUpdate(pekf.make_measurement(t0, [0, 0, 0], quat_wxyz(np.eye(3)), np.eye(3) * 0.01, np.eye(3) * 1e-4, source=pekf.PCM_INIT))
    for k in range(n_imu):
        t = t0 + k * imu_dt
        yaw = wz * (t - t0)
        gyro = np.array([0.0, 0.0, wz]) + rng.normal(0, 1e-3, 3)
        acc = np.array([0.0, v * wz, 9.81]) + rng.normal(0, 1e-2, 3)  # centripetal + gravity, body frame
        f.RunPredictionImu(t, gyro, acc)
        if k % pcm_every == pcm_every - 1:
            pos = np.array([v / wz * np.sin(yaw), v / wz * (1 - np.cos(yaw)), 0.0]) + rng.normal(0, 0.02, 3)
            R = synth.exp_so3([0, 0, yaw + rng.normal(0, 1e-3)])
            f.RunGnssUpdate(pekf.make_measurement(t, pos, quat_wxyz(R), np.eye(3) * 0.0625, np.eye(3) * (0.25 * np.pi / 180) ** 2, source=pekf.PCM))
        if k % 50 == 49:
            snaps.append(f.GetCurrentState().copy())
    return snaps


def test_oracle_init_and_guards():
    cfg = pekf.make_ekf_config(ekf_init_x_m=1.0, ekf_init_yaw_deg=90.0)
    f = O.EkfAlgorithm(cfg, _capi.EkfState)
    P = np.array(f.s.P[:]).reshape(N, N)
    assert np.allclose(np.diag(P)[:15], 100.0) and np.allclose(np.diag(P)[15:18], 1e-4) and np.allclose(np.diag(P)[24:27], 1e-4)
    assert abs(f.s.rot[0] - np.cos(np.pi / 4)) < 1e-15 and abs(f.s.rot[3] - np.sin(np.pi / 4)) < 1e-15 and f.s.grav[2] == 9.81
    # first call only latches the timestamp (ekf_algorithm.cpp:182-187); not initialised -> no prediction (:198-208)
    assert not f.RunPredictionImu(10.0, [0, 0, 0], [0, 0, 9.81]) and f.s.prev_timestamp == 10.0
    assert not f.RunPredictionImu(10.01, [0, 0, 0], [0, 0, 9.81]) and f.s.predictions == 0
    # PCM_INIT hard reset (:324-349): flags, P block, and predictions frozen until > 10 PCM updates (:189-194, :357-364)
    f.RunGnssUpdate(pekf.make_measurement(10.02, [5, 6, 7], [1, 0, 0, 0], np.eye(3), np.eye(3), source=pekf.PCM_INIT))
    assert f.s.state_initialized and f.s.yaw_initialized and f.s.pcm_init_on_going and list(f.s.pos) == [5, 6, 7]
    assert not f.RunPredictionImu(10.03, [0, 0, 0], [0, 0, 9.81])
    for k in range(12):
        f.RunGnssUpdate(pekf.make_measurement(10.1 + 0.1 * k, [5, 6, 7], [1, 0, 0, 0], np.eye(3) * 0.01, np.eye(3) * 1e-4, source=pekf.PCM))
    assert not f.s.pcm_init_on_going and f.s.pcm_update_count == 12
    assert f.RunPredictionImu(11.5, [0, 0, 0], [0, 0, 9.81]) and f.s.predictions == 1


def test_oracle_update_is_the_textbook_kalman_step():
    """RunGnssUpdate (ekf_algorithm.cpp:366-428): K = P H^T (H P H^T + R)^-1, P <- P - K H P with H = [I6 0]"""
    cfg = pekf.make_ekf_config(use_complementary_filter=0)
    f = O.EkfAlgorithm(cfg, _capi.EkfState)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(N, N))
    P0 = A @ A.T / N + np.eye(N) * 0.1
    f.s.P[:] = list(P0.reshape(-1))
    f.s.state_initialized = 1
    R6 = np.diag([0.04, 0.04, 0.09, 1e-4, 1e-4, 4e-4])
    m = pekf.make_measurement(1.0, [0.3, -0.2, 0.1], quat_wxyz(synth.exp_so3([0, 0, 0.02])), R6[:3, :3], R6[3:, 3:], source=pekf.PCM)
    f.RunGnssUpdate(m)
    H = np.zeros((6, N)); H[:6, :6] = np.eye(6)
    K = P0 @ H.T @ np.linalg.inv(H @ P0 @ H.T + R6)
    np.testing.assert_allclose(np.array(f.s.P[:]).reshape(N, N), P0 - K @ H @ P0, rtol=1e-10, atol=1e-12)
    Y = np.array([0.3, -0.2, 0.1, 0, 0, 0.02])
    du = K @ Y
    np.testing.assert_allclose(list(f.s.pos), du[:3], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(list(f.s.vel), du[6:9], rtol=1e-10, atol=1e-13)


def test_oracle_filter_tracks_the_arc():
    f = O.EkfAlgorithm(pekf.make_ekf_config(), _capi.EkfState)
    snaps = drive(f)
    last = snaps[-1]
    t = last[0] - 100.0
    assert abs(last[1] - 8.0 / 0.25 * np.sin(0.25 * t)) < 0.3 and abs(last[6] - 0.25 * t) < 0.06  # x, yaw
    assert f.s.predictions > 300 and f.s.updates >= 50


@pytest.mark.gpu
@pytest.mark.parametrize("ckf", [1, 0])
def test_gpu_ekf_matches_oracle(ckf):
    import elimaloc_b200 as E
    cfg = pekf.make_ekf_config(use_complementary_filter=ckf)
    g = E.EkfAlgorithm(cfg, device=0)
    o = O.EkfAlgorithm(pekf.make_ekf_config(use_complementary_filter=ckf), _capi.EkfState)
    sg, so = drive(g), drive(o)
    for a, b in zip(sg, so):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-10)
    gs, os_ = pekf.state_to_dict(g.state()), pekf.state_to_dict(o.s)
    for k in ("pos", "rot", "vel", "gyro", "acc", "bg", "ba", "grav", "imu_rot"):
        np.testing.assert_allclose(gs[k], os_[k], rtol=1e-9, atol=1e-11, err_msg=k)
    np.testing.assert_allclose(gs["P"], os_["P"], rtol=1e-8, atol=1e-13)
    for k in ("state_initialized", "yaw_initialized", "rotation_stabilized", "state_stabilized", "pcm_init_on_going", "pcm_update_count",
              "predictions", "updates", "ckf_has_prev"):
        assert gs[k] == os_[k], k


@pytest.mark.gpu
def test_gpu_ekf_guards_and_uninitialised_path():
    """without PCM_INIT the filter never leaves the 'not initialised' branch: only timestamps move"""
    import elimaloc_b200 as E
    g = E.EkfAlgorithm(pekf.make_ekf_config(), device=0)
    o = O.EkfAlgorithm(pekf.make_ekf_config(), _capi.EkfState)
    for f in (g, o):
        drive(f, n_imu=60, init=False)
    gs, os_ = pekf.state_to_dict(g.state()), pekf.state_to_dict(o.s)
    assert gs["predictions"] == os_["predictions"] and gs["updates"] == os_["updates"] == 6
    np.testing.assert_allclose(gs["P"], os_["P"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(gs["pos"], os_["pos"], rtol=1e-9, atol=1e-12)


# ---- golden vectors generated by the REFERENCE's own EkfAlgorithm (tests/golden/make_golden_stages.py) ------------------------
def load_reference_ekf(ckf):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"ref_ekf_drive_ckf{ckf}.npz"))


def check_against_reference_golden(f, state, ckf, slack=1.0):
    g = load_reference_ekf(ckf)
    snaps = np.array(drive(f))
    np.testing.assert_allclose(snaps, g["snaps"], rtol=1e-9 * slack, atol=1e-10 * slack)
    s = pekf.state_to_dict(state())
    for k in ("pos", "rot", "vel", "gyro", "acc", "bg", "ba", "grav", "imu_rot"):
        np.testing.assert_allclose(s[k], g[k], rtol=1e-9 * slack, atol=1e-11 * slack, err_msg=k)
    np.testing.assert_allclose(s["P"], g["P"], rtol=1e-8 * slack, atol=1e-13 * slack)
    for k in ("state_initialized", "yaw_initialized", "rotation_stabilized", "state_stabilized", "pcm_init_on_going", "pcm_update_count"):
        assert s[k] == g[k], k


@pytest.mark.parametrize("ckf", [1, 0])
def test_oracle_ekf_matches_the_reference_golden(ckf):
    o = O.EkfAlgorithm(pekf.make_ekf_config(use_complementary_filter=ckf), _capi.EkfState)
    check_against_reference_golden(o, lambda: o.s, ckf)


@pytest.mark.gpu
@pytest.mark.parametrize("ckf", [1, 0])
def test_gpu_ekf_matches_the_reference_golden(ckf):
    import elimaloc_b200 as E
    g = E.EkfAlgorithm(pekf.make_ekf_config(use_complementary_filter=ckf), device=0)
    check_against_reference_golden(g, g.state, ckf, slack=3.0)  # GPU-vs-oracle and oracle-vs-reference rounding add up
