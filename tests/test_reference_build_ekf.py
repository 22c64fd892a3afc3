"""The pin of the oracle's EKF: oracle/ekf.cpp against oracle/_ref/libref_ekf.so — the REFERENCE's own ekf_algorithm.cpp with
ekf_algorithm.hpp, localization_functions.hpp and localization_struct.hpp, compiled unmodified from /root/reference against
stand-in Eigen / ROS headers (oracle/ref_build/stubs, node_stubs).  Runs without a GPU.

The filter's members are read after every call and compared member for member: state vector, both quaternions, the whole
27 x 27 covariance, timestamps, the initialised / stabilised flags, the PCM-init counter, return values, GetCurrentState.
Same algorithm, different summation order in the 27 x 27 products: 1e-9 relative."""
import numpy as np
import pytest

from elimaloc_b200 import _capi, ekf as pekf, synth
from oracle import oracle as O
from oracle import reference_build as R

pytestmark = pytest.mark.skipif(not R.ekf_available(), reason="neither /root/reference nor a prebuilt oracle/_ref/libref_ekf.so is here")

FLAGS = ("reset_for_init_prediction", "state_initialized", "yaw_initialized", "rotation_stabilized", "state_stabilized", "pcm_init_on_going",
         "pcm_update_count")
VECS = ("pos", "rot", "vel", "gyro", "acc", "bg", "ba", "grav", "imu_rot")


def pair(**kw):
    return O.EkfAlgorithm(pekf.make_ekf_config(**kw), _capi.EkfState), R.EkfAlgorithm(pekf.make_ekf_config(**kw), _capi.EkfState)


def assert_same_state(fo, fr, tol=1e-9, where=""):
    so, sr = pekf.state_to_dict(fo.s), pekf.state_to_dict(fr.s)
    for k in FLAGS:
        assert int(so[k]) == int(sr[k]), (where, k, so[k], sr[k])
    for k in ("prev_timestamp", "prev_gnss_timestamp"):
        assert so[k] == sr[k], (where, k)
    for k in VECS:
        assert np.abs(so[k] - sr[k]).max() <= tol * max(1.0, np.abs(sr[k]).max()), (where, k, so[k], sr[k])
    assert np.abs(so["P"] - sr["P"]).max() <= tol * np.abs(sr["P"]).max(), (where, "P")


def quat_wxyz(Rm):
    w = np.sqrt(max(0.0, 1 + Rm[0, 0] + Rm[1, 1] + Rm[2, 2])) / 2
    return np.array([w, (Rm[2, 1] - Rm[1, 2]) / (4 * w), (Rm[0, 2] - Rm[2, 0]) / (4 * w), (Rm[1, 0] - Rm[0, 1]) / (4 * w)])


def drive_both(fo, fr, n_imu=400, imu_dt=0.01, pcm_every=10, seed=1, init=True, tilt=(0.0, 0.0)):
    """a car on a constant-twist arc: IMU at 100 Hz, a PCM pose at 10 Hz; both filters get identical inputs"""
    rng = np.random.default_rng(seed)
    v, wz, t0 = 8.0, 0.25, 100.0
    if init:
        m = pekf.make_measurement(t0, [0, 0, 0], quat_wxyz(synth.exp_so3([tilt[0], tilt[1], 0.0])), np.eye(3) * 0.01, np.eye(3) * 1e-4, source=pekf.PCM_INIT)
        assert fo.RunGnssUpdate(m) == fr.RunGnssUpdate(m)
        assert_same_state(fo, fr, where="pcm_init")
    for k in range(n_imu):
        t = t0 + k * imu_dt
        yaw = wz * (t - t0)
        gyro = np.array([0.0, 0.0, wz]) + rng.normal(0, 1e-3, 3)
        acc = np.array([0.0, v * wz, 9.81]) + rng.normal(0, 1e-2, 3)
        assert fo.RunPredictionImu(t, gyro, acc) == fr.RunPredictionImu(t, gyro, acc), k
        if k % pcm_every == pcm_every - 1:
            pos = np.array([v / wz * np.sin(yaw), v / wz * (1 - np.cos(yaw)), 0.0]) + rng.normal(0, 0.02, 3)
            Rm = synth.exp_so3([tilt[0], tilt[1], yaw + rng.normal(0, 1e-3)])
            m = pekf.make_measurement(t, pos, quat_wxyz(Rm), np.eye(3) * 0.0625, np.eye(3) * (0.25 * np.pi / 180) ** 2, source=pekf.PCM)
            assert fo.RunGnssUpdate(m) == fr.RunGnssUpdate(m)
        if k % 7 == 6:
            assert_same_state(fo, fr, where=f"imu {k}")
        if k % 50 == 49:
            eo, er = fo.GetCurrentState(), fr.GetCurrentState()
            assert np.abs(eo - er).max() <= 1e-9 * max(1.0, np.abs(er).max()), k
    assert_same_state(fo, fr, where="end")


def test_init_state_and_covariance():
    fo, fr = pair(ekf_init_x_m=1.0, ekf_init_y_m=-2.0, ekf_init_z_m=0.5, ekf_init_roll_deg=3.0, ekf_init_pitch_deg=-4.0, ekf_init_yaw_deg=90.0)
    assert_same_state(fo, fr, tol=1e-15, where="init")
    P = pekf.state_to_dict(fr.s)["P"]
    assert np.allclose(np.diag(P)[:15], 100.0) and np.allclose(np.diag(P)[15:], 1e-4) and np.count_nonzero(P - np.diag(np.diag(P))) == 0


def test_guards_and_pcm_init_sequence():
    """first call latches the timestamp; no prediction before initialisation; PCM_INIT freezes predictions until more than 10
    PCM updates arrived; a repeated timestamp is not new data"""
    fo, fr = pair(use_complementary_filter=0)
    g, a = [0.0, 0.0, 0.0], [0.0, 0.0, 9.81]
    for t in (10.0, 10.01):
        assert fo.RunPredictionImu(t, g, a) == fr.RunPredictionImu(t, g, a) == False  # noqa: E712
        assert_same_state(fo, fr, where=f"pre-init {t}")
    m = pekf.make_measurement(10.02, [5, 6, 7], [1, 0, 0, 0], np.eye(3), np.eye(3), source=pekf.PCM_INIT)
    assert fo.RunGnssUpdate(m) and fr.RunGnssUpdate(m)
    assert_same_state(fo, fr, where="pcm_init")
    assert fr.s.pcm_init_on_going and fr.s.state_initialized and list(fr.s.pos) == [5, 6, 7]
    assert fo.RunPredictionImu(10.03, g, a) == fr.RunPredictionImu(10.03, g, a) == False  # noqa: E712
    for k in range(12):
        m = pekf.make_measurement(10.1 + 0.1 * k, [5, 6, 7], [1, 0, 0, 0], np.eye(3) * 0.01, np.eye(3) * 1e-4, source=pekf.PCM)
        assert fo.RunGnssUpdate(m) == fr.RunGnssUpdate(m)
        assert_same_state(fo, fr, where=f"pcm {k}")
    assert not fr.s.pcm_init_on_going and fr.s.pcm_update_count == 12
    assert fo.RunPredictionImu(11.5, g, a) == fr.RunPredictionImu(11.5, g, a) == True  # noqa: E712
    assert fo.RunPredictionImu(11.5, g, a) == fr.RunPredictionImu(11.5, g, a) == False  # noqa: E712
    assert_same_state(fo, fr, where="after first prediction")


@pytest.mark.parametrize("ckf", [0, 1])
@pytest.mark.parametrize("gravity", [0, 1])
def test_arc_drive_matches(ckf, gravity):
    fo, fr = pair(use_complementary_filter=ckf, imu_estimate_gravity=gravity)
    drive_both(fo, fr, seed=3 + ckf)
    assert fr.s.state_initialized and fr.s.rotation_stabilized


def test_tilted_vehicle_and_sparse_measurements():
    fo, fr = pair(use_complementary_filter=1)
    drive_both(fo, fr, n_imu=300, pcm_every=25, seed=9, tilt=(0.05, -0.08))


def test_uninitialised_filter_runs_the_complementary_filter_once_yaw_is_known():
    """before the state is initialised RunPredictionImu only feeds the complementary filter, and only when yaw is initialised
    (ekf_algorithm.cpp:198-208): drive PCM updates without PCM_INIT until the yaw covariance shrinks"""
    fo, fr = pair(use_complementary_filter=1)
    rng = np.random.default_rng(4)
    for k in range(60):
        t = 50.0 + 0.01 * k
        gyro, acc = rng.normal(0, 1e-3, 3), np.array([0.1, -0.2, 9.8]) + rng.normal(0, 1e-2, 3)
        assert fo.RunPredictionImu(t, gyro, acc) == fr.RunPredictionImu(t, gyro, acc)
        if k % 5 == 4:
            m = pekf.make_measurement(t, [1.0, 2.0, 0.0], [1, 0, 0, 0], np.eye(3) * 0.04, np.eye(3) * 1e-4, source=pekf.PCM)
            assert fo.RunGnssUpdate(m) == fr.RunGnssUpdate(m)
        assert_same_state(fo, fr, where=f"step {k}")
    assert fr.s.yaw_initialized


@pytest.mark.parametrize("seed", range(8))
def test_randomised_event_streams(seed):
    """irregular and repeated IMU stamps, a backwards stamp, measurements at random times with random covariances, large attitude
    changes (yaw wrapping through +-pi, pitch pushed towards the gimbal-lock branch of RotToVec), re-initialisation by a second
    PCM_INIT in the middle of the stream — every member compared after every event"""
    rng = np.random.default_rng(500 + seed)
    ckf, grav = int(rng.integers(0, 2)), int(rng.integers(0, 2))
    fo, fr = pair(use_complementary_filter=ckf, imu_estimate_gravity=grav, ekf_init_yaw_deg=float(rng.uniform(-180, 180)),
                  ekf_init_pitch_deg=float(rng.uniform(-20, 20)))
    t = 10.0
    att = np.array([0.0, rng.uniform(-1.2, 1.2), rng.uniform(-3.1, 3.1)])   # roll, pitch, yaw of the "truth"
    pos = rng.normal(0, 5, 3)
    m0 = pekf.make_measurement(t, pos, quat_wxyz(synth.exp_so3([0, 0, att[2]]) @ synth.exp_so3([0, att[1], 0])), np.eye(3) * 1e-9, np.eye(3) * 1e-9,
                               source=pekf.PCM_INIT)
    assert fo.RunGnssUpdate(m0) == fr.RunGnssUpdate(m0)
    for k in range(260):
        ev = rng.random()
        if ev < 0.70:                                            # IMU sample
            dt = float(rng.choice([0.0, 1e-7, 0.005, 0.01, 0.01, 0.01, 0.05, -0.02]))
            t += dt
            gyro = rng.normal(0, 0.5, 3) * (3.0 if k % 40 < 5 else 1.0)
            acc = np.array([0.0, 0.0, 9.81]) + rng.normal(0, 1.0, 3)
            assert fo.RunPredictionImu(t, gyro, acc) == fr.RunPredictionImu(t, gyro, acc), k
            att += gyro * max(dt, 0.0)
        elif ev < 0.97:                                          # PCM pose
            att[2] = (att[2] + np.pi) % (2 * np.pi) - np.pi
            Rm = synth.exp_so3([0, 0, att[2]]) @ synth.exp_so3([0, np.clip(att[1], -1.5, 1.5), 0]) @ synth.exp_so3([att[0], 0, 0])
            pc = np.diag(rng.uniform(1e-4, 0.5, 3))
            rc = np.diag(rng.uniform(1e-6, 1e-2, 3))
            m = pekf.make_measurement(t - float(rng.uniform(0, 0.05)), pos + rng.normal(0, 0.1, 3), quat_wxyz(Rm), pc, rc, source=pekf.PCM)
            assert fo.RunGnssUpdate(m) == fr.RunGnssUpdate(m), k
        else:                                                    # forced re-initialisation
            m = pekf.make_measurement(t, pos, quat_wxyz(synth.exp_so3([0, 0, att[2]])), np.eye(3) * 1e-9, np.eye(3) * 1e-9, source=pekf.PCM_INIT)
            assert fo.RunGnssUpdate(m) == fr.RunGnssUpdate(m), k
        assert_same_state(fo, fr, tol=1e-8, where=f"seed {seed} event {k}")
        if k % 20 == 19:
            eo, er = fo.GetCurrentState(), fr.GetCurrentState()
            assert np.abs(eo[:25] - er[:25]).max() <= 1e-8 * max(1.0, np.abs(er[:25]).max()), k
    so = pekf.state_to_dict(fo.s)
    assert np.isfinite(so["P"]).all() and np.isfinite(so["pos"]).all() and so["updates"] > 30
