"""The pin of the stages around RunRegister: the REFERENCE's own ROS node class PcmMatching (pcm_matching.cpp, compiled unmodified
from /root/reference against stand-in ROS / tf / PCL / boost / Eigen headers into oracle/_ref/libref_node.so) with the test as
the middleware, against the oracle's restatements and the numpy glue of tests/pipeline_harness.py.  Runs without a GPU.

  FilterPointsByDistance   pcm_matching.cpp:451-465     vs oracle.scan_preprocess                      identical indices
  ImuDeskewInfo / OdomDeskewInfo  :533-729              vs oracle.deskew_tables                        tables 1e-15 / float increments
                                                        vs the PRODUCT's builders (elm_deskew_build_tables, a18)       same tolerances
  DeskewPoint over a scan  :499-511, 780-824            vs oracle.deskew_points                        bit-equal on the same tables
  GetInterpolatedPose      :933-1045                    vs pipeline_harness.get_interpolated_pose      float32 rounding
  PublishPcmOdom covariance :1047-1101                  vs oracle.shape_pcm_covariance + the product's host function   1e-12
  CallbackPointCloud end to end :198-324                vs the same chain composed from the oracle's pieces             1e-6
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pipeline_harness as H  # noqa: E402
from elimaloc_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import reference_build as R  # noqa: E402

pytestmark = pytest.mark.skipif(not R.node_available(), reason="neither /root/reference nor a prebuilt oracle/_ref/libref_node.so is here")

T0 = 100.0
LIDAR_XYZ, LIDAR_RPY_DEG = (0.0961, -0.1338, 0.3032), (-1.26, -0.876, 0.287)  # the reference's config/calibration.ini


def tf_quat_from_rpy(r, p, y):
    """tf::Quaternion::setRPY -> (w, x, y, z)"""
    cy, sy, cp, sp, cr, sr = np.cos(y / 2), np.sin(y / 2), np.cos(p / 2), np.sin(p / 2), np.cos(r / 2), np.sin(r / 2)
    return np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy])


def tf_rpy_from_quat(q):
    """tf::Matrix3x3(q).getRPY (non-degenerate branch)"""
    Rm = H.quat_to_R(np.asarray(q, dtype=np.float64))
    pitch = -np.arcsin(Rm[2, 0])
    return np.arctan2(Rm[2, 1] / np.cos(pitch), Rm[2, 2] / np.cos(pitch)), pitch, np.arctan2(Rm[1, 0] / np.cos(pitch), Rm[0, 0] / np.cos(pitch))


def ego_pose(t, centre):
    yaw = 0.3 * (t - T0)
    return np.array([centre[0] + 3 * np.sin(yaw), centre[1] + 3 * (1 - np.cos(yaw)), centre[2]]), (0.01, -0.02, yaw)


def feed(node, centre, k0=-5, k1=30, imu=True, odom=True):
    """100 Hz odometry and IMU; returns the odometry queue in the harness' format and the IMU arrays"""
    deq, stamps, gyro = [], [], []
    for k in range(k0, k1):
        t = T0 + 0.01 * k
        p, rpy = ego_pose(t, centre)
        q = tf_quat_from_rpy(*rpy)
        if odom:
            node.odom(t, p, q, lin=(0.9, 0.0, 0.0), ang=(0.0, 0.0, 0.3))
            deq.append(dict(t=t, pos=p, quat=q, vel_local=np.array([0.9, 0.0, 0.0]), rate=np.array([0.0, 0.0, 0.3])))
        if imu:
            g = [0.01 + 1e-3 * np.sin(k), -0.02, 0.3 + 1e-3 * np.cos(k)]
            node.imu(t + 0.003, g, [0.0, 0.0, 9.81])
            stamps.append(t + 0.003)
            gyro.append(g)
    return deq, np.array(stamps), np.array(gyro)


def oracle_tables(deq, stamps, gyro, t_cur, t_end):
    """the odometry selection of OdomDeskewInfo (first message not older than scan start / scan end) + the oracle's tables"""
    s = next(o for o in deq if not o["t"] < t_cur)
    e = next(o for o in deq if not o["t"] < t_end)
    return O.deskew_tables(stamps, gyro, t_cur, t_end, list(s["pos"]) + list(tf_rpy_from_quat(s["quat"])), s["t"],
                           list(e["pos"]) + list(tf_rpy_from_quat(e["quat"])), e["t"])


def product_tables(deq, stamps, gyro, t_cur, t_end):
    """the product's own ImuDeskewInfo / OdomDeskewInfo (elimaloc_b200/csrc/deskew_tables.cpp) on the same queues"""
    import elimaloc_b200 as E
    wxyz = np.array([o["quat"] for o in deq]).reshape(-1, 4)
    q = E.Queues(stamps, np.asarray(gyro).reshape(-1, 3), [o["t"] for o in deq], np.array([o["pos"] for o in deq]).reshape(-1, 3),
                 wxyz[:, [1, 2, 3, 0]], np.array([o["vel_local"] for o in deq]).reshape(-1, 3), np.array([o["rate"] for o in deq]).reshape(-1, 3))
    return E.build_deskew_tables(q, t_cur, t_end)


def assert_product_tables_equal_the_nodes(pt, tab, ok, incre_tol):
    assert bool(pt["odom_available"]) == bool(tab["odom_available"])
    if tab["odom_available"] or tab["imu_available"]:  # (the node leaves the IMU members untouched when it returns early)
        assert bool(pt["imu_available"]) == bool(tab["imu_available"])
    assert ok == (pt["imu_available"] and pt["odom_available"])
    if ok:
        k = tab["imu_pointer_cur"]
        assert pt["imu_pointer_cur"] == k
        for name in ("imu_time", "imu_rot_x", "imu_rot_y", "imu_rot_z"):
            assert np.abs(pt[name][:k + 1] - tab[name][:k + 1]).max() < 1e-14, name
        assert np.abs(pt["odom_incre"] - tab["odom_incre"]).max() < incre_tol


@pytest.fixture(scope="module")
def raw_map():
    return synth.map_u(40_000, 16.0, origin=-3.0)


def test_node_builds_the_same_map(raw_map):
    node = R.PcmMatchingNode(raw_map, icp_method=1)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw_map)
    assert node.map_points() == om.num_points()


def test_distance_filter(raw_map):
    node = R.PcmMatchingNode(raw_map[:100], input_max_dist=37.5)
    xyz = synth.scan_u(20_000, 40.0, seed=3)
    xyz[:5] = [[37.5, 0, 0], [0, -37.5, 0], [22.5, 30.0, 0], [37.500004, 0, 0], [0, 0, 0]]  # on and just beyond the sphere
    got = node.filter_by_distance(xyz)
    assert np.array_equal(got, O.scan_preprocess(xyz, 37.5, 0.0))
    assert 0 in got and 1 in got and 3 not in got and 0 < len(got) < len(xyz)


@pytest.mark.parametrize("scan_time_end", [0, 1])
def test_deskew_tables_and_points(raw_map, scan_time_end):
    node = R.PcmMatchingNode(raw_map[:100], scan_time_end=scan_time_end)
    deq, stamps, gyro = feed(node, (2.0, 1.0, 0.5))
    n = 4000
    rng = np.random.default_rng(0)
    xyz = synth.scan_u(n, 30.0, seed=1)
    rel = np.sort(rng.random(n).astype(np.float32) * np.float32(0.1))
    stamp = T0 + 0.05
    if scan_time_end:  # "last point is time 0": negative per-point times, stamp = scan end
        rel_in, stamp_in = (rel - rel[-1]).astype(np.float32), stamp + float(rel[-1])
        t_end = stamp_in
        t_cur = t_end + float(rel_in[0])
        rel_eff = (rel_in - rel_in[0]).astype(np.float32)
    else:
        rel_in, stamp_in, t_cur, t_end, rel_eff = rel, stamp, stamp, stamp + float(rel[-1]), rel
    ok, und, tab = node.deskew(stamp_in, xyz, rel_in)
    assert ok and tab["imu_available"] and tab["odom_available"]
    assert tab["time_scan_cur"] == t_cur and tab["time_scan_end"] == t_end
    # the per-point transform on the node's own tables: bit-equal
    assert np.array_equal(und, O.deskew_points(tab, xyz, rel_eff))
    # the tables themselves
    ot = oracle_tables(deq, stamps, gyro, t_cur, t_end)
    k = tab["imu_pointer_cur"]
    assert ot["imu_pointer_cur"] == k > 3
    for name in ("imu_time", "imu_rot_x", "imu_rot_y", "imu_rot_z"):
        assert np.abs(ot[name][:k + 1] - tab[name][:k + 1]).max() < 1e-15, name
    assert np.abs(ot["odom_incre"] - tab["odom_incre"]).max() < 2e-7  # float32; rpy goes through a quaternion on the node's side
    assert np.abs(und - O.deskew_points(ot, xyz, rel_eff)).max() < 2e-5
    assert_product_tables_equal_the_nodes(product_tables(deq, stamps, gyro, t_cur, t_end), tab, ok, 2e-7)


def test_deskew_needs_imu_and_odometry(raw_map):
    xyz = synth.scan_u(100, 10.0, seed=2)
    rel = np.linspace(0, 0.1, 100).astype(np.float32)
    node = R.PcmMatchingNode(raw_map[:100])
    feed(node, (2.0, 1.0, 0.5), imu=False)
    assert not node.deskew(T0 + 0.05, xyz, rel)[0]
    node = R.PcmMatchingNode(raw_map[:100])
    feed(node, (2.0, 1.0, 0.5), odom=False)
    assert not node.deskew(T0 + 0.05, xyz, rel)[0]
    node = R.PcmMatchingNode(raw_map[:100])
    feed(node, (2.0, 1.0, 0.5), k0=10)  # the first odometry message is newer than the scan start: no synced pose
    assert not node.deskew(T0 + 0.05, xyz, rel)[0]


def test_interpolated_pose(raw_map):
    node = R.PcmMatchingNode(raw_map[:100])
    deq, _, _ = feed(node, (2.0, 1.0, 0.5))
    for t in (T0 + 0.1234, T0 + 0.05, deq[-1]["t"], deq[-1]["t"] + 0.02, deq[-1]["t"] + 0.5):  # between, on, last, beyond the queue
        ok, T = node.interpolated_pose(t)
        want = H.get_interpolated_pose(deq, t)
        assert ok and np.abs(T - want).max() < 1e-6, t
    assert not node.interpolated_pose(deq[0]["t"] - 1.0)[0] and H.get_interpolated_pose(deq, deq[0]["t"] - 1.0) is None


def test_covariance_shaping(raw_map):
    import elimaloc_b200 as E
    node = R.PcmMatchingNode(raw_map[:100])
    rng = np.random.default_rng(5)
    for trial in range(12):
        A = rng.normal(size=(6, 6)) * (1e-6 if trial % 3 == 0 else 1e-2)
        local_cov = A @ A.T + np.diag(rng.random(6) * (1e-12 if trial % 4 == 0 else 1e-4))
        pose = synth.se3(rng.normal(0, 5, 3), rng.normal(0, 0.5, 3))
        std = [0.1, 0.25, 0.4, 1.3][trial % 4]
        got = node.shape_covariance(pose, local_cov, std)
        want = O.shape_pcm_covariance(pose[:3, :3], local_cov, std)
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
        assert np.abs(E.shape_pcm_covariance(pose[:3, :3], local_cov, std) - got).max() <= 1e-12 * max(1.0, np.abs(got).max())


@pytest.mark.parametrize("method", [O.P2P, O.GICP, O.VGICP])
def test_lidar_callback_end_to_end(raw_map, method):
    """one lidar message through the node's CallbackPointCloud against the same chain composed from the oracle's pieces:
    filter -> deskew -> pose interpolation -> first-in-voxel down-sampling -> RunRegister -> ego pose -> covariance"""
    node = R.PcmMatchingNode(raw_map, lidar_xyz=LIDAR_XYZ, lidar_rpy_deg=LIDAR_RPY_DEG, icp_method=method, input_max_dist=60.0)
    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw_map)
    om.CalVoxelCovAll()
    om.CalPointCovAll(0.4)
    centre = (5.0, 4.0, 5.0)
    deq, stamps, gyro = feed(node, centre)
    tf = np.eye(4)
    tf[:3, :3] = H.rpy_to_R(*np.deg2rad(LIDAR_RPY_DEG))
    tf[:3, 3] = LIDAR_XYZ
    # a scan of map points seen from the lidar at scan end, with per-point times
    stamp, n = T0 + 0.05, 3000
    rng = np.random.default_rng(7)
    rel = np.sort(rng.random(n).astype(np.float32) * np.float32(0.1))
    p_end, rpy_end = ego_pose(stamp + float(rel[-1]), centre)
    T_lidar = np.eye(4)
    T_lidar[:3, :3] = H.rpy_to_R(*rpy_end)
    T_lidar[:3, 3] = p_end
    T_lidar = T_lidar @ tf
    scan = synth.scan_m(om.export()["pxyz"], n, T_lidar, noise=0.02, seed=11)
    scan[:3] *= np.float32(40.0)  # three points beyond input_max_dist
    out = node.cloud(stamp, scan, rel)
    assert out is not None, "the node did not publish a pose"
    # --- the same chain from the oracle's pieces
    keep = O.scan_preprocess(scan, 60.0, 0.0)
    assert len(keep) == n - 3
    xyz, rt = scan[keep], rel[keep]
    t_cur, t_end = stamp, stamp + float(rt[-1])
    und = O.deskew_points(oracle_tables(deq, stamps, gyro, t_cur, t_end), xyz, rt)
    sync = H.get_interpolated_pose(deq, t_end).astype(np.float64) @ tf
    ds = und[O.scan_preprocess(und, 0.0, 1.5)]
    r = O.Registration().RunRegister(ds, om, sync, O.make_config(icp_method=method))
    assert r["is_success"]
    ego = r["pose"] @ np.linalg.inv(tf)
    assert out["stamp"] == t_end
    assert len(out["registered_world"]) == len(ds)
    assert np.abs(out["pos"] - ego[:3, 3]).max() < 2e-6
    q = H.R_to_quat_wxyz(ego[:3, :3])
    assert min(np.abs(out["quat_wxyz"] - q).max(), np.abs(out["quat_wxyz"] + q).max()) < 1e-7
    want_cov = O.shape_pcm_covariance(ego[:3, :3], r["local_cov"], r["fitness_score"])
    assert np.abs(out["cov"] - want_cov).max() <= 1e-5 * np.abs(want_cov).max()
    # and the pose is sensible: the node localised the scan (the synthetic scan is not motion-distorted, so the deskew itself
    # moves points by up to the 9 cm the vehicle travels during the sweep)
    assert np.linalg.norm(out["pos"] - p_end) < 0.15


# ------------------------------------------------------------------------------------------------ the closed loop
class PinnedSpanWorld(H.World):
    """scans whose first / last per-point times are exactly 0 and float32(0.1): the node derives the scan end from the last
    point's time, the harness takes it as given"""

    def scan(self, stored, t_end, span=0.1):
        xyz, tt = super().scan(stored, t_end, span)
        tt[0], tt[-1] = np.float32(0.0), np.float32(span)
        return xyz, tt


def run_reference_nodes(raw_map, world, n_scans, ekf_cfg, imu_dt=0.01, latency=0.03, scan_offset=0.0, input_voxel_ds_m=0.001):
    """the same sensor stream as pipeline_harness.run, through the reference's own two ROS nodes with this function as the
    middleware: IMU -> both nodes, EKF odometry -> PCM node, lidar -> PCM node, PCM odometry -> EKF node"""
    pcm = R.PcmMatchingNode(raw_map, icp_method=O.AVGICP, max_fitness_score=2.0, max_thread=4, input_voxel_ds_m=input_voxel_ds_m, input_max_dist=1000.0)
    ekf = R.EkfLocalizationNode(ekf_cfg)
    stored = None
    t0 = world.t0
    T0 = world.pose(t0)
    ekf.pcm_init_odom(t0, T0[:3, 3], H.R_to_quat_wxyz(T0[:3, :3]))
    out = dict(icp=[], ego=[], ok=[], t=[])
    k_imu = 0

    def step_imu(until):
        nonlocal k_imu
        while t0 + k_imu * imu_dt <= until + 1e-9:
            t = t0 + k_imu * imu_dt
            g, a = world.imu(t)
            od = ekf.imu(t, g, a)
            pcm.imu(t, g, a)
            if od is not None:
                pcm.odom(od["t"], od["pos"], od["quat"], od["vel_local"], od["rate"])
            k_imu += 1

    om = O.VoxelHashMap(1.0, 30)
    om.AddPoints(raw_map)
    stored = om.export()["pxyz"]
    for s in range(n_scans):
        t_end = t0 + 0.1 * (s + 1) + scan_offset
        step_imu(t_end + latency)
        xyz, rel = world.scan(stored, t_end)
        res = pcm.cloud(t_end - 0.1, xyz, rel)
        out["ok"].append(res is not None)
        out["t"].append(t_end)
        if res is not None:
            T = np.eye(4)
            T[:3, :3] = H.quat_to_R(res["quat_wxyz"])
            T[:3, 3] = res["pos"]
            out["icp"].append(T)
            ekf.pcm_odom(res["stamp"], res["pos"], res["quat_wxyz"], res["cov"])
        else:
            out["icp"].append(np.full((4, 4), np.nan))
        out["ego"].append(ekf.filter_pose())
    return {k: np.array(v) for k, v in out.items()}


def test_gnss_time_compensation(raw_map):
    from elimaloc_b200 import ekf as pekf
    ekf = R.EkfLocalizationNode(pekf.make_ekf_config())
    assert ekf.time_compensate(100.0, [0, 0, 0], [1, 0, 0, 0]) is None            # empty state queue
    ekf.pcm_init_odom(100.0, [1.0, 2.0, 0.5], [1, 0, 0, 0])
    for k in range(12):                                                            # leave the PCM-init phase
        ekf.imu(100.0 + 0.01 * k, [0, 0, 0.2], [0.1, 0.3, 9.81])
        ekf.pcm_odom(100.0 + 0.01 * k, [1.0, 2.0, 0.5], [1, 0, 0, 0], np.diag([0.01] * 3 + [1e-4] * 3))
    for k in range(12, 80):
        ekf.imu(100.0 + 0.01 * k, [0.01, -0.02, 0.2], [0.4, 0.3, 9.81])
    rows = ekf.state_queue()
    deq = [np.concatenate([r, np.zeros(19)]) for r in rows]
    assert len(rows) > 50 and np.ptp(rows[:, 1]) > 1e-3 and np.ptp(rows[:, 6]) > 1e-3  # the filter moved and turned
    q_in = H.R_to_quat_wxyz(synth.exp_so3([0.02, -0.01, 0.7]))
    for t in (rows[0, 0] - 0.5, rows[10, 0], rows[30, 0] + 0.004, rows[-1, 0] - 0.0301, rows[-1, 0], rows[-1, 0] + 0.2):
        got = ekf.time_compensate(t, [3.0, -1.0, 0.2], q_in)
        want = H.gnss_time_compensation(dict(t=t, pos=np.array([3.0, -1.0, 0.2]), quat=q_in), deq)
        assert (got is None) == (want is None), t
        if got is not None:
            assert got["t"] == want["t"] and np.abs(got["pos"] - want["pos"]).max() < 1e-12 and np.abs(got["quat"] - want["quat"]).max() < 1e-12, t


@pytest.mark.parametrize("world_cls,offset,m_raw,box,n_points,n_scans,ds", [(H.World, 0.0, 250_000, 30.0, 4096, 16, 0.0), (PinnedSpanWorld, 0.0037, 250_000, 30.0, 4096, 16, 0.0),
                                                                            (H.World, 0.0, 400_000, 40.0, 8192, 30, 0.0),   # the world of the GPU test
                                                                            (H.World, 0.0, 400_000, 40.0, 8192, 30, 1.5)])  # with the node's default down-sampling
def test_closed_loop_on_the_reference_nodes_matches_the_harness(world_cls, offset, m_raw, box, n_points, n_scans, ds):
    """BASELINE config 5: deskew -> AVGICP -> time compensation -> EKF update at 10 Hz with a 100 Hz IMU, 16 scans, on the
    reference's own two nodes (PcmMatching + EkfLocalization) against tests/pipeline_harness.run with the oracle arm — the
    harness whose GPU arm tests/test_pipeline.py checks on the B200.  offset 0: scan stamps coincide with IMU / odometry
    stamps, where the node's exact `<` comparisons decide which odometry sample is used."""
    from elimaloc_b200 import ekf as pekf
    raw = synth.map_s(m_raw, box)  # the worlds of tests/test_pipeline.py
    ref = run_reference_nodes(raw, world_cls(box, n_points, seed=7), n_scans, pekf.make_ekf_config(), scan_offset=offset, input_voxel_ds_m=ds or 0.001)
    har = H.run(H.OracleArm(raw, {}), world_cls(box, n_points, seed=7), n_scans, scan_offset=offset, input_voxel_ds_m=ds)
    assert ref["ok"].all() and har["ok"].all()
    d_icp, d_ego = np.abs(ref["icp"] - har["icp"]).max(), np.abs(ref["ego"] - har["ego"]).max()
    print("closed loop, reference nodes vs harness: max |d icp pose| %.3g, max |d filter pose| %.3g" % (d_icp, d_ego))
    tol = 2e-5 if ds == 0.0 else 1e-4  # down-sampled: the node feeds the survivors in hash-table order, the harness in input order
    assert d_icp < tol and d_ego < tol
    # and both follow the true trajectory (AVGICP with 1 m voxels is a coarse estimator: decimetres)
    # (not asserted with the 1.5 m down-sampling: on so few points the AVGICP loop loses track in this synthetic world after
    # ~20 scans — on the reference's nodes exactly as in the harness, which is what this test is about)
    if ds == 0.0:
        w = world_cls(box, n_points, seed=7)
        err = max(np.linalg.norm(ref["icp"][i][:3, 3] - w.pose(ref["t"][i])[:3, 3]) for i in range(n_scans))
        assert err < 0.6


@pytest.mark.parametrize("last_odom_k", [7, 12, 30])
def test_deskew_odometry_span_incl_the_integrate_branch(raw_map, last_odom_k):
    """OdomDeskewInfo's two poses: when the odometry queue ends BEFORE the scan end the node integrates the latest message forward
    with its own twist (pcm_matching.cpp:648-706) — the harness' deskew_odometry_span against the node's resulting increments"""
    node = R.PcmMatchingNode(raw_map[:100])
    deq, stamps, gyro = [], [], []
    for k in range(-5, 30):
        t = T0 + 0.01 * k
        p, rpy = ego_pose(t, (2.0, 1.0, 0.5))
        q = tf_quat_from_rpy(*rpy)
        if k <= last_odom_k:
            node.odom(t, p, q, lin=(0.9, 0.05, -0.02), ang=(0.01, -0.02, 0.3))
            deq.append(dict(t=t, pos=p, quat=q, vel_local=np.array([0.9, 0.05, -0.02]), rate=np.array([0.01, -0.02, 0.3])))
        g = [0.01, -0.02, 0.3]
        node.imu(t + 0.003, g, [0.0, 0.0, 9.81])
        stamps.append(t + 0.003)
        gyro.append(g)
    n = 500
    xyz = synth.scan_u(n, 30.0, seed=1)
    rel = np.linspace(0.0, 0.1, n).astype(np.float32)
    t_cur = T0 + 0.05
    t_end = t_cur + float(rel[-1])
    ok, und, tab = node.deskew(t_cur, xyz, rel)
    span = H.deskew_odometry_span(deq, t_cur, t_end)
    assert ok == (span is not None)
    if span is None:
        return
    assert (deq[-1]["t"] > t_end) == (last_odom_k == 30)                     # 7, 12: the integrate branch; 30: interpolation
    ot = O.deskew_tables(np.array(stamps), np.array(gyro), t_cur, t_end, *span)
    assert np.abs(ot["odom_incre"] - tab["odom_incre"]).max() < 2e-7
    assert np.abs(und - O.deskew_points(ot, xyz, rel)).max() < 2e-5
    assert_product_tables_equal_the_nodes(product_tables(deq, np.array(stamps), np.array(gyro), t_cur, t_end), tab, ok, 2e-7)


@pytest.mark.parametrize("seed", range(14))
def test_randomised_deskew_streams(raw_map, seed):
    """irregular IMU and odometry rates, gaps, scan spans of different length, both stamp conventions, queues that start late or
    end early: the node's decision (deskew or refuse), its tables and its undistorted cloud against the oracle + the harness'
    odometry selection"""
    rng = np.random.default_rng(900 + seed)
    mode = int(rng.integers(0, 2))
    node = R.PcmMatchingNode(raw_map[:100], scan_time_end=mode)
    span = float(rng.uniform(0.03, 0.12))
    t_cur = T0 + float(rng.uniform(0.02, 0.08))
    deq, stamps, gyro = [], [], []
    t = T0 - 0.1 + float(rng.uniform(0.0, 0.2))                        # sometimes the odometry starts after the scan start
    t_stop = t_cur + span + float(rng.uniform(-0.03, 0.06))           # sometimes it ends before the scan end
    while t < t_stop:
        p, rpy = ego_pose(t, (2.0, 1.0, 0.5))
        q = tf_quat_from_rpy(*rpy)
        lin, ang = rng.normal(0, 1.0, 3), rng.normal(0, 0.2, 3)
        node.odom(t, p, q, lin=lin, ang=ang)
        deq.append(dict(t=t, pos=p, quat=q, vel_local=lin, rate=ang))
        t += float(rng.choice([0.005, 0.01, 0.02, 0.035]))
    t = T0 - 0.05
    while t < t_cur + span + 0.05:
        g = rng.normal(0, 0.3, 3)
        node.imu(t, g, [0.0, 0.0, 9.81])
        stamps.append(t)
        gyro.append(g)
        t += float(rng.choice([0.002, 0.005, 0.01, 0.025]))
    n = 800
    xyz = synth.scan_u(n, 40.0, seed=seed)
    rel = np.sort(rng.random(n).astype(np.float32) * np.float32(span))
    if mode:  # stamp = scan end, per-point times <= 0
        rel_in, stamp_in = (rel - rel[-1]).astype(np.float32), t_cur + float(rel[-1])
        t_end = stamp_in
        t_start = t_end + float(rel_in[0])
        rel_eff = (rel_in - rel_in[0]).astype(np.float32)
    else:
        rel_in, stamp_in, t_start, t_end, rel_eff = rel, t_cur, t_cur, t_cur + float(rel[-1]), rel
    ok, und, tab = node.deskew(stamp_in, xyz, rel_in)
    sp = H.deskew_odometry_span(deq, t_start, t_end)
    ot = O.deskew_tables(np.array(stamps), np.array(gyro), t_start, t_end, *(sp if sp is not None else (None, 0.0, None, 0.0)))
    pt = product_tables(deq, np.array(stamps), np.array(gyro), t_start, t_end)
    if sp is None:
        assert not ok and not tab["odom_available"] and not pt["odom_available"]
        return
    assert ok == (ot["imu_available"] and ot["odom_available"])
    assert tab["imu_pointer_cur"] == ot["imu_pointer_cur"]
    if ok:
        k = tab["imu_pointer_cur"]
        for name in ("imu_time", "imu_rot_x", "imu_rot_y", "imu_rot_z"):
            assert np.abs(ot[name][:k + 1] - tab[name][:k + 1]).max() < 1e-14, name
        # float32 increments: when the two odometry samples are closer in time than the sweep is long the interpolation ratio
        # exceeds 1 and amplifies the float32 rounding of their difference (the node and the oracle round differently: the node's
        # angles go through a quaternion and back)
        ratio = (t_end - t_start) / max(sp[3] - sp[1], 1e-9)
        assert np.abs(ot["odom_incre"] - tab["odom_incre"]).max() < 5e-7 * max(1.0, ratio)
        assert np.array_equal(und, O.deskew_points(tab, xyz, rel_eff))
        assert_product_tables_equal_the_nodes(pt, tab, ok, 5e-7 * max(1.0, ratio))
    else:
        assert not (pt["imu_available"] and pt["odom_available"])
