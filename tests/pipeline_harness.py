"""BASELINE config 5 harness: deskew -> AVGICP registration -> covariance shaping -> time compensation -> EKF update, scans
streamed at 10 Hz with a 100 Hz IMU, closing the loop through the EKF pose like the two ROS nodes do.

The per-point / per-matrix work is done by an ARM (the oracle on the CPU or the product on the GPU); the small host glue
between those steps is restated ONCE here in numpy and shared by both arms, so it cannot be a source of divergence
(SURVEY.md section 8 row a22).  Glue restated (reference file:line):
  GetInterpolatedPose      pcm_matching/src/pcm_matching.cpp:933-1045 (float32 Affine3f, slerp lfun.hpp:216-241)
  sync_lidar_pose          pcm_matching.cpp:266                        icp_ego_pose  :298
  PublishPcmOdom shaping   pcm_matching.cpp:1082-1098, pcm_matching.hpp:222-273
  CallbackPcmOdom          ekf_localization/src/ekf_localization.cpp:147-179
  GnssTimeCompensation     ekf_localization.cpp:323-394
  ImuDeskewInfo / OdomDeskewInfo  pcm_matching.cpp:533-729 (through oracle.deskew_tables — table builders are glue)
Not modelled: ROS transport delays other than one fixed processing latency, VoxelDownsample (disabled: cell 0 keeps
every point, the default 1.5 m cell would leave a few thousand points), the Ouster index sampling (Q17)."""
import numpy as np

from elimaloc_b200 import ekf as pekf, synth
from oracle import oracle as O

F32 = np.float32


# ------------------------------------------------------------------------------------------------ small maths (glue)
def rpy_to_R(r, p, y):
    return synth.exp_so3([0, 0, y]) @ synth.exp_so3([0, p, 0]) @ synth.exp_so3([r, 0, 0])


def R_to_quat_wxyz(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        return np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
    q = np.zeros(4)
    q[0] = (R[k, j] - R[j, k]) / s
    q[1 + i] = 0.25 * s
    q[1 + j] = (R[j, i] + R[i, j]) / s
    q[1 + k] = (R[k, i] + R[i, k]) / s
    return q


def quat_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rot_to_vec(R):  # lfun.hpp RotToVec
    if abs(R[2, 0]) > 0.998:
        a = [0.0, np.pi / 2 * (1 if R[2, 0] >= 0 else -1), np.arctan2(-R[1, 2], R[1, 1])]
    else:
        p = np.arcsin(-R[2, 0])
        a = [np.arctan2(R[2, 1] / np.cos(p), R[2, 2] / np.cos(p)), p, np.arctan2(R[1, 0] / np.cos(p), R[0, 0] / np.cos(p))]
    return np.array([np.fmod(v + np.pi, 2 * np.pi) - np.pi for v in a])


def angle_diff(ref, rel):  # lfun.hpp AngleDiffRad
    d = rel - ref
    while d > np.pi:
        d -= 2 * np.pi
    while d < -np.pi:
        d += 2 * np.pi
    return d


def interpolate_tf_with_time(M, dt_scan, dt_trans):
    """InterpolateTfWithTime in float32 (lfun.hpp:216-241)"""
    if dt_trans == 0.0:
        return np.eye(4, dtype=F32)
    ratio = F32(dt_scan / dt_trans)
    trans = (M[:3, 3] * ratio).astype(F32)
    q = R_to_quat_wxyz(M[:3, :3].astype(np.float64)).astype(F32)
    ident = np.array([1, 0, 0, 0], F32)
    d = F32(np.dot(ident, q))
    ad = abs(d)
    if ad >= F32(1.0) - np.finfo(F32).eps:
        s0, s1 = F32(1) - ratio, ratio
    else:
        th = np.arccos(ad)
        st = np.sin(th)
        s0, s1 = F32(np.sin((F32(1) - ratio) * th) / st), F32(np.sin(ratio * th) / st)
    if d < 0:
        s1 = -s1
    qi = (s0 * ident + s1 * q).astype(F32)
    out = np.eye(4, dtype=F32)
    out[:3, :3] = quat_to_R(qi.astype(np.float64)).astype(F32)
    out[:3, 3] = trans
    return out


def odom_to_affine(o):
    M = np.eye(4, dtype=F32)
    M[:3, :3] = quat_to_R(o["quat"]).astype(F32)
    M[:3, 3] = o["pos"].astype(F32)
    return M


def get_interpolated_pose(deq, t):
    """pcm_matching.cpp:933-1045 (the extrapolation branch integrates the last twist)"""
    before = after = None
    for o in deq:
        if o["t"] <= t:
            before = o
        if o["t"] > t:
            after = o
            break
    if before is None:
        return None
    if after is None:
        last = deq[-1]
        dt = t - last["t"]
        r, p, y = rot_to_vec(quat_to_R(last["quat"]))
        Rg = rpy_to_R(r, p, y)
        pos = last["pos"] + Rg @ last["vel_local"] * dt
        r, p, y = r + last["rate"][0] * dt, p + last["rate"][1] * dt, y + last["rate"][2] * dt
        # odom_after is default-constructed in the reference and only its pose is filled: header.stamp stays 0, so
        # d_time_after = 0 and the ratio dt_scan / (0 - d_time_before) is a tiny NEGATIVE number — the extrapolated pose is in
        # effect discarded (verified against the node itself, tests/test_reference_build_node.py)
        after = dict(t=0.0, pos=pos, quat=R_to_quat_wxyz(rpy_to_R(r, p, y)))
    A, B = odom_to_affine(before), odom_to_affine(after)
    between = (np.linalg.inv(A.astype(np.float64)) @ B.astype(np.float64)).astype(F32)
    interp = interpolate_tf_with_time(between, t - before["t"], after["t"] - before["t"])
    return (A @ interp).astype(F32)


def normalize_covariance(c):  # pcm_matching.hpp:247-273
    c = c.copy()
    m = min(c[0, 0], c[1, 1], c[2, 2])
    if m <= 1e-9:
        c *= 1e9
        m = min(c[0, 0], c[1, 1], c[2, 2])
        if m < 1e-9:
            m = 1e-9
    return np.minimum(c / m, 5.0)


def shape_pcm_covariance(R_ego, local_cov, fitness):
    """PublishPcmOdom (pcm_matching.cpp:1082-1098): 6x6 row-major pose covariance"""
    std = max(fitness, 0.25)
    tc = R_ego @ local_cov[:3, :3] @ R_ego.T
    rc = local_cov[3:, 3:]
    ang = std * np.pi / 180.0
    return normalize_covariance(tc) * std * std, normalize_covariance(rc) * ang * ang


def gnss_time_compensation(meas, deq_state):
    """ekf_localization.cpp:323-394; meas = dict(t, pos, quat); deq_state = list of EgoState arrays (26)"""
    if not deq_state:
        return None
    cur = deq_state[-1]
    if deq_state[0][0] > meas["t"]:
        return None
    closest = None
    for s in deq_state:
        if s[0] > meas["t"]:
            closest = s
            break
        closest = s
    gap = cur[0] - meas["t"]
    if gap <= 0.0:
        return dict(meas)
    d = np.zeros(6)
    if abs(cur[0] - closest[0]) > 1e-5:
        ratio = gap / (cur[0] - closest[0])
        d[:3] = (cur[1:4] - closest[1:4]) * ratio
        d[3:] = [angle_diff(closest[4 + i], cur[4 + i]) * ratio for i in range(3)]
    dq = R_to_quat_wxyz(rpy_to_R(d[3], d[4], d[5]))
    w1, x1, y1, z1 = meas["quat"]
    w2, x2, y2, z2 = dq
    q = np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                  w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2])
    return dict(t=cur[0], pos=meas["pos"] + d[:3], quat=q / np.linalg.norm(q))


def deskew_odometry_span(deq, t_cur, t_scan_end):
    """OdomDeskewInfo's choice of the two poses the sweep is interpolated between (pcm_matching.cpp:587-729), as
    (pose6 start, stamp start, pose6 end, stamp end) with pose6 = x, y, z, roll, pitch, yaw; None when the node refuses
    (empty queue, or its first message is newer than the scan start).  Start: the first message not older than the scan
    start.  End: the first message not older than the scan end if the queue reaches beyond it, otherwise the LATEST message
    integrated forward with its own twist (local linear velocity rotated by its attitude, Euler rates added to the angles)."""
    if not deq or deq[0]["t"] > t_cur:
        return None
    s = deq[-1]
    for o in deq:
        s = o
        if not o["t"] < t_cur:
            break
    ps = np.concatenate([s["pos"], rot_to_vec(quat_to_R(s["quat"]))])
    if deq[-1]["t"] > t_scan_end:
        e = deq[-1]
        for o in deq:
            e = o
            if not o["t"] < t_scan_end:
                break
        return ps, s["t"], np.concatenate([e["pos"], rot_to_vec(quat_to_R(e["quat"]))]), e["t"]
    last = deq[-1]
    dt = t_scan_end - last["t"]
    r, p, y = rot_to_vec(quat_to_R(last["quat"]))
    pos = last["pos"] + rpy_to_R(r, p, y) @ last["vel_local"] * dt
    rpy = np.array([r, p, y]) + last["rate"] * dt
    rpy = np.array(rot_to_vec(rpy_to_R(*rpy)))  # the node goes through a quaternion (setRPY) and back (getRPY)
    return ps, s["t"], np.concatenate([pos, rpy]), t_scan_end


# ------------------------------------------------------------------------------------------------ arms
class OracleArm:
    name = "oracle"
    shape_covariance = staticmethod(lambda R, cov, fit: O.shape_pcm_covariance(R, cov, fit))

    def __init__(self, raw_map, ekf_cfg_kwargs):
        from elimaloc_b200 import _capi
        self.map = O.VoxelHashMap(1.0, 30)
        self.map.AddPoints(raw_map)
        self.map.CalVoxelCovAll()
        self.reg = O.Registration()
        self.ekf = O.EkfAlgorithm(pekf.make_ekf_config(**ekf_cfg_kwargs), _capi.EkfState)
        self.cfg = O.make_config(icp_method=O.AVGICP, max_iteration=10, max_thread=8, max_fitness_score=2.0)

    def stored(self):
        return self.map.export()["pxyz"]

    def deskew(self, xyz, rel, tab):
        return O.deskew_points(tab, xyz, rel)

    def downsample(self, xyz, voxel_size):
        return xyz[O.scan_preprocess(xyz, 0.0, voxel_size)]

    def register(self, scan, T0):
        r = self.reg.RunRegister(scan, self.map, T0, self.cfg)
        return r["pose"], r["is_success"], r["fitness_score"], r["local_cov"]

    def ekf_pose(self):
        return np.concatenate([np.array(self.ekf.s.pos[:]), np.array(self.ekf.s.rot[:])])


class GpuArm:
    name = "gpu"

    @staticmethod
    def shape_covariance(R, cov, fit):
        import elimaloc_b200 as E
        return E.shape_pcm_covariance(R, cov, fit)

    def __init__(self, raw_map, ekf_cfg_kwargs, device=0):
        import elimaloc_b200 as E
        self.map = E.VoxelHashMap(1.0, 30, device=device)
        self.map.AddPoints(raw_map)
        self.map.CalVoxelCovAll()
        self.reg = E.Registration(device=device)
        self.ekf = E.EkfAlgorithm(pekf.make_ekf_config(**ekf_cfg_kwargs), device=device)
        self.cfg = E.RegistrationConfig(icp_method=E.AVGICP, max_iteration=10, max_fitness_score=2.0)

    def stored(self):
        return self.map.Pointcloud()

    def deskew(self, xyz, rel, tab):
        return self.reg.DeskewPoints(xyz, rel, tab)

    def downsample(self, xyz, voxel_size):
        return self.reg.PreprocessScan(xyz, 0.0, voxel_size)[0]

    def register(self, scan, T0):
        return self.reg.RunRegister(scan, self.map, T0, self.cfg)

    def ekf_pose(self):
        st = self.ekf.state()
        return np.concatenate([np.array(st.pos[:]), np.array(st.rot[:])])


# ------------------------------------------------------------------------------------------------ world + loop
World = synth.ScanWorld  # constant-twist arc through the map; IMU in the body frame; raw scans with per-point motion distortion


def run(arm, world, n_scans, imu_dt=0.01, latency=0.03, scan_offset=0.0, input_voxel_ds_m=0.0, input_max_dist=0.0):
    """returns dict of per-scan arrays: icp pose, EKF pose (pos + quaternion) after the update, success flag, fitness.
    scan_offset shifts the scan stamps off the IMU grid (the reference selects odometry with exact `<` on the stamps);
    input_voxel_ds_m > 0 keeps the first point of every voxel of that size before the registration, as the node does
    (VoxelDownsample, pcm_matching.cpp:257-258; localization.ini: 1.5 m) — off by default."""
    stored = arm.stored()
    t0 = world.t0
    T0 = world.pose(t0)
    arm.ekf.RunGnssUpdate(pekf.make_measurement(t0, T0[:3, 3], R_to_quat_wxyz(T0[:3, :3]), np.eye(3) * 1e-9, np.eye(3) * 1e-9,
                                                source=pekf.PCM_INIT))  # CallbackPcmInitOdom (ekf_localization.cpp:181-210)
    deq_odom, deq_state, imu_log = [], [], []
    out = dict(icp=[], ego=[], ok=[], fit=[], t=[])
    k_imu = 0
    pending = None

    def step_imu(until):
        nonlocal k_imu
        while t0 + k_imu * imu_dt <= until + 1e-9:
            t = t0 + k_imu * imu_dt
            g, a = world.imu(t)
            imu_log.append((t, g))
            arm.ekf.RunPredictionImu(t, g, a)
            ego = arm.ekf.GetCurrentState()                       # PublishInThread (ekf_localization.cpp:400-...)
            deq_state.append(ego.copy())
            R = rpy_to_R(ego[4], ego[5], ego[6])
            deq_odom.append(dict(t=ego[0], pos=ego[1:4].copy(), quat=R_to_quat_wxyz(R), vel_local=ego[10:13].copy(), rate=ego[7:10].copy()))
            k_imu += 1

    for s in range(n_scans):
        t_end = t0 + 0.1 * (s + 1) + scan_offset
        t_cur = t_end - 0.1
        step_imu(t_end + latency)                                  # the result is applied `latency` after the scan end
        xyz, rel = world.scan(stored, t_end)
        if input_max_dist > 0.0:                                   # FilterPointsByDistance runs first (pcm_matching.cpp:235)
            keep = O.scan_preprocess(xyz, input_max_dist, 0.0)
            xyz, rel = xyz[keep], rel[keep]
        t_scan_end = t_cur + float(rel[-1])                        # d_time_scan_end_ = stamp + time of the last point (pcm.cpp:474)
        st = np.array([x[0] for x in imu_log])
        gy = np.array([x[1] for x in imu_log])
        sel = st >= t_cur - 0.05
        # start / end odometry of the scan span: OdomDeskewInfo skips messages with stamp < scan start / scan end — exact
        # comparisons on the doubles, no tolerance (pcm_matching.cpp:611-617, 640-647; checked against the node itself in
        # tests/test_reference_build_node.py, where a tolerance here showed up as centimetres on stamps that coincide)
        ps, ts, pe, te = deskew_odometry_span(deq_odom, t_cur, t_scan_end)
        tab = O.deskew_tables(st[sel], gy[sel], t_cur, t_scan_end, ps, ts, pe, te)
        und = arm.deskew(xyz, rel, tab)
        if input_voxel_ds_m > 0.0:
            und = arm.downsample(und, input_voxel_ds_m)
        sync = get_interpolated_pose(deq_odom, t_scan_end)
        T_init = sync.astype(np.float64)                           # tf_ego_to_lidar = identity (pcm_matching.cpp:266)
        pose, ok, fit, cov = arm.register(und, T_init)
        out["icp"].append(np.array(pose)); out["ok"].append(bool(ok)); out["fit"].append(float(fit)); out["t"].append(t_end)
        if ok:                                                     # pcm_matching.cpp:289-299
            c66 = arm.shape_covariance(pose[:3, :3], np.array(cov), fit)   # product / oracle implementation of the arm
            pc, rc = c66[:3, :3].copy(), c66[3:, 3:].copy()
            meas = gnss_time_compensation(dict(t=t_scan_end, pos=pose[:3, 3].copy(), quat=R_to_quat_wxyz(pose[:3, :3])), deq_state)
            if meas is not None:
                arm.ekf.RunGnssUpdate(pekf.make_measurement(meas["t"], meas["pos"], meas["quat"], pc, rc, source=pekf.PCM))
        out["ego"].append(arm.ekf_pose())                          # raw filter pose (pos, quaternion) after the update
    return {k: np.array(v) for k, v in out.items()}


class ChainArm(GpuArm):
    """the product's device-resident scan chain (elm_scan_pipeline_*): tables built by the product, point data never leaves HBM,
    the EKF update reads the IcpState where it lies"""
    name = "chain"

    def __init__(self, raw_map, ekf_cfg_kwargs, device=0, input_voxel_ds_m=0.0, input_max_dist=0.0):
        import elimaloc_b200 as E
        super().__init__(raw_map, ekf_cfg_kwargs, device)
        self.ekf.enable_state_ring(True)
        self.pipe = E.ScanPipeline(self.reg, input_max_dist=input_max_dist, input_voxel_ds_m=input_voxel_ds_m)


def run_chain(arm, world, n_scans, imu_dt=0.01, latency=0.03, scan_offset=0.0):
    """run() with the per-scan work done by ChainArm.pipe: the same event order, the same numpy GetInterpolatedPose; what the
    stage-wise arms compute on the host (tables, covariance shaping, time compensation) is computed by the product here."""
    import elimaloc_b200 as E
    stored = arm.stored()
    t0 = world.t0
    T0 = world.pose(t0)
    arm.ekf.RunGnssUpdate(pekf.make_measurement(t0, T0[:3, 3], R_to_quat_wxyz(T0[:3, :3]), np.eye(3) * 1e-9, np.eye(3) * 1e-9, source=pekf.PCM_INIT))
    deq_odom, imu_t, imu_g = [], [], []
    out = dict(icp=[], ego=[], ok=[], fit=[], t=[], n=[])
    k_imu = 0

    def step_imu(until):
        nonlocal k_imu
        while t0 + k_imu * imu_dt <= until + 1e-9:
            t = t0 + k_imu * imu_dt
            g, a = world.imu(t)
            imu_t.append(t)
            imu_g.append(g)
            arm.ekf.RunPredictionImu(t, g, a)                      # (+ GetCurrentState and the deque push, on the device)
            ego = arm.ekf.GetCurrentState()
            R = rpy_to_R(ego[4], ego[5], ego[6])
            deq_odom.append(dict(t=ego[0], pos=ego[1:4].copy(), quat=R_to_quat_wxyz(R), vel_local=ego[10:13].copy(), rate=ego[7:10].copy()))
            k_imu += 1

    for s in range(n_scans):
        t_end = t0 + 0.1 * (s + 1) + scan_offset
        t_cur = t_end - 0.1
        step_imu(t_end + latency)
        xyz, rel = world.scan(stored, t_end)
        wxyz = np.array([o["quat"] for o in deq_odom])
        q = E.Queues(imu_t, np.array(imu_g), [o["t"] for o in deq_odom], np.array([o["pos"] for o in deq_odom]), wxyz[:, [1, 2, 3, 0]],
                     np.array([o["vel_local"] for o in deq_odom]), np.array([o["rate"] for o in deq_odom]))
        ok_d, _, t_scan_end = arm.pipe.deskew(xyz, rel, t_cur, q)
        assert ok_d
        sync = get_interpolated_pose(deq_odom, t_scan_end)
        arm.pipe.register(arm.map, sync.astype(np.float64), arm.cfg)
        arm.pipe.ekf_update(arm.ekf)
        r = arm.pipe.fetch()
        out["icp"].append(r["T_lidar"]); out["ok"].append(r["is_success"]); out["fit"].append(r["fitness_score"]); out["t"].append(t_end)
        out["n"].append(r["n_registered"])
        out["ego"].append(arm.ekf_pose())
    return {k: np.array(v) for k, v in out.items()}
