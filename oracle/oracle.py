"""ORACLE — TEST INFRASTRUCTURE ONLY.  Parity status: registration + voxel map PINNED on oracle/_ref (the reference's own
sources compiled against stand-in Eigen/oneTBB/ROS/PCL headers, oracle/reference_build.py), and so are the EKF and the
node-side stages (deskew, distance filter, pose interpolation, covariance shaping).

ctypes front-end of oracle/liboracle.so, the CPU restatement of the reference's
pcm_matching hot path (registration.cpp / voxel_hash_map.{hpp,cpp}).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

P2P, GICP, VGICP, AVGICP = 0, 1, 2, 3


class RegConfig(C.Structure):
    """Field-for-field mirror of orc_reg_config (oracle/capi.cpp); reg.hpp:62-85 subset."""
    _fields_ = [("icp_method", C.c_int32), ("max_iteration", C.c_int32), ("max_thread", C.c_int32),
                ("use_radar_cov", C.c_int32), ("debug_print", C.c_int32), ("reserved0", C.c_int32),
                ("max_search_dist", C.c_double), ("lm_lambda", C.c_double),
                ("icp_termination_threshold_m", C.c_double), ("min_overlap_ratio", C.c_double),
                ("max_fitness_score", C.c_double), ("range_variance_m", C.c_double),
                ("azimuth_variance_deg", C.c_double), ("elevation_variance_deg", C.c_double)]


def make_config(**kw):
    """Defaults are config/localization.ini:80-109 of the reference."""
    d = dict(icp_method=GICP, max_iteration=10, max_thread=1, use_radar_cov=0, debug_print=0, reserved0=0,
             max_search_dist=5.0, lm_lambda=0.5, icp_termination_threshold_m=0.02, min_overlap_ratio=0.4,
             max_fitness_score=0.5, range_variance_m=1.0, azimuth_variance_deg=0.4, elevation_variance_deg=0.4)
    d.update(kw)
    return RegConfig(**d)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp")) or f == "Makefile"]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        # shared hosts: never let idle OpenMP workers spin (an oversubscribed box turns every barrier into seconds)
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        os.environ.setdefault("OMP_NUM_THREADS", str(min(os.cpu_count() or 1, 16)))
        L = C.CDLL(build())
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.orc_map_create.restype = C.c_void_p
        L.orc_map_create.argtypes = [C.c_double, C.c_int]
        L.orc_map_destroy.argtypes = [C.c_void_p]
        L.orc_map_add_points.argtypes = [C.c_void_p, fp, C.c_size_t]
        L.orc_map_cal_voxel_cov.argtypes = [C.c_void_p]
        L.orc_map_cal_point_cov.argtypes = [C.c_void_p, C.c_double]
        L.orc_map_num_voxels.restype = C.c_size_t
        L.orc_map_num_voxels.argtypes = [C.c_void_p]
        L.orc_map_num_points.restype = C.c_size_t
        L.orc_map_num_points.argtypes = [C.c_void_p]
        L.orc_map_export.argtypes = [C.c_void_p, ip, ip, dp, dp, fp, dp, dp]
        L.orc_reg_create.restype = C.c_void_p
        L.orc_reg_destroy.argtypes = [C.c_void_p]
        L.orc_run_register.argtypes = [C.c_void_p, C.c_void_p, fp, C.c_size_t, dp, C.POINTER(RegConfig), dp, ip, dp,
                                       dp, C.c_int32, ip, dp, dp, dp, dp, dp, dp]
        L.orc_linearize.argtypes = [C.c_void_p, C.c_void_p, fp, C.c_size_t, dp, C.POINTER(RegConfig), dp, dp, dp,
                                    C.POINTER(C.c_longlong)]
        L.orc_correspondences.argtypes = [C.c_void_p, fp, C.c_size_t, dp, C.c_int, C.c_double, ip, dp]
        L.orc_time_register.restype = C.c_double
        L.orc_time_register.argtypes = [C.c_void_p, C.c_void_p, fp, C.c_size_t, dp, C.POINTER(RegConfig), ip]
        L.orc_deskew_tables.argtypes = [dp, dp, C.c_int, dp, C.c_double, dp, C.c_double, C.c_double, C.c_double, dp, dp, dp, dp, ip, fp]
        L.orc_deskew_points.argtypes = [dp, dp, dp, dp, ip, fp, C.c_double, C.c_double, fp, fp, C.c_size_t, fp]
        L.orc_ekf_state_size.restype = C.c_size_t
        L.orc_ekf_init.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_ekf_predict_imu.argtypes = [C.c_void_p, C.c_void_p, C.c_double, dp, dp]
        L.orc_ekf_update_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_ekf_get_current_state.argtypes = [C.c_void_p, dp]
        L.orc_shape_pcm_covariance.argtypes = [dp, dp, C.c_double, dp]
        L.orc_find_ground_height.argtypes = [C.c_void_p, C.c_double, C.c_double, dp]
        L.orc_scan_preprocess.restype = C.c_size_t
        L.orc_scan_preprocess.argtypes = [fp, C.c_size_t, C.c_double, C.c_double, ip]
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _xyz(a):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)
    return a


class VoxelHashMap:
    """Mirror of the reference's VoxelHashMap (vhm.hpp:89-335) over the oracle."""

    def __init__(self, voxel_size=1.0, max_points_per_voxel=30):
        self._h = lib().orc_map_create(float(voxel_size), int(max_points_per_voxel))
        self.voxel_size = float(voxel_size)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_map_destroy(self._h)
            self._h = None

    def AddPoints(self, xyz):
        xyz = _xyz(xyz)
        lib().orc_map_add_points(self._h, _f(xyz), xyz.shape[0])

    def CalVoxelCovAll(self):
        lib().orc_map_cal_voxel_cov(self._h)

    def CalPointCovAll(self, d):
        lib().orc_map_cal_point_cov(self._h, float(d))

    def num_voxels(self):
        return lib().orc_map_num_voxels(self._h)

    def num_points(self):
        return lib().orc_map_num_points(self._h)

    def Empty(self):
        return self.num_voxels() == 0

    def FindGroundHeight(self, position_xy):
        """vhm.hpp:285-322: (found, ground_z)"""
        z = C.c_double(0.0)
        found = lib().orc_find_ground_height(self._h, float(position_xy[0]), float(position_xy[1]), C.byref(z))
        return bool(found), float(z.value)

    def export(self):
        V, P = self.num_voxels(), self.num_points()
        out = dict(keys=np.zeros((V, 3), np.int32), counts=np.zeros(V, np.int32), vmean=np.zeros((V, 3)),
                   vcov=np.zeros((V, 3, 3)), pxyz=np.zeros((P, 3), np.float32), pmean=np.zeros((P, 3)),
                   pcov=np.zeros((P, 3, 3)))
        lib().orc_map_export(self._h, _i(out["keys"]), _i(out["counts"]), _d(out["vmean"]), _d(out["vcov"]),
                             _f(out["pxyz"]), _d(out["pmean"]), _d(out["pcov"]))
        return out


class Registration:
    """Mirror of the reference's Registration (reg.hpp:101-230) over the oracle."""

    def __init__(self):
        self._h = lib().orc_reg_create()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_reg_destroy(self._h)
            self._h = None

    def RunRegister(self, source_local, voxel_map, initial_guess, cfg, fitness_in=0.0, max_trace=64):
        src = _xyz(source_local)
        T0 = np.ascontiguousarray(initial_guess, dtype=np.float64).reshape(4, 4)
        T = np.zeros((4, 4))
        ok = np.zeros(1, np.int32)
        fit = np.array([fitness_in], np.float64)
        cov = np.zeros((6, 6))
        nit = np.zeros(1, np.int32)
        tr = dict(pose_in=np.zeros((max_trace, 4, 4)), JTJ=np.zeros((max_trace, 6, 6)), JTr=np.zeros((max_trace, 6)),
                  res=np.zeros(max_trace), ncorr=np.zeros(max_trace), pose_out=np.zeros((max_trace, 4, 4)))
        lib().orc_run_register(self._h, voxel_map._h, _f(src), src.shape[0], _d(T0), C.byref(cfg), _d(T), _i(ok),
                               _d(fit), _d(cov), max_trace, _i(nit), _d(tr["pose_in"]), _d(tr["JTJ"]), _d(tr["JTr"]),
                               _d(tr["res"]), _d(tr["ncorr"]), _d(tr["pose_out"]))
        n = int(nit[0])
        tr = {k: v[:min(n, max_trace)] for k, v in tr.items()}
        return dict(pose=T, is_success=bool(ok[0]), fitness_score=float(fit[0]), local_cov=cov, n_iter=n, trace=tr)

    def linearize(self, source_local, voxel_map, pose, cfg):
        src = _xyz(source_local)
        T0 = np.ascontiguousarray(pose, dtype=np.float64).reshape(4, 4)
        JTJ, JTr, res = np.zeros((6, 6)), np.zeros(6), np.zeros(1)
        nc = C.c_longlong(0)
        lib().orc_linearize(self._h, voxel_map._h, _f(src), src.shape[0], _d(T0), C.byref(cfg), _d(JTJ), _d(JTr),
                            _d(res), C.byref(nc))
        return dict(JTJ=JTJ, JTr=JTr, residual_sum=float(res[0]), n_corr=int(nc.value))

    def time_register(self, source_local, voxel_map, initial_guess, cfg):
        src = _xyz(source_local)
        T0 = np.ascontiguousarray(initial_guess, dtype=np.float64).reshape(4, 4)
        it = np.zeros(1, np.int32)
        sec = lib().orc_time_register(self._h, voxel_map._h, _f(src), src.shape[0], _d(T0), C.byref(cfg), _i(it))
        return sec, int(it[0])


def correspondences(voxel_map, source_local, pose, method, max_dist):
    src = _xyz(source_local)
    T0 = np.ascontiguousarray(pose, dtype=np.float64).reshape(4, 4)
    K = 7 if method == AVGICP else 1
    cnt = np.zeros(src.shape[0], np.int32)
    tgt = np.zeros((src.shape[0], K, 3))
    lib().orc_correspondences(voxel_map._h, _f(src), src.shape[0], _d(T0), int(method), float(max_dist), _i(cnt),
                              _d(tgt))
    return cnt, tgt


IMU_QUEUE_LENGTH = 2000


def deskew_tables(imu_stamp, gyro_xyz, time_scan_cur, time_scan_end, start_pose=None, start_stamp=0.0, end_pose=None, end_stamp=0.0):
    """ImuDeskewInfo + OdomDeskewInfo (pcm_matching.cpp:533-729) on plain arrays -> dict of the member tables."""
    st = np.ascontiguousarray(imu_stamp, dtype=np.float64)
    gy = np.ascontiguousarray(gyro_xyz, dtype=np.float64).reshape(-1, 3)
    t = dict(imu_time=np.zeros(IMU_QUEUE_LENGTH), imu_rot_x=np.zeros(IMU_QUEUE_LENGTH), imu_rot_y=np.zeros(IMU_QUEUE_LENGTH),
             imu_rot_z=np.zeros(IMU_QUEUE_LENGTH))
    mi, mf = np.zeros(3, np.int32), np.zeros(3, np.float32)
    sp = np.ascontiguousarray(start_pose, dtype=np.float64) if start_pose is not None else None
    ep = np.ascontiguousarray(end_pose, dtype=np.float64) if end_pose is not None else None
    lib().orc_deskew_tables(_d(st), _d(gy), len(st), _d(sp) if sp is not None else None, float(start_stamp),
                            _d(ep) if ep is not None else None, float(end_stamp), float(time_scan_cur), float(time_scan_end),
                            _d(t["imu_time"]), _d(t["imu_rot_x"]), _d(t["imu_rot_y"]), _d(t["imu_rot_z"]), _i(mi), _f(mf))
    t.update(imu_pointer_cur=int(mi[0]), imu_available=bool(mi[1]), odom_available=bool(mi[2]), odom_incre=mf.copy(),
             time_scan_cur=float(time_scan_cur), time_scan_end=float(time_scan_end))
    return t


def deskew_points(t, xyz, rel_time):
    """DeskewPoint over a scan (pcm_matching.cpp:499-511, 780-824), float32."""
    xyz = _xyz(xyz)
    rt = np.ascontiguousarray(rel_time, dtype=np.float32)
    out = np.zeros_like(xyz)
    mi = np.array([t["imu_pointer_cur"], int(t["imu_available"]), int(t["odom_available"])], np.int32)
    mf = np.ascontiguousarray(t["odom_incre"], dtype=np.float32)
    lib().orc_deskew_points(_d(t["imu_time"]), _d(t["imu_rot_x"]), _d(t["imu_rot_y"]), _d(t["imu_rot_z"]), _i(mi), _f(mf),
                            t["time_scan_cur"], t["time_scan_end"], _f(xyz), _f(rt), xyz.shape[0], _f(out))
    return out


class EkfAlgorithm:
    """Oracle EKF (oracle/ekf.cpp).  cfg / state / measurement are ctypes structures with the layout of
    elm_ekf_config / elm_ekf_state / elm_ekf_measurement (passed in by the tests so that both sides share them)."""

    def __init__(self, cfg, state_type):
        self.cfg = cfg
        self.s = state_type()
        assert C.sizeof(self.s) == lib().orc_ekf_state_size(), "EKF state layout mismatch between oracle and C ABI"
        lib().orc_ekf_init(C.byref(self.cfg), C.byref(self.s))

    def RunPredictionImu(self, t, gyro, acc):
        g = np.ascontiguousarray(gyro, dtype=np.float64)
        a = np.ascontiguousarray(acc, dtype=np.float64)
        return bool(lib().orc_ekf_predict_imu(C.byref(self.cfg), C.byref(self.s), float(t), _d(g), _d(a)))

    def RunGnssUpdate(self, meas):
        return bool(lib().orc_ekf_update_pose(C.byref(self.cfg), C.byref(self.s), C.byref(meas)))

    def GetCurrentState(self):
        ego = np.zeros(26)
        lib().orc_ekf_get_current_state(C.byref(self.s), _d(ego))
        return ego


def shape_pcm_covariance(R_ego, local_cov, icp_pose_std_m, cov36=None):
    """PublishPcmOdom's covariance shaping (pcm_matching.cpp:1082-1098): fills the two 3x3 blocks of a 6x6 row-major pose
    covariance (other entries keep the values of `cov36`, zeros by default)."""
    R = np.ascontiguousarray(R_ego, dtype=np.float64).reshape(3, 3)
    lc = np.ascontiguousarray(local_cov, dtype=np.float64).reshape(6, 6)
    out = np.zeros((6, 6)) if cov36 is None else np.ascontiguousarray(cov36, dtype=np.float64).reshape(6, 6).copy()
    lib().orc_shape_pcm_covariance(_d(R), _d(lc), float(icp_pose_std_m), _d(out))
    return out


def scan_preprocess(xyz, max_dist=0.0, voxel_size=0.0):
    """FilterPointsByDistance + VoxelDownsample (pcm_matching.cpp:451-465, voxel_hash_map.hpp:260-283): input indices of the
    survivors, in input order."""
    src = _xyz(xyz)
    idx = np.zeros(src.shape[0], np.int32)
    m = lib().orc_scan_preprocess(_f(src), src.shape[0], float(max_dist), float(voxel_size), _i(idx))
    return idx[:m]
