"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

ctypes front-end of oracle/_ref/libref.so: the REFERENCE's own registration.cpp + voxel_hash_map.{hpp,cpp}, compiled
unmodified from /root/reference against the stand-in third-party headers in oracle/ref_build/stubs (Eigen3, oneTBB and PCL are
absent from this image; see stubs/mini_eigen.hpp for what the stand-in does and does not preserve), and of
oracle/_ref/libref_ekf.so: the reference's own ekf_algorithm.cpp built the same way (plus ref_build/node_stubs), and of
oracle/_ref/libref_node.so: the ROS node class of pcm_matching.cpp itself (plus ref_build/node_stubs: ROS, tf, PCL, boost).
They are the pin of the oracle: tests/test_reference_build.py, test_reference_build_ekf.py and test_reference_build_node.py
run the oracle and these libraries on the same seeded inputs.

/root/reference exists only in the build container; `build()` compiles there, the GPU box uses the prebuilt file (git-ignored,
not gpurun-ignored).  Same call surface as oracle/oracle.py so one test body drives both.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import AVGICP, RegConfig, _d, _f, _i, _xyz

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref.so")
REFERENCE_ROOT = os.environ.get("ELM_REFERENCE_ROOT", "/root/reference")
_PCM = os.path.join(REFERENCE_ROOT, "src", "app", "localization", "pcm_matching")
_SO_EKF = os.path.join(_HERE, "_ref", "libref_ekf.so")
_SO_NODE = os.path.join(_HERE, "_ref", "libref_node.so")
_SO_EKFNODE = os.path.join(_HERE, "_ref", "libref_ekfnode.so")
_SO_NODE_ON_CUDA = os.path.join(_HERE, "_ref", "libref_node_on_cuda.so")
_LIB = None


def sources_present():
    return os.path.isfile(os.path.join(_PCM, "src", "registration.cpp"))


def available():
    return os.path.isfile(_SO) or sources_present()


def node_available():
    return (os.path.isfile(_SO_NODE) and os.path.isfile(_SO_EKFNODE)) or sources_present()


def node_on_cuda_available():
    return os.path.isfile(_SO_NODE_ON_CUDA) or sources_present()


def build_node_on_cuda():
    """The unmodified reference node compiled against shim/registration_shim.hpp and linked with the product library
    (needs elimaloc_b200/libelimaloc_b200.so to be built first)."""
    if sources_present():
        subprocess.check_call(["make", "-C", _HERE, "-s", "_ref/libref_node_on_cuda.so", "_ref/libref_node_on_shim_host.so", "REFERENCE_ROOT=" + REFERENCE_ROOT])
    if not os.path.isfile(_SO_NODE_ON_CUDA):
        raise FileNotFoundError("oracle/_ref/libref_node_on_cuda.so is not built and the reference sources are not here")
    return _SO_NODE_ON_CUDA


def ekf_available():
    return os.path.isfile(_SO_EKF) or sources_present()


def build(force=False):
    """Compile the reference's two translation units where they lie (never copied) + ref_build/ref_capi.cpp."""
    if sources_present():  # make decides whether anything is stale
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []) + ["_ref/libref.so", "REFERENCE_ROOT=" + REFERENCE_ROOT])
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []) + ["_ref/libref_ekf.so", "REFERENCE_ROOT=" + REFERENCE_ROOT])
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []) + ["_ref/libref_node.so", "REFERENCE_ROOT=" + REFERENCE_ROOT])
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []) + ["_ref/libref_ekfnode.so", "REFERENCE_ROOT=" + REFERENCE_ROOT])
    if not os.path.isfile(_SO):
        raise FileNotFoundError("oracle/_ref/libref.so is not built and the reference sources are not here")
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.ref_map_create.restype = C.c_void_p
        L.ref_map_create.argtypes = [C.c_double, C.c_int]
        L.ref_map_destroy.argtypes = [C.c_void_p]
        L.ref_map_add_points.argtypes = [C.c_void_p, fp, C.c_size_t]
        L.ref_map_cal_voxel_cov.argtypes = [C.c_void_p]
        L.ref_map_cal_point_cov.argtypes = [C.c_void_p, C.c_double]
        L.ref_map_num_voxels.restype = C.c_size_t
        L.ref_map_num_voxels.argtypes = [C.c_void_p]
        L.ref_map_num_points.restype = C.c_size_t
        L.ref_map_num_points.argtypes = [C.c_void_p]
        L.ref_map_empty.argtypes = [C.c_void_p]
        L.ref_map_export.argtypes = [C.c_void_p, ip, ip, dp, dp, fp, dp, dp]
        L.ref_reg_create.restype = C.c_void_p
        L.ref_reg_destroy.argtypes = [C.c_void_p]
        L.ref_run_register.argtypes = [C.c_void_p, C.c_void_p, fp, C.c_size_t, dp, C.POINTER(RegConfig), dp, ip, dp, dp, C.c_int32, ip, dp, dp]
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_time_register.restype = C.c_double
        L.ref_time_register.argtypes = [C.c_void_p, C.c_void_p, fp, C.c_size_t, dp, C.POINTER(RegConfig), ip]
        L.ref_reg_fitness.restype = C.c_double
        L.ref_reg_fitness.argtypes = [C.c_void_p]
        L.ref_linearize.argtypes = [C.c_void_p, C.c_void_p, fp, C.c_size_t, dp, C.POINTER(RegConfig), dp, dp, dp, C.POINTER(C.c_longlong)]
        L.ref_correspondences.argtypes = [C.c_void_p, fp, C.c_size_t, dp, C.c_int, C.c_double, ip, dp]
        L.ref_search_pairs.restype = C.c_size_t
        L.ref_search_pairs.argtypes = [C.c_void_p, fp, C.c_size_t, dp, C.c_int, C.c_double, ip, dp, C.c_size_t]
        L.ref_find_ground_height.argtypes = [C.c_void_p, C.c_double, C.c_double, dp]
        L.ref_voxel_downsample.restype = C.c_size_t
        L.ref_voxel_downsample.argtypes = [fp, C.c_size_t, C.c_double, ip]
        _LIB = L
    return _LIB


def set_threads(n):
    """Threads of the oneTBB stand-in (searches, CalVoxelCovAll, CalPointCovAll); results do not depend on it."""
    lib().ref_set_threads(int(n))


class VoxelHashMap:
    """The reference's VoxelHashMap (voxel_hash_map.hpp:89-335), itself."""

    def __init__(self, voxel_size=1.0, max_points_per_voxel=30):
        self._h = lib().ref_map_create(float(voxel_size), int(max_points_per_voxel))
        self.voxel_size = float(voxel_size)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_map_destroy(self._h)
            self._h = None

    def AddPoints(self, xyz):
        xyz = _xyz(xyz)
        lib().ref_map_add_points(self._h, _f(xyz), xyz.shape[0])

    def CalVoxelCovAll(self):
        lib().ref_map_cal_voxel_cov(self._h)

    def CalPointCovAll(self, d):
        lib().ref_map_cal_point_cov(self._h, float(d))

    def num_voxels(self):
        return lib().ref_map_num_voxels(self._h)

    def num_points(self):
        return lib().ref_map_num_points(self._h)

    def Empty(self):
        return bool(lib().ref_map_empty(self._h))

    def FindGroundHeight(self, position_xy):
        z = C.c_double(0.0)
        found = lib().ref_find_ground_height(self._h, float(position_xy[0]), float(position_xy[1]), C.byref(z))
        return bool(found), float(z.value)

    def export(self):
        V, P = self.num_voxels(), self.num_points()
        out = dict(keys=np.zeros((V, 3), np.int32), counts=np.zeros(V, np.int32), vmean=np.zeros((V, 3)), vcov=np.zeros((V, 3, 3)),
                   pxyz=np.zeros((P, 3), np.float32), pmean=np.zeros((P, 3)), pcov=np.zeros((P, 3, 3)))
        lib().ref_map_export(self._h, _i(out["keys"]), _i(out["counts"]), _d(out["vmean"]), _d(out["vcov"]), _f(out["pxyz"]), _d(out["pmean"]),
                             _d(out["pcov"]))
        return out


class Registration:
    """The reference's Registration (registration.hpp:101-230), itself."""

    def __init__(self):
        self._h = lib().ref_reg_create()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_reg_destroy(self._h)
            self._h = None

    def RunRegister(self, source_local, voxel_map, initial_guess, cfg, fitness_in=0.0, max_trace=64):
        """trace: A[i] = JTJ + lm_lambda * diag(JTJ) and b[i] = JTr of iteration i (the system handed to ldlt().solve)."""
        src = _xyz(source_local)
        T0 = np.ascontiguousarray(initial_guess, dtype=np.float64).reshape(4, 4)
        T = np.zeros((4, 4))
        ok = np.zeros(1, np.int32)
        fit = np.array([fitness_in], np.float64)
        cov = np.zeros((6, 6))
        nit = np.zeros(1, np.int32)
        A, b = np.zeros((max_trace, 6, 6)), np.zeros((max_trace, 6))
        lib().ref_run_register(self._h, voxel_map._h, _f(src), src.shape[0], _d(T0), C.byref(cfg), _d(T), _i(ok), _d(fit), _d(cov), max_trace,
                               _i(nit), _d(A), _d(b))
        n = int(nit[0])
        return dict(pose=T, is_success=bool(ok[0]), fitness_score=float(fit[0]), local_cov=cov, n_iter=n,
                    trace=dict(A=A[:min(n, max_trace)], b=b[:min(n, max_trace)]), d_fitness_score=float(lib().ref_reg_fitness(self._h)))

    def time_register(self, source_local, voxel_map, initial_guess, cfg):
        src = _xyz(source_local)
        T0 = np.ascontiguousarray(initial_guess, dtype=np.float64).reshape(4, 4)
        it = np.zeros(1, np.int32)
        sec = lib().ref_time_register(self._h, voxel_map._h, _f(src), src.shape[0], _d(T0), C.byref(cfg), _i(it))
        return sec, int(it[0])

    def linearize(self, source_local, voxel_map, pose, cfg):
        src = _xyz(source_local)
        T0 = np.ascontiguousarray(pose, dtype=np.float64).reshape(4, 4)
        JTJ, JTr, res = np.zeros((6, 6)), np.zeros(6), np.zeros(1)
        nc = C.c_longlong(0)
        lib().ref_linearize(self._h, voxel_map._h, _f(src), src.shape[0], _d(T0), C.byref(cfg), _d(JTJ), _d(JTr), _d(res), C.byref(nc))
        return dict(JTJ=JTJ, JTr=JTr, residual_sum=float(res[0]), n_corr=int(nc.value))


def correspondences(voxel_map, source_local, pose, method, max_dist):
    src = _xyz(source_local)
    T0 = np.ascontiguousarray(pose, dtype=np.float64).reshape(4, 4)
    K = 7 if method == AVGICP else 1
    cnt = np.zeros(src.shape[0], np.int32)
    tgt = np.zeros((src.shape[0], K, 3))
    lib().ref_correspondences(voxel_map._h, _f(src), src.shape[0], _d(T0), int(method), float(max_dist), _i(cnt), _d(tgt))
    return cnt, tgt


def search_pairs(voxel_map, source_local, pose, method, max_dist):
    """One whole-scan search: (scan index of every emitted pair, its target position), in the reference's emission order."""
    src = _xyz(source_local)
    T0 = np.ascontiguousarray(pose, dtype=np.float64).reshape(4, 4)
    cap = src.shape[0] * (7 if method == AVGICP else 1)
    idx = np.zeros(max(cap, 1), np.int32)
    tgt = np.zeros((max(cap, 1), 3))
    n = lib().ref_search_pairs(voxel_map._h, _f(src), src.shape[0], _d(T0), int(method), float(max_dist), _i(idx), _d(tgt), cap)
    return idx[:n], tgt[:n]


def voxel_downsample(xyz, voxel_size):
    """VoxelDownsample (voxel_hash_map.hpp:260-283): surviving input indices in the reference's (hash-table) order."""
    xyz = _xyz(xyz)
    idx = np.zeros(max(xyz.shape[0], 1), np.int32)
    n = lib().ref_voxel_downsample(_f(xyz), xyz.shape[0], float(voxel_size), _i(idx))
    return idx[:n]


# ---------------------------------------------------------------------------------------------------------------- EKF
def ekf_lib(fresh_statics=False):
    """oracle/_ref/libref_ekf.so.  ComplementaryKalmanFilter keeps its previous sample in function-static variables
    (ekf_algorithm.cpp:613-614), shared by every filter of the process: fresh_statics loads a private copy of the library."""
    import shutil
    import tempfile
    build()
    if not os.path.isfile(_SO_EKF):
        raise FileNotFoundError("oracle/_ref/libref_ekf.so is not built and the reference sources are not here")
    path = _SO_EKF
    if fresh_statics:
        fd, path = tempfile.mkstemp(suffix=".so", prefix="libref_ekf_")
        os.close(fd)
        shutil.copyfile(_SO_EKF, path)
    L = C.CDLL(path)
    if fresh_statics:
        os.unlink(path)  # stays mapped
    dp = C.POINTER(C.c_double)
    L.ref_ekf_create.restype = C.c_void_p
    L.ref_ekf_create.argtypes = [C.c_void_p]
    L.ref_ekf_destroy.argtypes = [C.c_void_p]
    L.ref_ekf_predict_imu.argtypes = [C.c_void_p, C.c_double, dp, dp]
    L.ref_ekf_update_pose.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_ekf_get_current_state.argtypes = [C.c_void_p, dp]
    L.ref_ekf_dump.argtypes = [C.c_void_p, C.c_void_p]
    return L


class EkfAlgorithm:
    """The reference's EkfAlgorithm (ekf_algorithm.hpp:79-290), itself; same call surface as oracle.EkfAlgorithm.
    `s` is refreshed from the filter's members after every call (the ComplementaryKalmanFilter statics are not reachable)."""

    def __init__(self, cfg, state_type, fresh_statics=True):
        self._L = ekf_lib(fresh_statics)
        self.cfg = cfg
        self.s = state_type()
        self._h = self._L.ref_ekf_create(C.byref(cfg))
        self._sync()

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_ekf_destroy(self._h)
            self._h = None

    def _sync(self):
        self._L.ref_ekf_dump(self._h, C.byref(self.s))

    def RunPredictionImu(self, t, gyro, acc):
        g = np.ascontiguousarray(gyro, dtype=np.float64)
        a = np.ascontiguousarray(acc, dtype=np.float64)
        r = bool(self._L.ref_ekf_predict_imu(self._h, float(t), _d(g), _d(a)))
        self._sync()
        return r

    def RunGnssUpdate(self, meas):
        r = bool(self._L.ref_ekf_update_pose(self._h, C.byref(meas)))
        self._sync()
        return r

    def GetCurrentState(self):
        ego = np.zeros(26)
        self._L.ref_ekf_get_current_state(self._h, _d(ego))
        return ego


# --------------------------------------------------------------------------------------------------------------- the node
LOCALIZATION_INI = """[common_variable]
lidar_type = velodyne
lidar_scan_time_end = {scan_time_end}
lidar_time_delay = {time_delay}
lidar_topic_name = /velodyne_points
imu_topic_name = /imu/data

[pcm_matching]
debug_print = 0
pcm_voxel_size = {voxel_size}
pcm_voxel_max_point = {voxel_max_point}
run_deskew = {run_deskew}
input_max_dist = {input_max_dist}
input_index_sampling = 1
input_voxel_ds_m = {input_voxel_ds_m}
icp_method = {icp_method}
voxel_search_method = 2
gicp_cov_search_dist = 0.4
max_thread = {max_thread}
max_iteration = {max_iteration}
max_search_dist = {max_search_dist}
lm_lambda = {lm_lambda}
icp_termination_threshold_m = {icp_termination_threshold_m}
min_overlap_ratio = {min_overlap_ratio}
max_fitness_score = {max_fitness_score}
use_radar_cov = 0
doppler_trans_lambda = 0.5
range_variance_m = 1.0
azimuth_variance_deg = 0.4
elevation_variance_deg = 0.4
"""
EKF_INI = """
[ekf_localization]
debug_print = 0
debug_imu_print = 0
imu_gravity = {imu_gravity}
imu_estimate_gravity = {imu_estimate_gravity}
imu_estimate_calibration = 0
use_zupt = 0
use_complementary_filter = {use_complementary_filter}
gps_type = 2
gnss_uncertainy_max_m = 1.0
use_gps = 0
use_imu = 1
use_can = 0
use_pcm_matching = 1
can_vel_scale_factor = 1.0
ekf_init_x_m = {ekf_init_x_m}
ekf_init_y_m = {ekf_init_y_m}
ekf_init_z_m = {ekf_init_z_m}
ekf_init_roll_deg = {ekf_init_roll_deg}
ekf_init_pitch_deg = {ekf_init_pitch_deg}
ekf_init_yaw_deg = {ekf_init_yaw_deg}
ekf_state_uncertainty_pos_m = {state_std_pos_m}
ekf_state_uncertainty_rot_deg = {state_std_rot_deg}
ekf_state_uncertainty_vel_mps = {state_std_vel_mps}
ekf_state_uncertainty_gyro_dps = 5.0
ekf_state_uncertainty_acc_mps = 100.0
ekf_imu_uncertainty_gyro_dps = {imu_std_gyro_dps}
ekf_imu_uncertainty_acc_mps = {imu_std_acc_mps}
ekf_imu_bias_cov_gyro = {imu_bias_cov_gyro}
ekf_imu_bias_cov_acc = {imu_bias_cov_acc}
ekf_gnss_min_cov_x_m = 0.2
ekf_gnss_min_cov_y_m = 0.2
ekf_gnss_min_cov_z_m = 0.7
ekf_gnss_min_cov_roll_deg = 0.0
ekf_gnss_min_cov_pitch_deg = 0.0
ekf_gnss_min_cov_yaw_deg = 0.0
ekf_can_meas_uncertainty_vel_mps = 2.0
ekf_can_meas_uncertainty_yaw_rate_deg = 10.0
ekf_bestvel_meas_uncertainty_vel_mps = 1.0
"""
EKF_CALIBRATION_INI = """[Rear To Imu]
transform_xyz_m = 0.0 0.0 0.0
rotation_rpy_deg = 0.0 0.0 0.0

[Rear To Gps]
transform_xyz_m = 0.0 0.0 0.0
rotation_rpy_deg = 0.0 0.0 0.0
"""
CALIBRATION_INI = """[Rear To Imu]
transform_xyz_m = 0.0 0.0 0.0
rotation_rpy_deg = {imu_rpy}

[Rear To Main LiDAR]
transform_xyz_m = {lidar_xyz}
rotation_rpy_deg = {lidar_rpy}
"""


class PcmMatchingNode:
    """The reference's ROS node class PcmMatching (pcm_matching.hpp:112-380), itself, with this wrapper as the middleware.
    Configuration goes through the node's own ini parser: the two files are written into a temporary $PWD/config."""

    def __init__(self, map_xyz, lidar_xyz=(0.0, 0.0, 0.0), lidar_rpy_deg=(0.0, 0.0, 0.0), imu_rpy_deg=(0.0, 0.0, 0.0), on_cuda=False, **kw):
        """on_cuda: the same unmodified node, but compiled against shim/registration_shim.hpp and linked with the product
        library — its VoxelHashMap / Registration are then the CUDA drop-in (needs a GPU to do anything but fail softly)."""
        import tempfile
        if on_cuda:
            so = build_node_on_cuda()
            if on_cuda == "host":  # the shim with ELM_SHIM_DEVICE = -1: host-only map, registration fails softly
                so = so.replace("libref_node_on_cuda.so", "libref_node_on_shim_host.so")
        else:
            build()
            so = _SO_NODE
        if not os.path.isfile(so):
            raise FileNotFoundError(so + " is not built and the reference sources are not here")
        d = dict(scan_time_end=0, time_delay=0.0, voxel_size=1.0, voxel_max_point=30, run_deskew=1, input_max_dist=100.0, input_voxel_ds_m=1.5,
                 icp_method=1, max_thread=1, max_iteration=10, max_search_dist=5.0, lm_lambda=0.5, icp_termination_threshold_m=0.02,
                 min_overlap_ratio=0.4, max_fitness_score=0.5)
        d.update(kw)
        self.cfg = d
        self._dir = tempfile.TemporaryDirectory(prefix="pcm_node_")
        os.makedirs(os.path.join(self._dir.name, "config"))
        with open(os.path.join(self._dir.name, "config", "localization.ini"), "w") as f:
            f.write(LOCALIZATION_INI.format(**d))
        with open(os.path.join(self._dir.name, "config", "calibration.ini"), "w") as f:
            f.write(CALIBRATION_INI.format(imu_rpy=" ".join(map(str, imu_rpy_deg)), lidar_xyz=" ".join(map(str, lidar_xyz)),
                                           lidar_rpy=" ".join(map(str, lidar_rpy_deg))))
        L = C.CDLL(so)
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.ref_node_create.restype = C.c_void_p
        L.ref_node_create.argtypes = [C.c_char_p, fp, C.c_size_t]
        L.ref_node_destroy.argtypes = [C.c_void_p]
        L.ref_node_map_points.restype = C.c_size_t
        L.ref_node_map_points.argtypes = [C.c_void_p]
        L.ref_node_imu.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.ref_node_odom.argtypes = [C.c_void_p, C.c_double, dp, dp, dp, dp]
        L.ref_node_cloud.restype = C.c_size_t
        L.ref_node_cloud.argtypes = [C.c_void_p, C.c_double, fp, fp, C.c_size_t]
        L.ref_node_last_pcm_odom.argtypes = [dp, dp, dp, dp]
        L.ref_node_last_registered_cloud.restype = C.c_size_t
        L.ref_node_last_registered_cloud.argtypes = [fp, C.c_size_t]
        L.ref_node_filter_by_distance.restype = C.c_size_t
        L.ref_node_filter_by_distance.argtypes = [C.c_void_p, fp, C.c_size_t, ip]
        L.ref_node_deskew.argtypes = [C.c_void_p, C.c_double, fp, fp, C.c_size_t]
        L.ref_node_undistorted.restype = C.c_size_t
        L.ref_node_undistorted.argtypes = [C.c_void_p, fp, C.c_size_t]
        L.ref_node_deskew_tables.argtypes = [C.c_void_p, dp, dp, dp, dp, ip, fp, dp]
        L.ref_node_interpolated_pose.argtypes = [C.c_void_p, C.c_double, fp]
        L.ref_node_shape_covariance.argtypes = [C.c_void_p, dp, dp, C.c_double, dp]
        L.ref_node_init_map_cloud.restype = C.c_size_t
        L.ref_node_init_map_cloud.argtypes = [fp, C.c_size_t]
        L.ref_node_init_markers.restype = C.c_size_t
        L.ref_node_init_markers.argtypes = [dp, C.c_size_t]
        L.ref_node_initial_pose.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp]
        self._L = L
        m = _xyz(map_xyz)
        self._h = L.ref_node_create(self._dir.name.encode(), _f(m), m.shape[0])

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_node_destroy(self._h)
            self._h = None

    def map_points(self):
        return self._L.ref_node_map_points(self._h)

    def init_publications(self):
        """what Init() published: (voxel-map cloud [P, 3], covariance markers [V, 6] = position + scale); call before any
        other node of the same library is created"""
        n = self.map_points()
        cloud = np.zeros((max(n, 1), 3), np.float32)
        nc = self._L.ref_node_init_map_cloud(_f(cloud), cloud.shape[0])
        mk = np.zeros((max(n, 1), 6))
        nm = self._L.ref_node_init_markers(_d(mk), mk.shape[0])
        return cloud[:nc], mk[:nm]

    def initial_pose(self, x, y, yaw, z=0.0):
        """CallbackInitialPose; returns dict(pos, quat_wxyz) of the published /app/loc/pcm_init_odom or None"""
        pos, q = np.zeros(3), np.zeros(4)
        if not self._L.ref_node_initial_pose(self._h, float(x), float(y), float(z), float(yaw), _d(pos), _d(q)):
            return None
        return dict(pos=pos, quat_wxyz=np.array([q[3], q[0], q[1], q[2]]))

    def imu(self, t, gyro, acc):
        g, a = np.ascontiguousarray(gyro, dtype=np.float64), np.ascontiguousarray(acc, dtype=np.float64)
        self._L.ref_node_imu(self._h, float(t), _d(g), _d(a))

    def odom(self, t, pos, quat_wxyz, lin=(0, 0, 0), ang=(0, 0, 0)):
        q = np.asarray(quat_wxyz, dtype=np.float64)
        p, qx = np.ascontiguousarray(pos, dtype=np.float64), np.ascontiguousarray([q[1], q[2], q[3], q[0]], dtype=np.float64)
        li, an = np.ascontiguousarray(lin, dtype=np.float64), np.ascontiguousarray(ang, dtype=np.float64)
        self._L.ref_node_odom(self._h, float(t), _d(p), _d(qx), _d(li), _d(an))

    def cloud(self, stamp, xyz, rel_time):
        """CallbackPointCloud; returns the published pcm odometry of THIS call (dict) or None"""
        x, rt = _xyz(xyz), np.ascontiguousarray(rel_time, dtype=np.float32)
        before = getattr(self, "_n_odom", 0)
        self._n_odom = self._L.ref_node_cloud(self._h, float(stamp), _f(x), _f(rt), x.shape[0])
        if self._n_odom == before:
            return None
        st, pos, q, cov = np.zeros(1), np.zeros(3), np.zeros(4), np.zeros((6, 6))
        self._L.ref_node_last_pcm_odom(_d(st), _d(pos), _d(q), _d(cov))
        reg = np.zeros((x.shape[0], 3), np.float32)
        n = self._L.ref_node_last_registered_cloud(_f(reg), x.shape[0])
        return dict(stamp=float(st[0]), pos=pos, quat_wxyz=np.array([q[3], q[0], q[1], q[2]]), cov=cov, registered_world=reg[:n])

    def filter_by_distance(self, xyz):
        x = _xyz(xyz)
        idx = np.zeros(max(1, x.shape[0]), np.int32)
        n = self._L.ref_node_filter_by_distance(self._h, _f(x), x.shape[0], _i(idx))
        return idx[:n]

    def deskew(self, stamp, xyz, rel_time):
        """DeskewPointCloud; returns (ok, undistorted cloud, tables in the layout of oracle.deskew_tables)"""
        x, rt = _xyz(xyz), np.ascontiguousarray(rel_time, dtype=np.float32)
        ok = bool(self._L.ref_node_deskew(self._h, float(stamp), _f(x), _f(rt), x.shape[0]))
        out = np.zeros_like(x)
        n = self._L.ref_node_undistorted(self._h, _f(out), x.shape[0])
        t = dict(imu_time=np.zeros(2000), imu_rot_x=np.zeros(2000), imu_rot_y=np.zeros(2000), imu_rot_z=np.zeros(2000))
        mi, mf, tm = np.zeros(3, np.int32), np.zeros(3, np.float32), np.zeros(2)
        self._L.ref_node_deskew_tables(self._h, _d(t["imu_time"]), _d(t["imu_rot_x"]), _d(t["imu_rot_y"]), _d(t["imu_rot_z"]), _i(mi), _f(mf), _d(tm))
        t.update(imu_pointer_cur=int(mi[0]), imu_available=bool(mi[1]), odom_available=bool(mi[2]), odom_incre=mf.copy(),
                 time_scan_cur=float(tm[0]), time_scan_end=float(tm[1]))
        return ok, out[:n], t

    def interpolated_pose(self, t):
        T = np.zeros((4, 4), np.float32)
        ok = bool(self._L.ref_node_interpolated_pose(self._h, float(t), _f(T)))
        return ok, T

    def shape_covariance(self, pose, local_cov, icp_pose_std_m):
        T = np.ascontiguousarray(pose, dtype=np.float64).reshape(4, 4)
        lc = np.ascontiguousarray(local_cov, dtype=np.float64).reshape(6, 6)
        out = np.zeros((6, 6))
        self._L.ref_node_shape_covariance(self._h, _d(T), _d(lc), float(icp_pose_std_m), _d(out))
        return out


def _fresh_cdll(path):
    """a private copy of a library: its function-static variables start from scratch"""
    import shutil
    import tempfile
    fd, tmp = tempfile.mkstemp(suffix=".so", prefix="libref_")
    os.close(fd)
    shutil.copyfile(path, tmp)
    L = C.CDLL(tmp)
    os.unlink(tmp)
    return L


class EkfLocalizationNode:
    """The reference's EKF ROS node class EkfLocalization (ekf_localization.hpp), itself; configured through its own ini
    parser from the field names of elimaloc_b200.ekf.make_ekf_config (`ekf_cfg` is that ctypes structure)."""

    def __init__(self, ekf_cfg):
        import tempfile
        build()
        if not os.path.isfile(_SO_EKFNODE):
            raise FileNotFoundError("oracle/_ref/libref_ekfnode.so is not built and the reference sources are not here")
        d = {name: getattr(ekf_cfg, name) for name, _ in ekf_cfg._fields_ if name != "reserved"}
        self._dir = tempfile.TemporaryDirectory(prefix="ekf_node_")
        os.makedirs(os.path.join(self._dir.name, "config"))
        with open(os.path.join(self._dir.name, "config", "localization.ini"), "w") as f:
            f.write("[common_variable]\ncan_topic_name = /can\nimu_topic_name = /imu/data\nnavsatfix_topic_name = /gps/fix\nprojection_mode = Cartesian\n")
            f.write(EKF_INI.format(**{k: repr(v) if isinstance(v, float) else v for k, v in d.items()}))
        with open(os.path.join(self._dir.name, "config", "calibration.ini"), "w") as f:
            f.write(EKF_CALIBRATION_INI)
        L = _fresh_cdll(_SO_EKFNODE)
        dp = C.POINTER(C.c_double)
        L.ref_ekfnode_create.restype = C.c_void_p
        L.ref_ekfnode_create.argtypes = [C.c_char_p]
        L.ref_ekfnode_destroy.argtypes = [C.c_void_p]
        L.ref_ekfnode_imu.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.ref_ekfnode_pcm_odom.argtypes = [C.c_void_p, C.c_double, dp, dp, dp]
        L.ref_ekfnode_pcm_init_odom.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.ref_ekfnode_last_odom.argtypes = [dp, dp, dp, dp, dp]
        L.ref_ekfnode_filter_pose.argtypes = [C.c_void_p, dp, dp]
        L.ref_ekfnode_time_compensate.argtypes = [C.c_void_p, C.c_double, dp, dp, dp, dp, dp]
        L.ref_ekfnode_state_queue.restype = C.c_size_t
        L.ref_ekfnode_state_queue.argtypes = [C.c_void_p, dp, C.c_size_t]
        self._L = L
        self._h = L.ref_ekfnode_create(self._dir.name.encode())

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_ekfnode_destroy(self._h)
            self._h = None

    def imu(self, t, gyro, acc):
        """CallbackImu; returns the /app/loc/ekf_pose_odom it published: dict(t, pos, quat wxyz, vel_local, rate)"""
        g, a = np.ascontiguousarray(gyro, dtype=np.float64), np.ascontiguousarray(acc, dtype=np.float64)
        self._L.ref_ekfnode_imu(self._h, float(t), _d(g), _d(a))
        st, pos, q, lin, ang = np.zeros(1), np.zeros(3), np.zeros(4), np.zeros(3), np.zeros(3)
        if not self._L.ref_ekfnode_last_odom(_d(st), _d(pos), _d(q), _d(lin), _d(ang)):
            return None
        return dict(t=float(st[0]), pos=pos, quat=np.array([q[3], q[0], q[1], q[2]]), vel_local=lin, rate=ang)

    def pcm_odom(self, t, pos, quat_wxyz, cov66):
        q = np.asarray(quat_wxyz, dtype=np.float64)
        p, qx = np.ascontiguousarray(pos, dtype=np.float64), np.ascontiguousarray([q[1], q[2], q[3], q[0]], dtype=np.float64)
        c = np.ascontiguousarray(cov66, dtype=np.float64).reshape(36)
        self._L.ref_ekfnode_pcm_odom(self._h, float(t), _d(p), _d(qx), _d(c))

    def pcm_init_odom(self, t, pos, quat_wxyz):
        q = np.asarray(quat_wxyz, dtype=np.float64)
        p, qx = np.ascontiguousarray(pos, dtype=np.float64), np.ascontiguousarray([q[1], q[2], q[3], q[0]], dtype=np.float64)
        self._L.ref_ekfnode_pcm_init_odom(self._h, float(t), _d(p), _d(qx))

    def filter_pose(self):
        pos, q = np.zeros(3), np.zeros(4)
        self._L.ref_ekfnode_filter_pose(self._h, _d(pos), _d(q))
        return np.concatenate([pos, q])

    def time_compensate(self, t, pos, quat_wxyz):
        p, q = np.ascontiguousarray(pos, dtype=np.float64), np.ascontiguousarray(quat_wxyz, dtype=np.float64)
        to, po, qo = np.zeros(1), np.zeros(3), np.zeros(4)
        if not self._L.ref_ekfnode_time_compensate(self._h, float(t), _d(p), _d(q), _d(to), _d(po), _d(qo)):
            return None
        return dict(t=float(to[0]), pos=po, quat=qo)

    def state_queue(self):
        rows = np.zeros((1000, 7))
        n = self._L.ref_ekfnode_state_queue(self._h, _d(rows), 1000)
        return rows[:n]
