"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

ctypes front-end of oracle/_ref/libref.so: the REFERENCE's own registration.cpp + voxel_hash_map.{hpp,cpp}, compiled
unmodified from /root/reference against the stand-in third-party headers in oracle/ref_build/stubs (Eigen3, oneTBB and PCL are
absent from this image; see stubs/mini_eigen.hpp for what the stand-in does and does not preserve), and of
oracle/_ref/libref_ekf.so: the reference's own ekf_algorithm.cpp built the same way (plus ref_build/ros_stubs).  They are the
pin of the oracle: tests/test_reference_build.py and tests/test_reference_build_ekf.py run the oracle and these libraries on
the same seeded inputs.

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
_LIB = None


def sources_present():
    return os.path.isfile(os.path.join(_PCM, "src", "registration.cpp"))


def available():
    return os.path.isfile(_SO) or sources_present()


def ekf_available():
    return os.path.isfile(_SO_EKF) or sources_present()


def build(force=False):
    """Compile the reference's two translation units where they lie (never copied) + ref_build/ref_capi.cpp."""
    if sources_present():  # make decides whether anything is stale
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []) + ["_ref/libref.so", "REFERENCE_ROOT=" + REFERENCE_ROOT])
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []) + ["_ref/libref_ekf.so", "REFERENCE_ROOT=" + REFERENCE_ROOT])
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
