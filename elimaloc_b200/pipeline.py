"""Host-side mirror of the per-scan chain of PcmMatching::CallbackPointCloud (pcm_matching.cpp:198-299) over the C ABI:
deskew tables (ImuDeskewInfo / OdomDeskewInfo, :533-729) and the device-resident scan pipeline (filter -> deskew -> down-sampling
-> RunRegister -> EKF update) of include/elimaloc_b200.h."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import DeskewTables, ImuQueue, OdomQueue, ScanPipelineConfig, ScanResult, check, lib


def _d(a):
    return a.ctypes.data_as(_capi._dp)


class Queues:
    """deq_imu_ / deq_odom_ of the node as contiguous arrays (kept alive for the duration of a call)."""

    def __init__(self, imu_stamp, imu_gyro, odom_stamp, odom_pos, odom_quat_xyzw, odom_lin_vel, odom_ang_vel):
        c = lambda a, w: np.ascontiguousarray(a, dtype=np.float64).reshape(-1, w) if w > 1 else np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        self.keep = [c(imu_stamp, 1), c(imu_gyro, 3), c(odom_stamp, 1), c(odom_pos, 3), c(odom_quat_xyzw, 4), c(odom_lin_vel, 3), c(odom_ang_vel, 3)]
        k = self.keep
        assert len(k[0]) == len(k[1]) and all(len(k[2]) == len(a) for a in k[3:])
        self.imu = ImuQueue(_d(k[0]), _d(k[1]), len(k[0]))
        self.odom = OdomQueue(_d(k[2]), _d(k[3]), _d(k[4]), _d(k[5]), _d(k[6]), len(k[2]))


def build_deskew_tables(queues, time_scan_cur, time_scan_end):
    """ImuDeskewInfo + OdomDeskewInfo -> dict in the layout Registration.DeskewPoints takes (+ imu_drop / odom_drop)."""
    storage = np.zeros(4 * _capi.ELM_IMU_QUEUE_LENGTH)
    t = DeskewTables()
    t.time_scan_cur, t.time_scan_end = float(time_scan_cur), float(time_scan_end)
    di, do = C.c_size_t(0), C.c_size_t(0)
    check(lib().elm_deskew_build_tables(C.byref(queues.imu), C.byref(queues.odom), C.byref(t), _d(storage), C.byref(di), C.byref(do)))
    L = _capi.ELM_IMU_QUEUE_LENGTH
    return dict(imu_time=storage[:L].copy(), imu_rot_x=storage[L:2 * L].copy(), imu_rot_y=storage[2 * L:3 * L].copy(), imu_rot_z=storage[3 * L:].copy(),
                imu_pointer_cur=int(t.imu_pointer_cur), imu_available=bool(t.imu_available), odom_available=bool(t.odom_available),
                odom_incre=np.array([t.odom_incre_x, t.odom_incre_y, t.odom_incre_z], np.float32), time_scan_cur=float(t.time_scan_cur),
                time_scan_end=float(t.time_scan_end), imu_drop=int(di.value), odom_drop=int(do.value))


class ScanPipeline:
    def __init__(self, registration, input_max_dist=0.0, input_voxel_ds_m=0.0, run_deskew=True, lidar_scan_time_end=False, tf_ego_to_lidar=None):
        cfg = ScanPipelineConfig()
        cfg.input_max_dist, cfg.input_voxel_ds_m = float(input_max_dist), float(input_voxel_ds_m)
        cfg.run_deskew, cfg.lidar_scan_time_end = int(bool(run_deskew)), int(bool(lidar_scan_time_end))
        T = np.eye(4) if tf_ego_to_lidar is None else np.ascontiguousarray(tf_ego_to_lidar, dtype=np.float64).reshape(4, 4)
        cfg.tf_ego_to_lidar[:] = [float(v) for v in T.reshape(16)]
        self._h = C.c_void_p()
        self._reg = registration  # (keeps the registration alive)
        check(lib().elm_scan_pipeline_create(C.byref(self._h), registration._h, C.byref(cfg)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().elm_scan_pipeline_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown: the module globals may already be gone)
                pass

    def deskew(self, xyz, point_time, stamp, queues):
        """-> (deskew_ok, time_scan_cur, time_scan_end); enqueues upload + distance filter + deskew."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        pt = np.ascontiguousarray(point_time, dtype=np.float32).reshape(-1)
        tc, te, ok = C.c_double(0.0), C.c_double(0.0), C.c_int32(0)
        check(lib().elm_scan_pipeline_deskew(self._h, xyz.ctypes.data_as(_capi._fp), pt.ctypes.data_as(_capi._fp), xyz.shape[0], float(stamp),
                                             C.byref(queues.imu), C.byref(queues.odom), C.byref(tc), C.byref(te), C.byref(ok)))
        return bool(ok.value), float(tc.value), float(te.value)

    def register(self, voxel_map, sync_lidar_pose, cfg):
        T = np.ascontiguousarray(sync_lidar_pose, dtype=np.float64).reshape(4, 4)
        check(lib().elm_scan_pipeline_register(self._h, voxel_map._h, _d(T), C.byref(cfg)))

    def ekf_update(self, ekf):
        check(lib().elm_scan_pipeline_ekf_update(self._h, ekf._h))

    def fetch(self):
        r = ScanResult()
        check(lib().elm_scan_pipeline_fetch(self._h, C.byref(r)))
        return dict(T_lidar=np.array(r.T_lidar[:]).reshape(4, 4), T_ego=np.array(r.T_ego[:]).reshape(4, 4), fitness_score=float(r.fitness_score),
                    local_cov=np.array(r.local_cov[:]).reshape(6, 6), pose_cov=np.array(r.pose_cov[:]).reshape(6, 6), is_success=bool(r.is_success),
                    iterations=int(r.iterations), n_raw=int(r.n_raw), n_after_filter=int(r.n_after_filter), n_registered=int(r.n_registered),
                    time_scan_cur=float(r.time_scan_cur), time_scan_end=float(r.time_scan_end))
