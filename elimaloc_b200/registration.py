"""Host-side mirror of the reference's registration interface over the C ABI.

Names, argument meaning and failure behaviour follow the reference
(src/app/localization/pcm_matching/include/{registration,voxel_hash_map}.hpp):

    VoxelHashMap.Init / AddPoints / CalVoxelCovAll / CalPointCovAll / Empty / Pointcloud / Covariances
    Registration.Init / RunRegister

so the parity tests read like tests of the reference.  All compute happens in libelimaloc_b200.so on the GPU."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import AVGICP, GICP, P2P, VGICP, RegConfig, check, lib  # noqa: F401


def RegistrationConfig(**kw):
    """RegistrationConfig (registration.hpp:62-85); defaults = config/localization.ini:80-109 of the reference."""
    d = dict(icp_method=GICP, max_iteration=10, max_thread=10, use_radar_cov=0, debug_print=0, reserved0=0,
             max_search_dist=5.0, lm_lambda=0.5, icp_termination_threshold_m=0.02, min_overlap_ratio=0.4,
             max_fitness_score=0.5, range_variance_m=1.0, azimuth_variance_deg=0.4, elevation_variance_deg=0.4)
    d.update(kw)
    return RegConfig(**d)


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _xyz(a):
    return np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)


def _pose(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(4, 4)


class VoxelHashMap:
    """voxel_hash_map.hpp:89-335.  device=-1 builds a host-only map (builder tests on a machine without a GPU)."""

    def __init__(self, voxel_size=1.0, max_points_per_voxel=30, device=0):
        self._h = C.c_void_p()
        self.device = device
        self.Init(voxel_size, max_points_per_voxel)

    def Init(self, voxel_size, max_points_per_voxel):
        if self._h:
            lib().elm_map_destroy(self._h)
            self._h = C.c_void_p()
        check(lib().elm_map_create(C.byref(self._h), float(voxel_size), int(max_points_per_voxel), int(self.device)))
        self.voxel_size_ = float(voxel_size)
        self.max_points_per_voxel_ = int(max_points_per_voxel)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().elm_map_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown: the module globals may already be gone)
                pass

    def AddPoints(self, xyz):
        xyz = _xyz(xyz)
        check(lib().elm_map_add_points(self._h, _f(xyz), xyz.shape[0]))

    def set_gpu_build(self, enable):
        """AddPoints / CalVoxelCovAll / CalPointCovAll on the GPU (default for a map on a device) or on the host; same results."""
        check(lib().elm_map_set_gpu_build(self._h, int(bool(enable))))

    def build_times(self):
        """milliseconds the last AddPoints / CalVoxelCovAll / CalPointCovAll spent in the builder proper"""
        ms = np.zeros(3)
        check(lib().elm_map_build_times(self._h, _d(ms)))
        return dict(add_points_ms=float(ms[0]), voxel_cov_ms=float(ms[1]), point_cov_ms=float(ms[2]))

    def AddPointsFromPcd(self, path):
        """loadPCDFile + AddPoints (pcm_matching.cpp:69-88); returns the number of points read"""
        n = C.c_size_t(0)
        check(lib().elm_map_add_points_pcd(self._h, str(path).encode(), C.byref(n)))
        return int(n.value)

    def FindGroundHeight(self, position_xy):
        """voxel_hash_map.hpp:285-322: (found, ground_z)"""
        z, found = C.c_double(0.0), C.c_int32(0)
        check(lib().elm_map_find_ground_height(self._h, float(position_xy[0]), float(position_xy[1]), C.byref(z), C.byref(found)))
        return bool(found.value), float(z.value)

    def CalVoxelCovAll(self):
        check(lib().elm_map_cal_voxel_cov(self._h))

    def CalPointCovAll(self, d_search_dist):
        check(lib().elm_map_cal_point_cov(self._h, float(d_search_dist)))

    def Empty(self):
        return bool(lib().elm_map_empty(self._h))

    def num_voxels(self):
        return lib().elm_map_num_voxels(self._h)

    def num_points(self):
        return lib().elm_map_num_points(self._h)

    def Save(self, path):
        """Write the built map (points, voxel table, directory, covariances) to `path` (elm_map_save)."""
        check(lib().elm_map_save(self._h, str(path).encode()))

    @classmethod
    def Load(cls, path, device=0):
        """A map restored from a file written by Save, uploaded to `device` (elm_map_load)."""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        self.device = device
        check(lib().elm_map_load(C.byref(self._h), str(path).encode(), int(device)))
        self.voxel_size_ = self.max_points_per_voxel_ = None
        return self

    def directory_check(self):
        """(centre keys stored, table slots, mismatches) of the neighbourhood directory; mismatches must be 0."""
        e, s, m = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        check(lib().elm_map_directory_check(self._h, C.byref(e), C.byref(s), C.byref(m)))
        return int(e.value), int(s.value), int(m.value)

    def export(self, voxel_cov=False, point_cov=False):
        V, P = self.num_voxels(), self.num_points()
        out = dict(keys=np.zeros((V, 3), np.int32), counts=np.zeros(V, np.int32), pxyz=np.zeros((P, 3), np.float32))
        vmean = vcov = pmean = pcov = None
        if voxel_cov:
            out["vmean"], out["vcov"] = np.zeros((V, 3)), np.zeros((V, 3, 3))
            vmean, vcov = _d(out["vmean"]), _d(out["vcov"])
        if point_cov:
            out["pmean"], out["pcov"] = np.zeros((P, 3)), np.zeros((P, 3, 3))
            pmean, pcov = _d(out["pmean"]), _d(out["pcov"])
        check(lib().elm_map_export(self._h, _i(out["keys"]), _i(out["counts"]), vmean, vcov, _f(out["pxyz"]), pmean, pcov))
        return out

    def Pointcloud(self):
        """voxel_hash_map.cpp:245-255 (canonical order instead of unordered_map iteration order)."""
        return self.export()["pxyz"]

    def Covariances(self):
        """voxel_hash_map.cpp:257-265: covariances of voxels holding more than two points."""
        e = self.export(voxel_cov=True)
        keep = e["counts"] > 2
        return e["vmean"][keep], e["vcov"][keep]


def read_pcd_xyz(path):
    """x, y, z of a PCD v0.7 file (ascii / binary / binary_compressed) as float32 (n, 3); returns (xyz, dropped non-finite)."""
    n, dropped = C.c_size_t(0), C.c_size_t(0)
    check(lib().elm_pcd_read_xyz(str(path).encode(), None, 0, C.byref(n), C.byref(dropped)))
    xyz = np.zeros((int(n.value), 3), np.float32)
    check(lib().elm_pcd_read_xyz(str(path).encode(), _f(xyz), int(n.value), C.byref(n), C.byref(dropped)))
    return xyz, int(dropped.value)


def shape_pcm_covariance(R_ego, local_cov, icp_pose_std_m, cov36=None):
    """PcmMatching::PublishPcmOdom's covariance shaping (pcm_matching.cpp:1082-1098) -> 6x6 row-major pose covariance."""
    R = np.ascontiguousarray(R_ego, dtype=np.float64).reshape(3, 3)
    lc = np.ascontiguousarray(local_cov, dtype=np.float64).reshape(6, 6)
    out = np.zeros((6, 6)) if cov36 is None else np.ascontiguousarray(cov36, dtype=np.float64).reshape(6, 6).copy()
    check(lib().elm_shape_pcm_covariance(_d(R), _d(lc), float(icp_pose_std_m), _d(out)))
    return out


class Registration:
    """registration.hpp:101-230.  Holds the device scratch of the ICP loop and d_fitness_score_."""

    def __init__(self, config=None, device=0, stream=None):
        self._h = C.c_void_p()
        self.device = device
        check(lib().elm_registration_create(C.byref(self._h), int(device), C.c_void_p(stream) if stream else None))
        self.config_ = config

    def Init(self, config):
        self.config_ = config

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().elm_registration_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown: the module globals may already be gone)
                pass

    def RunRegister(self, source_local, voxel_map, initial_guess, m_config=None, fitness_score=0.0):
        """Returns (pose 4x4, is_success, fitness_score, local_cov 6x6); fitness_score is passed through
        untouched on failure exactly like the reference's out-parameter (registration.cpp:415)."""
        cfg = m_config if m_config is not None else self.config_
        src = _xyz(source_local)
        T0 = _pose(initial_guess)
        T = np.zeros((4, 4))
        ok = np.zeros(1, np.int32)
        fit = np.array([fitness_score], np.float64)
        cov = np.zeros((6, 6))
        check(lib().elm_run_register(self._h, voxel_map._h, _f(src), src.shape[0], _d(T0), C.byref(cfg), _d(T), _i(ok),
                                     _d(fit), _d(cov)))
        return T, bool(ok[0]), float(fit[0]), cov

    # ---- device-resident variant (what bench.py times as `value`) ----
    def enqueue(self, d_src_ptr, n, voxel_map, initial_guess, cfg):
        T0 = _pose(initial_guess)
        check(lib().elm_register_enqueue(self._h, voxel_map._h, C.c_void_p(d_src_ptr), int(n), _d(T0), C.byref(cfg)))

    def fetch(self, fitness_score=0.0):
        T = np.zeros((4, 4))
        ok = np.zeros(1, np.int32)
        fit = np.array([fitness_score], np.float64)
        cov = np.zeros((6, 6))
        it = np.zeros(1, np.int32)
        check(lib().elm_register_fetch(self._h, _d(T), _i(ok), _d(fit), _d(cov), _i(it)))
        return T, bool(ok[0]), float(fit[0]), cov, int(it[0])

    def launch_count(self):
        n = C.c_int64(0)
        check(lib().elm_registration_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def set_profiling(self, enable):
        check(lib().elm_registration_set_profiling(self._h, int(bool(enable))))

    def profile(self):
        """(ms in the search kernel, ms in the accumulate kernel, iterations timed) since set_profiling(True)."""
        a, b = C.c_double(0.0), C.c_double(0.0)
        n = C.c_int64(0)
        check(lib().elm_registration_profile(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return float(a.value), float(b.value), int(n.value)

    def profile_by_kind(self):
        """{kind: (ms in the first kernel, ms in the second kernel, iterations)} for kind in cold / first_warm / warm."""
        ms = np.zeros(6)
        n = (C.c_int64 * 3)()
        check(lib().elm_registration_profile_by_kind(self._h, _d(ms), n))
        return {k: (float(ms[2 * i]), float(ms[2 * i + 1]), int(n[i])) for i, k in enumerate(("cold", "first_warm", "warm"))}

    def set_stats(self, enable):
        check(lib().elm_registration_set_stats(self._h, int(bool(enable))))

    def stats(self):
        """(map points visited by the P2P/GICP search, queries searched) since set_stats()."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        check(lib().elm_registration_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def stats_raw(self):
        """All 32 device counters (see elm_registration_stats_raw)."""
        buf = (C.c_uint64 * 32)()
        check(lib().elm_registration_stats_raw(self._h, buf))
        return [int(v) for v in buf]

    def set_fused(self, enable):
        """P2P / GICP: one fused kernel per iteration (default) or search + accumulate as two launches."""
        check(lib().elm_registration_set_fused(self._h, int(bool(enable))))

    def set_binning(self, enable):
        """Search the scan in spatially binned order (default) or in the caller's order; results are identical."""
        check(lib().elm_registration_set_binning(self._h, int(bool(enable))))

    def set_exhaustive(self, exhaustive):
        """True: visit all 27 voxels like the reference; False (default): exact pruning."""
        check(lib().elm_registration_set_exhaustive(self._h, int(bool(exhaustive))))

    # ---- test hooks ----
    def linearize(self, source_local, voxel_map, pose, cfg):
        src = _xyz(source_local)
        T0 = _pose(pose)
        JTJ, JTr, res = np.zeros((6, 6)), np.zeros(6), np.zeros(1)
        nc = C.c_int64(0)
        check(lib().elm_linearize(self._h, voxel_map._h, _f(src), src.shape[0], _d(T0), C.byref(cfg), _d(JTJ), _d(JTr),
                                  _d(res), C.byref(nc)))
        return dict(JTJ=JTJ, JTr=JTr, residual_sum=float(res[0]), n_corr=int(nc.value))

    def correspondences(self, source_local, voxel_map, pose, method, max_dist):
        src = _xyz(source_local)
        T0 = _pose(pose)
        K = 7 if method == AVGICP else 1
        cnt = np.zeros(src.shape[0], np.int32)
        tgt = np.zeros((src.shape[0], K, 3))
        check(lib().elm_correspondences(self._h, voxel_map._h, _f(src), src.shape[0], _d(T0), int(method), float(max_dist),
                                        _i(cnt), _d(tgt)))
        return cnt, tgt

    def correspondences_sequence(self, source_local, voxel_map, poses, method, max_dist):
        """Correspondences at poses[-1] after searching poses[0], poses[1], ... in turn the way the ICP loop does
        (cold search first, warm-started searches after it)."""
        src = _xyz(source_local)
        Ts = np.ascontiguousarray(np.stack([_pose(T) for T in poses]), dtype=np.float64)
        K = 7 if method == AVGICP else 1
        cnt = np.zeros(src.shape[0], np.int32)
        tgt = np.zeros((src.shape[0], K, 3))
        check(lib().elm_correspondences_sequence(self._h, voxel_map._h, _f(src), src.shape[0], _d(Ts), len(poses), int(method),
                                                 float(max_dist), _i(cnt), _d(tgt)))
        return cnt, tgt

    def set_warm_start(self, enable):
        """P2P / GICP: iterations after the first start their search from the previous match (default on; same results)."""
        check(lib().elm_registration_set_warm_start(self._h, int(bool(enable))))

    # ---- deskew (PcmMatching::DeskewPointCloud's per-point loop, pcm_matching.cpp:499-511, 780-824) ----
    @staticmethod
    def _deskew_tables(t):
        keep = [np.ascontiguousarray(t[k], dtype=np.float64) for k in ("imu_time", "imu_rot_x", "imu_rot_y", "imu_rot_z")]
        inc = np.asarray(t["odom_incre"], dtype=np.float32)
        st = _capi.DeskewTables(_d(keep[0]), _d(keep[1]), _d(keep[2]), _d(keep[3]), int(t["imu_pointer_cur"]), int(t["imu_available"]),
                                int(t["odom_available"]), 0, float(inc[0]), float(inc[1]), float(inc[2]), 0.0,
                                float(t["time_scan_cur"]), float(t["time_scan_end"]))
        return st, keep

    def DeskewPoints(self, xyz, rel_time, tables):
        """tables: dict with imu_time, imu_rot_{x,y,z}, imu_pointer_cur, imu_available, odom_available, odom_incre[3],
        time_scan_cur, time_scan_end (the node's member variables of the same meaning)."""
        xyz = _xyz(xyz)
        rt = np.ascontiguousarray(rel_time, dtype=np.float32)
        out = np.zeros_like(xyz)
        st, keep = self._deskew_tables(tables)
        check(lib().elm_deskew_points(self._h, _f(xyz), _f(rt), xyz.shape[0], C.byref(st), _f(out)))
        return out

    def deskew_device(self, d_xyz_ptr, d_time_ptr, n, tables, d_out_ptr):
        st, keep = self._deskew_tables(tables)
        check(lib().elm_deskew_points_device(self._h, C.c_void_p(d_xyz_ptr), C.c_void_p(d_time_ptr), int(n), C.byref(st),
                                             C.c_void_p(d_out_ptr)))

    # ---- scan pre-processing (FilterPointsByDistance + VoxelDownsample, pcm_matching.cpp:451-465, voxel_hash_map.hpp:260-283) ----
    def PreprocessScan(self, xyz, max_dist=0.0, voxel_size=0.0, aux=None):
        """Returns (xyz of the survivors in input order, their aux values or None, their input indices)."""
        xyz = _xyz(xyz)
        n = xyz.shape[0]
        out = np.zeros_like(xyz)
        idx = np.zeros(n, np.int32)
        a = np.ascontiguousarray(aux, dtype=np.float32) if aux is not None else None
        ao = np.zeros(n, np.float32) if aux is not None else None
        m = C.c_size_t(0)
        check(lib().elm_scan_preprocess(self._h, _f(xyz), _f(a) if a is not None else None, n, float(max_dist), float(voxel_size), _f(out),
                                        _f(ao) if ao is not None else None, _i(idx), C.byref(m)))
        k = int(m.value)
        return out[:k], (ao[:k] if ao is not None else None), idx[:k]

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * 128)()
        check(lib().elm_comm_unique_id(buf))
        return bytes(buf)

    def set_comm(self, unique_id, rank, world_size):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(lib().elm_registration_set_comm(self._h, buf, int(rank), int(world_size)))

    def peer_export(self):
        """64-byte CUDA IPC handle of this rank's accumulator mailbox (elm_registration_peer_export)."""
        buf = (C.c_uint8 * 64)()
        check(lib().elm_registration_peer_export(self._h, buf))
        return bytes(buf)

    def peer_attach(self, handles, rank, world_size):
        """handles: world_size handles in rank order (this rank's own entry is ignored)."""
        flat = b"".join(handles)
        assert len(flat) == 64 * world_size
        buf = (C.c_uint8 * len(flat)).from_buffer_copy(flat)
        check(lib().elm_registration_peer_attach(self._h, buf, int(rank), int(world_size)))

    def peer_detach(self):
        check(lib().elm_registration_peer_detach(self._h))

    def peer_setup(self, dist):
        """Export / all-gather / attach over an initialised torch.distributed process group, then barrier."""
        rank, world = dist.get_rank(), dist.get_world_size()
        handles = [None] * world
        dist.all_gather_object(handles, self.peer_export())
        err = None
        try:
            self.peer_attach(handles, rank, world)
        except _capi.ElmError as exc:  # keep the collective call sequence identical on every rank, then report
            err = exc
        dist.barrier()
        if err is not None:
            raise err
