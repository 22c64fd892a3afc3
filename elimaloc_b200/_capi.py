"""ctypes binding of libelimaloc_b200.so (include/elimaloc_b200.h).

The library is the product; this file only declares its C ABI.  There is no fallback: if the shared object is
missing the import fails loudly (build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C elimaloc_b200/csrc`)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ELIMALOC_B200_LIB: developer override to A/B-test another build of the same library (profiles/quick_bench.sh)
LIB_PATH = os.environ.get("ELIMALOC_B200_LIB") or os.path.join(_HERE, "libelimaloc_b200.so")

ELM_OK, ELM_ERR_INVALID, ELM_ERR_CUDA, ELM_ERR_NCCL, ELM_ERR_UNSUPPORTED, ELM_ERR_RANGE, ELM_ERR_STATE, ELM_ERR_IO = range(8)
P2P, GICP, VGICP, AVGICP = 0, 1, 2, 3


class RegConfig(C.Structure):
    """elm_reg_config — RegistrationConfig of the reference (registration.hpp:62-85), solver-read fields only."""
    _fields_ = [("icp_method", C.c_int32), ("max_iteration", C.c_int32), ("max_thread", C.c_int32),
                ("use_radar_cov", C.c_int32), ("debug_print", C.c_int32), ("reserved0", C.c_int32),
                ("max_search_dist", C.c_double), ("lm_lambda", C.c_double),
                ("icp_termination_threshold_m", C.c_double), ("min_overlap_ratio", C.c_double),
                ("max_fitness_score", C.c_double), ("range_variance_m", C.c_double),
                ("azimuth_variance_deg", C.c_double), ("elevation_variance_deg", C.c_double)]


class DeskewTables(C.Structure):
    """elm_deskew_tables — the member tables ImuDeskewInfo / OdomDeskewInfo fill (pcm_matching.cpp:533-729)."""
    _fields_ = [("imu_time", C.POINTER(C.c_double)), ("imu_rot_x", C.POINTER(C.c_double)), ("imu_rot_y", C.POINTER(C.c_double)),
                ("imu_rot_z", C.POINTER(C.c_double)), ("imu_pointer_cur", C.c_int32), ("imu_available", C.c_int32),
                ("odom_available", C.c_int32), ("reserved0", C.c_int32), ("odom_incre_x", C.c_float), ("odom_incre_y", C.c_float),
                ("odom_incre_z", C.c_float), ("reserved1", C.c_float), ("time_scan_cur", C.c_double), ("time_scan_end", C.c_double)]


class EkfConfig(C.Structure):
    """elm_ekf_config (EkfLocalizationConfig subset); defaults of config/localization.ini in ekf.make_ekf_config"""
    _fields_ = [(n, C.c_double) for n in ("imu_gravity", "ekf_init_x_m", "ekf_init_y_m", "ekf_init_z_m", "ekf_init_roll_deg",
                                          "ekf_init_pitch_deg", "ekf_init_yaw_deg", "state_std_pos_m", "state_std_rot_deg",
                                          "state_std_vel_mps", "imu_std_gyro_dps", "imu_std_acc_mps", "imu_bias_cov_gyro",
                                          "imu_bias_cov_acc")] + [("imu_estimate_gravity", C.c_int32),
                                                                  ("use_complementary_filter", C.c_int32), ("reserved", C.c_int32 * 2)]


class EkfState(C.Structure):
    """elm_ekf_state"""
    _fields_ = [("pos", C.c_double * 3), ("rot", C.c_double * 4), ("vel", C.c_double * 3), ("gyro", C.c_double * 3),
                ("acc", C.c_double * 3), ("bg", C.c_double * 3), ("ba", C.c_double * 3), ("grav", C.c_double * 3),
                ("imu_rot", C.c_double * 4), ("P", C.c_double * (27 * 27)), ("prev_timestamp", C.c_double),
                ("prev_gnss_timestamp", C.c_double), ("ckf_prev_vel_local_x", C.c_double), ("ckf_prev_time", C.c_double),
                ("ego", C.c_double * 26), ("ego_prev_timestamp", C.c_double)] + \
               [(n, C.c_int32) for n in ("reset_for_init_prediction", "state_initialized", "yaw_initialized", "rotation_stabilized",
                                         "state_stabilized", "pcm_init_on_going", "pcm_update_count", "ckf_has_prev", "predictions",
                                         "updates")] + [("reserved", C.c_int32 * 2)]


class EkfMeasurement(C.Structure):
    """elm_ekf_measurement; source 3 = PCM, 4 = PCM_INIT"""
    _fields_ = [("timestamp", C.c_double), ("pos", C.c_double * 3), ("rot", C.c_double * 4), ("pos_cov", C.c_double * 9),
                ("rot_cov", C.c_double * 9), ("source", C.c_int32), ("reserved", C.c_int32)]


class ImuQueue(C.Structure):
    """elm_imu_queue"""
    _fields_ = [("stamp", C.POINTER(C.c_double)), ("gyro", C.POINTER(C.c_double)), ("n", C.c_size_t)]


class OdomQueue(C.Structure):
    """elm_odom_queue"""
    _fields_ = [("stamp", C.POINTER(C.c_double)), ("pos", C.POINTER(C.c_double)), ("quat_xyzw", C.POINTER(C.c_double)),
                ("lin_vel", C.POINTER(C.c_double)), ("ang_vel", C.POINTER(C.c_double)), ("n", C.c_size_t)]


class ScanPipelineConfig(C.Structure):
    """elm_scan_pipeline_config"""
    _fields_ = [("input_max_dist", C.c_double), ("input_voxel_ds_m", C.c_double), ("run_deskew", C.c_int32),
                ("lidar_scan_time_end", C.c_int32), ("tf_ego_to_lidar", C.c_double * 16)]


class ScanResult(C.Structure):
    """elm_scan_result"""
    _fields_ = [("T_lidar", C.c_double * 16), ("T_ego", C.c_double * 16), ("fitness_score", C.c_double), ("local_cov", C.c_double * 36),
                ("pose_cov", C.c_double * 36), ("time_scan_cur", C.c_double), ("time_scan_end", C.c_double)] + \
               [(n, C.c_int32) for n in ("is_success", "iterations", "n_raw", "n_after_filter", "n_registered", "deskew_ok")]


ELM_IMU_QUEUE_LENGTH = 2000


class ElmError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"elimaloc_b200 status {status}: {message}")
        self.status = status


_dp, _fp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); every symbol include/elimaloc_b200.h declares
SIGNATURES = {
    "elm_last_error": (C.c_char_p, []),
    "elm_device_count": (C.c_int, []),
    "elm_map_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_double, C.c_int, C.c_int]),
    "elm_map_destroy": (None, [C.c_void_p]),
    "elm_map_add_points": (C.c_int, [C.c_void_p, _fp, C.c_size_t]),
    "elm_map_cal_voxel_cov": (C.c_int, [C.c_void_p]),
    "elm_map_cal_point_cov": (C.c_int, [C.c_void_p, C.c_double]),
    "elm_map_set_gpu_build": (C.c_int, [C.c_void_p, C.c_int]),
    "elm_map_build_times": (C.c_int, [C.c_void_p, _dp]),
    "elm_map_empty": (C.c_int, [C.c_void_p]),
    "elm_map_num_voxels": (C.c_size_t, [C.c_void_p]),
    "elm_map_num_points": (C.c_size_t, [C.c_void_p]),
    "elm_map_export": (C.c_int, [C.c_void_p, _ip, _ip, _dp, _dp, _fp, _dp, _dp]),
    "elm_shape_pcm_covariance": (C.c_int, [_dp, _dp, C.c_double, _dp]),
    "elm_map_find_ground_height": (C.c_int, [C.c_void_p, C.c_double, C.c_double, _dp, _ip]),
    "elm_pcd_read_xyz": (C.c_int, [C.c_char_p, _fp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "elm_map_add_points_pcd": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t)]),
    "elm_map_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "elm_map_load": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p, C.c_int]),
    "elm_map_directory_check": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "elm_registration_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "elm_registration_destroy": (None, [C.c_void_p]),
    "elm_run_register": (C.c_int, [C.c_void_p, C.c_void_p, _fp, C.c_size_t, _dp, C.POINTER(RegConfig), _dp, _ip, _dp, _dp]),
    "elm_register_enqueue": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, _dp, C.POINTER(RegConfig)]),
    "elm_register_fetch": (C.c_int, [C.c_void_p, _dp, _ip, _dp, _dp, _ip]),
    "elm_linearize": (C.c_int, [C.c_void_p, C.c_void_p, _fp, C.c_size_t, _dp, C.POINTER(RegConfig), _dp, _dp, _dp,
                                C.POINTER(C.c_int64)]),
    "elm_correspondences": (C.c_int, [C.c_void_p, C.c_void_p, _fp, C.c_size_t, _dp, C.c_int, C.c_double, _ip, _dp]),
    "elm_correspondences_sequence": (C.c_int, [C.c_void_p, C.c_void_p, _fp, C.c_size_t, _dp, C.c_int, C.c_int, C.c_double, _ip, _dp]),
    "elm_registration_set_warm_start": (C.c_int, [C.c_void_p, C.c_int]),
    "elm_registration_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "elm_registration_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "elm_registration_profile": (C.c_int, [C.c_void_p, _dp, _dp, C.POINTER(C.c_int64)]),
    "elm_registration_profile_by_kind": (C.c_int, [C.c_void_p, _dp, C.POINTER(C.c_int64)]),
    "elm_registration_set_stats": (C.c_int, [C.c_void_p, C.c_int]),
    "elm_registration_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "elm_registration_stats_raw": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "elm_registration_set_fused": (C.c_int, [C.c_void_p, C.c_int]),
    "elm_registration_set_binning": (C.c_int, [C.c_void_p, C.c_int]),
    "elm_registration_set_exhaustive": (C.c_int, [C.c_void_p, C.c_int]),
    "elm_deskew_points": (C.c_int, [C.c_void_p, _fp, _fp, C.c_size_t, C.POINTER(DeskewTables), _fp]),
    "elm_deskew_points_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(DeskewTables), C.c_void_p]),
    "elm_scan_preprocess": (C.c_int, [C.c_void_p, _fp, _fp, C.c_size_t, C.c_double, C.c_double, _fp, _fp, _ip, C.POINTER(C.c_size_t)]),
    "elm_scan_preprocess_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]),
    "elm_ekf_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(EkfConfig), C.c_int, C.c_void_p]),
    "elm_ekf_destroy": (None, [C.c_void_p]),
    "elm_ekf_predict_imu": (C.c_int, [C.c_void_p, C.c_double, _dp, _dp]),
    "elm_ekf_update_pose": (C.c_int, [C.c_void_p, C.POINTER(EkfMeasurement)]),
    "elm_ekf_get_state": (C.c_int, [C.c_void_p, C.POINTER(EkfState)]),
    "elm_ekf_set_state": (C.c_int, [C.c_void_p, C.POINTER(EkfState)]),
    "elm_ekf_get_current_state": (C.c_int, [C.c_void_p, _dp]),
    "elm_deskew_build_tables": (C.c_int, [C.POINTER(ImuQueue), C.POINTER(OdomQueue), C.POINTER(DeskewTables), _dp, C.POINTER(C.c_size_t),
                                          C.POINTER(C.c_size_t)]),
    "elm_scan_pipeline_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(ScanPipelineConfig)]),
    "elm_scan_pipeline_destroy": (None, [C.c_void_p]),
    "elm_scan_pipeline_deskew": (C.c_int, [C.c_void_p, _fp, _fp, C.c_size_t, C.c_double, C.POINTER(ImuQueue), C.POINTER(OdomQueue), _dp, _dp, _ip]),
    "elm_scan_pipeline_register": (C.c_int, [C.c_void_p, C.c_void_p, _dp, C.POINTER(RegConfig)]),
    "elm_scan_pipeline_ekf_update": (C.c_int, [C.c_void_p, C.c_void_p]),
    "elm_scan_pipeline_fetch": (C.c_int, [C.c_void_p, C.POINTER(ScanResult)]),
    "elm_ekf_enable_state_ring": (C.c_int, [C.c_void_p, C.c_int]),
    "elm_comm_unique_id": (C.c_int, [_u8p]),
    "elm_registration_set_comm": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int]),
    "elm_registration_peer_export": (C.c_int, [C.c_void_p, _u8p]),
    "elm_registration_peer_attach": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int]),
    "elm_registration_peer_detach": (C.c_int, [C.c_void_p]),
}

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                              "(run __graft_entry__.build() or `make -C elimaloc_b200/csrc`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(status):
    if status != ELM_OK:
        raise ElmError(status, lib().elm_last_error().decode("utf-8", "replace"))
