"""Host-side mirror of the reference's EkfAlgorithm (ekf_localization/include/ekf_algorithm.hpp:81-104) over the C ABI."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import EkfConfig, EkfMeasurement, EkfState, check, lib

PCM, PCM_INIT = 3, 4  # GnssSource (localization_struct.hpp:28)


def make_ekf_config(**kw):
    """defaults = config/localization.ini:15-66 of the reference"""
    d = dict(imu_gravity=9.81, ekf_init_x_m=0.0, ekf_init_y_m=0.0, ekf_init_z_m=0.0, ekf_init_roll_deg=0.0, ekf_init_pitch_deg=0.0,
             ekf_init_yaw_deg=0.0, state_std_pos_m=0.02, state_std_rot_deg=0.2, state_std_vel_mps=2.0, imu_std_gyro_dps=0.01,
             imu_std_acc_mps=0.001, imu_bias_cov_gyro=0.0001, imu_bias_cov_acc=0.0001, imu_estimate_gravity=1,
             use_complementary_filter=1)
    d.update(kw)
    return EkfConfig(**d)


def make_measurement(timestamp, pos, rot_wxyz, pos_cov, rot_cov, source=PCM):
    m = EkfMeasurement()
    m.timestamp = float(timestamp)
    m.pos[:] = [float(v) for v in pos]
    m.rot[:] = [float(v) for v in rot_wxyz]
    m.pos_cov[:] = [float(v) for v in np.asarray(pos_cov, dtype=np.float64).reshape(9)]
    m.rot_cov[:] = [float(v) for v in np.asarray(rot_cov, dtype=np.float64).reshape(9)]
    m.source = int(source)
    return m


def state_to_dict(s):
    out = {}
    for name, _ in EkfState._fields_:
        v = getattr(s, name)
        out[name] = np.array(v[:]) if hasattr(v, "__len__") else v
    out["P"] = out["P"].reshape(27, 27)
    return out


class EkfAlgorithm:
    def __init__(self, cfg, device=0, stream=None):
        self._h = C.c_void_p()
        self.cfg_ = cfg
        check(lib().elm_ekf_create(C.byref(self._h), C.byref(cfg), int(device), C.c_void_p(stream) if stream else None))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().elm_ekf_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown: the module globals may already be gone)
                pass

    def RunPredictionImu(self, cur_timestamp, gyro, acc):
        g = np.ascontiguousarray(gyro, dtype=np.float64)
        a = np.ascontiguousarray(acc, dtype=np.float64)
        check(lib().elm_ekf_predict_imu(self._h, float(cur_timestamp), g.ctypes.data_as(_capi._dp), a.ctypes.data_as(_capi._dp)))

    def RunGnssUpdate(self, meas):
        check(lib().elm_ekf_update_pose(self._h, C.byref(meas)))

    def state(self):
        s = EkfState()
        check(lib().elm_ekf_get_state(self._h, C.byref(s)))
        return s

    def set_state(self, s):
        check(lib().elm_ekf_set_state(self._h, C.byref(s)))

    def enable_state_ring(self, enable=True):
        """PublishInThread's deque of EgoStates kept in HBM (needed by ScanPipeline.ekf_update)."""
        check(lib().elm_ekf_enable_state_ring(self._h, int(bool(enable))))

    def GetCurrentState(self):
        ego = np.zeros(26)
        check(lib().elm_ekf_get_current_state(self._h, ego.ctypes.data_as(_capi._dp)))
        return ego
