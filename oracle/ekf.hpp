// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.hpp header).
// Parity status: PINNED on oracle/_ref/libref_ekf.so — the reference's own ekf_algorithm.cpp (+ ekf_algorithm.hpp,
// localization_functions.hpp, localization_struct.hpp) compiled unmodified against stand-in Eigen / ROS headers
// (tests/test_reference_build_ekf.py compares every member after every call); Eigen's arithmetic kernels stay restated.
// CPU restatement of the 27-state EKF of ekf_localization (README says "24-DOF"; STATE_ORDER is 27):
//   /root/reference/src/app/localization/ekf_localization/include/ekf_algorithm.hpp   (ekf_alg.hpp)
//   /root/reference/src/app/localization/ekf_localization/src/ekf_algorithm.cpp       (ekf_alg.cpp)
//   /root/reference/src/app/localization/localization_interface/localization_functions.hpp (lfun.hpp)
// In scope: Init, RunPredictionImu, RunGnssUpdate for the PCM / PCM_INIT sources,
// UpdateEkfState<M>, ComplementaryKalmanFilter, the Check* flags, GetCurrentState.
// Out of scope: RunPrediction (non-IMU), RunCanUpdate, ZUPT, NavSat/BESTPOS branches, CalibrateVehicleToImu
// (all off in config/localization.ini:23-31).
#pragma once
#include <cstdint>

namespace orc {

constexpr int kEkfN = 27;  // STATE_ORDER, ekf_alg.hpp:41-69

// Plain mirror of EkfAlgorithm's members (ekf_alg.hpp:269-289) + the function-static memory of
// ComplementaryKalmanFilter (ekf_alg.cpp:613-614).  Same layout as elm_ekf_state in include/elimaloc_b200.h.
struct EkfStateBlob {
    double pos[3], rot[4] /* w x y z */, vel[3], gyro[3], acc[3], bg[3], ba[3], grav[3], imu_rot[4];
    double P[kEkfN * kEkfN];  // row-major
    double prev_timestamp, prev_gnss_timestamp;
    double ckf_prev_vel_local_x, ckf_prev_time;
    double ego[26], ego_prev_timestamp;  // prev_ego_state_ cache of GetCurrentState (ekf_alg.cpp:786-789,830)
    int32_t reset_for_init_prediction, state_initialized, yaw_initialized, rotation_stabilized, state_stabilized;
    int32_t pcm_init_on_going, pcm_update_count, ckf_has_prev, predictions, updates, reserved[2];
};

// EkfLocalizationConfig subset the in-scope functions read (ekf_localization_config.hpp:20-95)
struct EkfConfig {
    double imu_gravity, ekf_init_x_m, ekf_init_y_m, ekf_init_z_m, ekf_init_roll_deg, ekf_init_pitch_deg, ekf_init_yaw_deg;
    double state_std_pos_m, state_std_rot_deg, state_std_vel_mps, imu_std_gyro_dps, imu_std_acc_mps;
    double imu_bias_cov_gyro, imu_bias_cov_acc;
    int32_t imu_estimate_gravity, use_complementary_filter, reserved[2];
};

// EkfGnssMeasurement (localization_struct.hpp:146-153); source: 3 = PCM, 4 = PCM_INIT (GnssSource enum order)
struct EkfMeasurement {
    double timestamp, pos[3], rot[4] /* w x y z */, pos_cov[9], rot_cov[9];
    int32_t source, reserved;
};

void EkfInit(const EkfConfig& c, EkfStateBlob& s);                                                   // ekf_alg.cpp:22-66
bool EkfPredictImu(const EkfConfig& c, EkfStateBlob& s, double t, const double gyro[3], const double acc[3]);  // :167-316
bool EkfUpdatePose(const EkfConfig& c, EkfStateBlob& s, const EkfMeasurement& m);                    // :318-432
void EkfGetCurrentState(EkfStateBlob& s, double ego_out[26]);                                        // :778-833

}  // namespace orc
