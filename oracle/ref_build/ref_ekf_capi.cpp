// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// C entry points over the REFERENCE's own EkfAlgorithm (ekf_localization/src/ekf_algorithm.cpp + include/ekf_algorithm.hpp +
// localization_interface/localization_functions.hpp / localization_struct.hpp, all compiled unmodified from /root/reference
// against the stand-in headers of stubs/ and node_stubs/), with the state / config / measurement layouts of oracle/ekf.hpp so
// that tests/test_reference_build_ekf.py can compare the oracle's EKF member for member.
// Built by oracle/Makefile into oracle/_ref/libref_ekf.so (git-ignored).
#include <deque>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "localization_functions.hpp"  // the reference's (include guard: a second inclusion below is a no-op)
#include "ekf_localization_config.hpp"

// The filter keeps its state in private members; the dump below only READS them.  Every standard / stand-in header the class
// header pulls in is already included above, so the macro touches nothing but the reference's own class definition.
#define private public
#include "ekf_algorithm.hpp"
#undef private

#include "../ekf.hpp"  // orc::EkfStateBlob, orc::EkfConfig, orc::EkfMeasurement (plain layouts)

namespace {

struct Quiet {  // PrintState / Init / PCM-init messages go to std::cout; keep test logs clean
    Quiet() { std::cout.setstate(std::ios_base::failbit); }
    ~Quiet() { std::cout.clear(); }
};

EkfLocalizationConfig to_cfg(const orc::EkfConfig& c) {
    EkfLocalizationConfig r{};
    r.b_debug_print = false;
    r.b_debug_imu_print = false;
    r.i_gps_type = GpsType::ODOMETRY;  // the NavSatFix / BESTPOS branches are out of scope
    r.b_use_gps = false;
    r.b_use_can = false;
    r.b_use_imu = true;
    r.b_use_pcm_matching = true;
    r.b_use_zupt = false;
    r.b_imu_estimate_calibration = false;
    r.b_use_complementary_filter = c.use_complementary_filter != 0;
    r.b_imu_estimate_gravity = c.imu_estimate_gravity != 0;
    r.d_imu_gravity = c.imu_gravity;
    r.d_ekf_init_x_m = c.ekf_init_x_m;
    r.d_ekf_init_y_m = c.ekf_init_y_m;
    r.d_ekf_init_z_m = c.ekf_init_z_m;
    r.d_ekf_init_roll_deg = c.ekf_init_roll_deg;
    r.d_ekf_init_pitch_deg = c.ekf_init_pitch_deg;
    r.d_ekf_init_yaw_deg = c.ekf_init_yaw_deg;
    r.d_state_std_pos_m = c.state_std_pos_m;
    r.d_state_std_rot_deg = c.state_std_rot_deg;
    r.d_state_std_vel_mps = c.state_std_vel_mps;
    r.d_imu_std_gyro_dps = c.imu_std_gyro_dps;
    r.d_imu_std_acc_mps = c.imu_std_acc_mps;
    r.d_ekf_imu_bias_cov_gyro = c.imu_bias_cov_gyro;
    r.d_ekf_imu_bias_cov_acc = c.imu_bias_cov_acc;
    r.d_can_vel_scale_factor = 1.0;
    return r;
}

void put3(double* o, const Eigen::Vector3d& v) { o[0] = v.x(); o[1] = v.y(); o[2] = v.z(); }
void putq(double* o, const Eigen::Quaterniond& q) { o[0] = q.w(); o[1] = q.x(); o[2] = q.y(); o[3] = q.z(); }

}  // namespace

extern "C" {

void* ref_ekf_create(const orc::EkfConfig* c) {
    Quiet q;
    auto* f = new EkfAlgorithm(to_cfg(*c));
    f->Init();
    return f;
}
void ref_ekf_destroy(void* f) { delete static_cast<EkfAlgorithm*>(f); }

int ref_ekf_predict_imu(void* f, double t, const double* gyro, const double* acc) {
    Quiet q;
    ImuStruct imu;
    imu.timestamp = t;
    imu.gyro = Eigen::Vector3d(gyro[0], gyro[1], gyro[2]);
    imu.acc = Eigen::Vector3d(acc[0], acc[1], acc[2]);
    return static_cast<EkfAlgorithm*>(f)->RunPredictionImu(t, imu) ? 1 : 0;
}

int ref_ekf_update_pose(void* f, const orc::EkfMeasurement* m) {
    Quiet q;
    EkfGnssMeasurement g;
    g.timestamp = m->timestamp;
    g.gnss_source = static_cast<GnssSource>(m->source);  // enum order: NOVATEL, NAVSATFIX, BESTPOS, PCM (3), PCM_INIT (4)
    g.pos = Eigen::Vector3d(m->pos[0], m->pos[1], m->pos[2]);
    g.rot = Eigen::Quaterniond(m->rot[0], m->rot[1], m->rot[2], m->rot[3]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { g.pos_cov(i, j) = m->pos_cov[3 * i + j]; g.rot_cov(i, j) = m->rot_cov[3 * i + j]; }
    return static_cast<EkfAlgorithm*>(f)->RunGnssUpdate(g) ? 1 : 0;
}

// EgoState in the oracle's 26-value order (oracle/ekf.cpp EkfGetCurrentState)
void ref_ekf_get_current_state(void* f, double* o) {
    const EgoState e = static_cast<EkfAlgorithm*>(f)->GetCurrentState();
    const double v[26] = {e.timestamp, e.x_m, e.y_m, e.z_m, e.roll_rad, e.pitch_rad, e.yaw_rad, e.roll_vel, e.pitch_vel, e.yaw_vel,
                          e.vx, e.vy, e.vz, e.ax, e.ay, e.az, e.x_cov_m, e.y_cov_m, e.z_cov_m, e.latitude_std, e.longitude_std,
                          e.height_std, e.roll_cov_rad, e.pitch_cov_rad, e.yaw_cov_rad, 0.0};
    for (int i = 0; i < 26; ++i) o[i] = v[i];
}

// Members -> the oracle's blob.  Not reachable: the function-static memory of ComplementaryKalmanFilter (left 0).
void ref_ekf_dump(void* fp, orc::EkfStateBlob* s) {
    auto* f = static_cast<EkfAlgorithm*>(fp);
    std::memset(s, 0, sizeof *s);
    put3(s->pos, f->S_.pos);
    putq(s->rot, f->S_.rot);
    put3(s->vel, f->S_.vel);
    put3(s->gyro, f->S_.gyro);
    put3(s->acc, f->S_.acc);
    put3(s->bg, f->S_.bg);
    put3(s->ba, f->S_.ba);
    put3(s->grav, f->S_.grav);
    putq(s->imu_rot, f->S_.imu_rot);
    for (int i = 0; i < STATE_ORDER; ++i) for (int j = 0; j < STATE_ORDER; ++j) s->P[i * STATE_ORDER + j] = f->P_(i, j);
    s->prev_timestamp = f->prev_timestamp_;
    s->prev_gnss_timestamp = f->prev_gnss_.timestamp;
    s->reset_for_init_prediction = f->b_reset_for_init_prediction_;
    s->state_initialized = f->b_state_initialized_;
    s->yaw_initialized = f->b_yaw_initialized_;
    s->rotation_stabilized = f->b_rotation_stabilized_;
    s->state_stabilized = f->b_state_stabilized_;
    s->pcm_init_on_going = f->b_pcm_init_on_going_;
    s->pcm_update_count = f->i_pcm_update_count_;
}

}  // extern "C"
