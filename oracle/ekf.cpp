// ORACLE — TEST INFRASTRUCTURE ONLY (see ekf.hpp).
#include "ekf.hpp"

#include <cmath>
#include <cstring>

#include "smallmat.hpp"

namespace orc {

namespace {
constexpr int N = kEkfN;
enum { S_X = 0, S_Y, S_Z, S_ROLL, S_PITCH, S_YAW, S_VX, S_VY, S_VZ, S_ROLL_RATE, S_PITCH_RATE, S_YAW_RATE, S_AX, S_AY, S_AZ,
       S_B_ROLL_RATE, S_B_PITCH_RATE, S_B_YAW_RATE, S_B_AX, S_B_AY, S_B_AZ, S_G_X, S_G_Y, S_G_Z, S_IMU_ROLL, S_IMU_PITCH, S_IMU_YAW };
constexpr double INIT_STATE_COV = 100.0;  // ekf_alg.hpp:73
constexpr double kPi = 3.14159265358979323846;

struct Quat { double w, x, y, z; };
Quat qmul(const Quat& a, const Quat& b) {  // Eigen quaternion product
    return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
Quat qnormalized(const Quat& q) {
    const double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return Quat{q.w / n, q.x / n, q.y / n, q.z / n};
}
M3 qtoR(const Quat& q) {  // Quaterniond::toRotationMatrix()
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    M3 R;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1 - (txx + tyy);
    return R;
}
Quat qfromR(const M3& m) {  // Quaterniond(Matrix3d): Shepperd
    Quat q;
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t; t = 0.5 / t;
        q.x = (m(2, 1) - m(1, 2)) * t; q.y = (m(0, 2) - m(2, 0)) * t; q.z = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        double v[3];
        v[i] = 0.5 * t; t = 0.5 / t;
        q.w = (m(k, j) - m(j, k)) * t;
        v[j] = (m(j, i) + m(i, j)) * t;
        v[k] = (m(k, i) + m(i, k)) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}
Quat qfromAngleAxis(double angle, const V3& axis) {  // Quaterniond(AngleAxisd)
    const double h = 0.5 * angle, s = std::sin(h);
    return Quat{std::cos(h), s * axis.x, s * axis.y, s * axis.z};
}
V3 qrotate(const Quat& q, const V3& v) {  // Quaterniond * Vector3d (_transformVector)
    const V3 u(q.x, q.y, q.z);
    V3 uv(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
    uv = uv + uv;
    const V3 c(u.y * uv.z - u.z * uv.y, u.z * uv.x - u.x * uv.z, u.x * uv.y - u.y * uv.x);
    return v + q.w * uv + c;
}
Quat qinverse(const Quat& q) {
    const double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    return Quat{q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
}
double NormAngleRad(double a) {  // lfun.hpp NormAngleRad
    while (a > kPi) a -= kPi * 2.;
    while (a < -kPi) a += kPi * 2.;
    return a;
}
V3 RotToVec(const M3& R) {  // lfun.hpp RotToVec
    double a0, a1, a2;
    if (std::fabs(R(2, 0)) > 0.998) {
        a2 = std::atan2(-R(1, 2), R(1, 1));
        a1 = kPi / 2 * (R(2, 0) >= 0 ? 1 : -1);
        a0 = 0;
    } else {
        a1 = std::asin(-R(2, 0));
        a0 = std::atan2(R(2, 1) / std::cos(a1), R(2, 2) / std::cos(a1));
        a2 = std::atan2(R(1, 0) / std::cos(a1), R(0, 0) / std::cos(a1));
    }
    a0 = std::fmod(a0 + kPi, 2 * kPi) - kPi;
    a1 = std::fmod(a1 + kPi, 2 * kPi) - kPi;
    a2 = std::fmod(a2 + kPi, 2 * kPi) - kPi;
    return V3(a0, a1, a2);
}
M3 skew(const V3& v) { M3 m; m(0, 1) = -v.z; m(0, 2) = v.y; m(1, 0) = v.z; m(1, 2) = -v.x; m(2, 0) = -v.y; m(2, 1) = v.x; return m; }
M3 ExpSO3(const V3& omega) {  // lfun.hpp Exp
    const double theta = norm(omega);
    if (theta < 1e-5) return M3::Identity();
    const V3 axis(omega.x / theta, omega.y / theta, omega.z / theta);
    const M3 K = skew(axis), KK = mul(K, K);
    M3 R = M3::Identity();
    const double s = std::sin(theta), c1 = 1 - std::cos(theta);
    for (int i = 0; i < 9; ++i) R.m[i] += s * K.m[i] + c1 * KK.m[i];
    return R;
}
Quat ExpGyroToQuat(const V3& gyro, double dt) { return qfromR(ExpSO3(dt * gyro)); }  // lfun.hpp ExpGyroToQuat
M3 PartialDerivativeRotWrtGyro(const V3& gyro, double dt) {                            // lfun.hpp
    const V3 omega = dt * gyro;
    const double theta = norm(omega);
    M3 Z;
    if (theta < 1e-5) return Z;
    const V3 axis(omega.x / theta, omega.y / theta, omega.z / theta);
    const M3 K = skew(axis), KK = mul(K, K);
    M3 R = M3::Identity();
    const double a = (1 - std::cos(theta)) / (theta * theta), b = (theta - std::sin(theta)) / (theta * theta * theta);
    for (int i = 0; i < 9; ++i) R.m[i] = dt * (R.m[i] + a * K.m[i] + b * KK.m[i]);
    return R;
}
Quat getq(const double* r) { return Quat{r[0], r[1], r[2], r[3]}; }
void setq(double* r, const Quat& q) { r[0] = q.w; r[1] = q.x; r[2] = q.y; r[3] = q.z; }
double& P_(EkfStateBlob& s, int i, int j) { return s.P[i * N + j]; }

void CheckYawInitialized(EkfStateBlob& s) { s.yaw_initialized = std::sqrt(P_(s, S_YAW, S_YAW)) < 5.0 * kPi / 180.0; }  // ekf_alg.hpp:164-177
void CheckStateInitialized(EkfStateBlob& s) {                                                                             // :148-162
    s.state_initialized = std::sqrt(P_(s, S_ROLL, S_ROLL)) < 5.0 * kPi / 180.0 && std::sqrt(P_(s, S_PITCH, S_PITCH)) < 5.0 * kPi / 180.0 &&
                          std::sqrt(P_(s, S_YAW, S_YAW)) < 5.0 * kPi / 180.0 && std::sqrt(P_(s, S_X, S_X)) < 1.0 && std::sqrt(P_(s, S_Y, S_Y)) < 1.0;
}
void CheckRotationStabilized(EkfStateBlob& s) {  // :179-193
    s.rotation_stabilized = std::sqrt(P_(s, S_ROLL, S_ROLL)) < 0.2 * kPi / 180.0 && std::sqrt(P_(s, S_PITCH, S_PITCH)) < 0.2 * kPi / 180.0 &&
                            std::sqrt(P_(s, S_YAW, S_YAW)) < 0.2 * kPi / 180.0;
}
void CheckStateStabilized(EkfStateBlob& s) {  // :195-209
    s.state_stabilized = std::sqrt(P_(s, S_ROLL, S_ROLL)) < 0.2 * kPi / 180.0 && std::sqrt(P_(s, S_PITCH, S_PITCH)) < 0.2 * kPi / 180.0 &&
                         std::sqrt(P_(s, S_YAW, S_YAW)) < 0.2 * kPi / 180.0 && std::sqrt(P_(s, S_X, S_X)) < 0.5 && std::sqrt(P_(s, S_Y, S_Y)) < 0.5;
}

// UpdateEkfState<M, M> (ekf_alg.hpp:116-145) with H = rows `hrow[0..M)` of the identity (all in-scope H are selectors)
template <int M>
void UpdateEkfState(EkfStateBlob& s, const double* K /* N x M */, const double* Y, const int* hrow) {
    double du[N];
    for (int i = 0; i < N; ++i) { double a = 0; for (int k = 0; k < M; ++k) a += K[i * M + k] * Y[k]; du[i] = a; }
    for (int k = 0; k < 3; ++k) {
        s.pos[k] += du[S_X + k]; s.vel[k] += du[S_VX + k]; s.gyro[k] += du[S_ROLL_RATE + k]; s.acc[k] += du[S_AX + k];
        s.bg[k] += du[S_B_ROLL_RATE + k]; s.ba[k] += du[S_B_AX + k]; s.grav[k] += du[S_G_X + k];
    }
    const V3 rd(du[3], du[4], du[5]);
    setq(s.rot, qnormalized(qmul(getq(s.rot), qfromAngleAxis(norm(rd), normalized(rd)))));
    const V3 id(du[24], du[25], du[26]);
    setq(s.imu_rot, qnormalized(qmul(getq(s.imu_rot), qfromAngleAxis(norm(id), normalized(id)))));
    // P = P - K * H * P
    double HP[M][N];
    for (int k = 0; k < M; ++k) for (int j = 0; j < N; ++j) HP[k][j] = s.P[hrow[k] * N + j];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) { double a = 0; for (int k = 0; k < M; ++k) a += K[i * M + k] * HP[k][j]; s.P[i * N + j] -= a; }
}

// ComplementaryKalmanFilter (ekf_alg.cpp:597-701)
void ComplementaryKalmanFilter(const EkfConfig&, EkfStateBlob& s, double timestamp, const double acc_in[3]) {
    const V3 acc_meas(acc_in[0] - s.ba[0], acc_in[1] - s.ba[1], acc_in[2] - s.ba[2]);
    const Quat rot = getq(s.rot);
    const V3 vel_local = qrotate(qinverse(rot), V3(s.vel[0], s.vel[1], s.vel[2]));
    const double centripetal_acc = vel_local.x * s.gyro[2];
    if (!s.ckf_has_prev) { s.ckf_prev_vel_local_x = vel_local.x; s.ckf_prev_time = timestamp; s.ckf_has_prev = 1; }  // function statics :613-614
    const double dt = timestamp - s.ckf_prev_time;
    if (dt < 1e-6) return;
    const double est_acc_x = (vel_local.x - s.ckf_prev_vel_local_x) / dt;
    s.ckf_prev_vel_local_x = vel_local.x;
    s.ckf_prev_time = timestamp;
    V3 comp(acc_meas.x, acc_meas.y - centripetal_acc, acc_meas.z);
    if (s.rotation_stabilized) comp.x -= est_acc_x;
    const double d_acc_diff = norm(acc_meas) - norm(V3(s.grav[0], s.grav[1], s.grav[2]));
    const V3 g = normalized(comp);
    double z[2] = {std::atan2(g.y, g.z), -std::asin(g.x)};
    const V3 rpy = RotToVec(qtoR(rot));
    double innov[2] = {NormAngleRad(z[0] - rpy.x), NormAngleRad(z[1] - rpy.y)};
    double base = 1.0 * kPi / 180.0;
    if (!s.state_initialized) base = 10.0 * kPi / 180.0;
    const double cu = std::fabs(centripetal_acc) / 9.81 * 10.0, lu = std::fabs(est_acc_x) / 9.81 * 10.0, au = std::fabs(d_acc_diff) / 9.81 * 10.0;
    const double lat = 1.0 + au + cu, lon = 1.0 + au + lu;
    const double R0 = std::fmax(std::pow(base * lat, 2), std::pow(1.0 * kPi / 180.0, 2));
    const double R1 = std::fmax(std::pow(base * lon, 2), std::pow(1.0 * kPi / 180.0, 2));
    // S = H P H^T + R (2x2), K = P H^T S^-1
    const double S00 = P_(s, S_ROLL, S_ROLL) + R0, S01 = P_(s, S_ROLL, S_PITCH), S10 = P_(s, S_PITCH, S_ROLL), S11 = P_(s, S_PITCH, S_PITCH) + R1;
    const double det = S00 * S11 - S01 * S10;
    const double i00 = S11 / det, i01 = -S01 / det, i10 = -S10 / det, i11 = S00 / det;
    double K[N * 2];
    for (int i = 0; i < N; ++i) {
        const double a = P_(s, i, S_ROLL), b = P_(s, i, S_PITCH);
        K[i * 2] = a * i00 + b * i10;
        K[i * 2 + 1] = a * i01 + b * i11;
    }
    const int hrow[2] = {S_ROLL, S_PITCH};
    UpdateEkfState<2>(s, K, innov, hrow);
}
}  // namespace

void EkfInit(const EkfConfig& c, EkfStateBlob& s) {
    std::memset(&s, 0, sizeof s);
    s.pos[0] = c.ekf_init_x_m; s.pos[1] = c.ekf_init_y_m; s.pos[2] = c.ekf_init_z_m;
    const Quat q = qmul(qmul(qfromAngleAxis(c.ekf_init_yaw_deg * kPi / 180.0, V3(0, 0, 1)), qfromAngleAxis(c.ekf_init_pitch_deg * kPi / 180.0, V3(0, 1, 0))),
                        qfromAngleAxis(c.ekf_init_roll_deg * kPi / 180.0, V3(1, 0, 0)));
    setq(s.rot, q);
    s.imu_rot[0] = 1.0;  // EkfState default: identity
    s.grav[2] = c.imu_gravity;
    for (int i = 0; i < N; ++i) s.P[i * N + i] = INIT_STATE_COV;
    for (int k = 0; k < 3; ++k) {
        s.P[(S_B_ROLL_RATE + k) * N + S_B_ROLL_RATE + k] = c.imu_bias_cov_gyro;
        s.P[(S_B_AX + k) * N + S_B_AX + k] = c.imu_bias_cov_acc;
        s.P[(S_G_X + k) * N + S_G_X + k] = c.imu_bias_cov_acc;
        s.P[(S_IMU_ROLL + k) * N + S_IMU_ROLL + k] = c.imu_bias_cov_gyro;
    }
    s.reset_for_init_prediction = 1;
}

bool EkfPredictImu(const EkfConfig& c, EkfStateBlob& s, double t, const double gyro_in[3], const double acc_in[3]) {
    if (s.reset_for_init_prediction) { s.prev_timestamp = t; s.reset_for_init_prediction = 0; return false; }  // :182-187
    if (s.pcm_init_on_going) { s.prev_timestamp = t; return false; }                                           // :189-194
    CheckRotationStabilized(s);                                                                                // :196
    if (!s.state_initialized) {                                                                                // :198-208
        s.prev_timestamp = t;
        if (s.yaw_initialized && c.use_complementary_filter) ComplementaryKalmanFilter(c, s, t, acc_in);
        return false;
    }
    if (std::fabs(t - s.prev_timestamp) < 1e-6) return false;  // :210-213
    const double dt = t - s.prev_timestamp;
    const Quat rot_prev = getq(s.rot);
    const M3 G_R_I = qtoR(rot_prev);                                                    // :231
    const V3 cg(gyro_in[0] - s.bg[0], gyro_in[1] - s.bg[1], gyro_in[2] - s.bg[2]);      // :234
    setq(s.rot, qnormalized(qmul(rot_prev, ExpGyroToQuat(cg, dt))));                    // :235-236
    const V3 ca(acc_in[0] - s.ba[0], acc_in[1] - s.ba[1], acc_in[2] - s.ba[2]);         // :239
    const V3 ag = mul(G_R_I, ca) - V3(s.grav[0], s.grav[1], s.grav[2]);                 // :240
    const double agv[3] = {ag.x, ag.y, ag.z}, cgv[3] = {cg.x, cg.y, cg.z};
    for (int k = 0; k < 3; ++k) {
        s.pos[k] += s.vel[k] * dt + 0.5 * agv[k] * dt * dt;  // :243 (uses the previous velocity)
    }
    for (int k = 0; k < 3; ++k) { s.vel[k] += agv[k] * dt; s.gyro[k] = cgv[k]; s.acc[k] = agv[k]; }  // :244-248
    // Q (:256-272)
    double Q[N];
    for (int i = 0; i < N; ++i) Q[i] = 0.0;
    const double d2 = dt * dt;
    for (int k = 0; k < 3; ++k) {
        Q[S_X + k] = std::pow(c.state_std_pos_m, 2) * dt * dt;
        Q[S_ROLL + k] = std::pow(c.state_std_rot_deg * kPi / 180.0, 2) * dt * dt;
        Q[S_VX + k] = std::pow(c.state_std_vel_mps, 2) * dt * dt;
        Q[S_ROLL_RATE + k] = std::pow(c.imu_std_gyro_dps * kPi / 180.0, 2) * dt * dt;
        Q[S_AX + k] = std::pow(c.imu_std_acc_mps, 2) * dt * dt;
        Q[S_B_ROLL_RATE + k] = std::pow(c.imu_bias_cov_gyro, 2) * dt * dt;
        Q[S_B_AX + k] = std::pow(c.imu_bias_cov_acc, 2) * dt * dt;
        Q[S_G_X + k] = std::pow(c.imu_bias_cov_acc, 2) * dt * dt;
        Q[S_IMU_ROLL + k] = std::pow(c.state_std_rot_deg * kPi / 180.0, 2) * dt * dt;
    }
    (void)d2;
    // F (:275-297)
    static thread_local double F[N * N], FP[N * N];
    for (int i = 0; i < N * N; ++i) F[i] = 0.0;
    for (int i = 0; i < N; ++i) F[i * N + i] = 1.0;
    const M3 dRdg = PartialDerivativeRotWrtGyro(cg, dt);
    for (int a = 0; a < 3; ++a) {
        F[(S_X + a) * N + S_VX + a] = dt;
        F[(S_ROLL_RATE + a) * N + S_B_ROLL_RATE + a] = -1.0;
        for (int b = 0; b < 3; ++b) {
            F[(S_X + a) * N + S_B_AX + b] = -0.5 * G_R_I(a, b) * dt * dt;
            F[(S_ROLL + a) * N + S_B_ROLL_RATE + b] = -dRdg(a, b);
            F[(S_VX + a) * N + S_B_AX + b] = -G_R_I(a, b) * dt;
            F[(S_AX + a) * N + S_B_AX + b] = -G_R_I(a, b);
        }
    }
    if (c.imu_estimate_gravity) {
        F[S_Z * N + S_G_Z] = -0.5 * dt * dt;
        F[S_VZ * N + S_G_Z] = -dt;
        F[S_AZ * N + S_G_Z] = -1.0;
    }
    // P = F P F^T + Q (:300)
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) { double a = 0; for (int k = 0; k < N; ++k) a += F[i * N + k] * s.P[k * N + j]; FP[i * N + j] = a; }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) { double a = 0; for (int k = 0; k < N; ++k) a += FP[i * N + k] * F[j * N + k]; s.P[i * N + j] = a + (i == j ? Q[i] : 0.0); }
    s.prev_timestamp = t;
    s.predictions += 1;
    if (c.use_complementary_filter) ComplementaryKalmanFilter(c, s, t, acc_in);  // :312
    return true;
}

bool EkfUpdatePose(const EkfConfig& c, EkfStateBlob& s, const EkfMeasurement& m) {
    if (m.source == 4) {  // PCM_INIT (:324-349)
        for (int k = 0; k < 3; ++k) { s.pos[k] = m.pos[k]; s.vel[k] = s.gyro[k] = s.acc[k] = s.bg[k] = s.ba[k] = 0.0; s.grav[k] = 0.0; }
        for (int k = 0; k < 4; ++k) s.rot[k] = m.rot[k];
        s.grav[2] = c.imu_gravity;
        for (int i = 0; i <= S_AZ; ++i) for (int j = 0; j <= S_AZ; ++j) s.P[i * N + j] = (i == j) ? INIT_STATE_COV : 0.0;
        s.state_initialized = 1; s.yaw_initialized = 1; s.pcm_init_on_going = 1;
        return true;
    }
    CheckYawInitialized(s); CheckStateInitialized(s); CheckRotationStabilized(s); CheckStateStabilized(s);  // :351-354
    if (s.pcm_init_on_going && m.source == 3) {                                                             // :357-364
        if (s.pcm_update_count > 10) s.pcm_init_on_going = 0;
        s.pcm_update_count++;
    }
    // S = H P H^T + R with H = [I6 0] (:369-400)
    M6 S;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) S(i, j) = s.P[i * N + j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { S(i, j) += m.pos_cov[i * 3 + j]; S(3 + i, 3 + j) += m.rot_cov[i * 3 + j]; }
    const M6 Si = inverse(S);
    double K[N * 6];                                                                                         // :403
    for (int i = 0; i < N; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += s.P[i * N + k] * Si(k, j); K[i * 6 + j] = a; }
    // Y (:406-410)
    const V3 sa = RotToVec(qtoR(qnormalized(getq(s.rot)))), ma = RotToVec(qtoR(qnormalized(getq(m.rot))));
    double Y[6] = {m.pos[0] - s.pos[0], m.pos[1] - s.pos[1], m.pos[2] - s.pos[2], NormAngleRad(ma.x - sa.x), NormAngleRad(ma.y - sa.y), NormAngleRad(ma.z - sa.z)};
    const int hrow[6] = {0, 1, 2, 3, 4, 5};
    UpdateEkfState<6>(s, K, Y, hrow);  // :427
    s.prev_gnss_timestamp = m.timestamp;
    s.updates += 1;
    return true;
}

void EkfGetCurrentState(EkfStateBlob& s, double o[26]) {
    const double ts = s.prev_timestamp;
    if (ts - s.ego_prev_timestamp < 1e-6) { std::memcpy(o, s.ego, 26 * sizeof(double)); return; }  // :786-789
    const V3 e = RotToVec(qtoR(getq(s.rot)));
    const double cy = std::cos(e.z), sy = std::sin(e.z), cp = std::cos(e.y), sp = std::sin(e.y), cr = std::cos(e.x), sr = std::sin(e.x);
    auto g2l = [&](double gx, double gy, double gz, double& lx, double& ly, double& lz) {  // ConvertGlobalToLocalVelocity
        lx = gx * (cy * cp) + gy * (sy * cp) + gz * (-sp);
        ly = gx * (cy * sp * sr - sy * cr) + gy * (sy * sp * sr + cy * cr) + gz * (cp * sr);
        lz = gx * (cy * sp * cr + sy * sr) + gy * (sy * sp * cr - cy * sr) + gz * (cp * cr);
    };
    o[0] = ts; o[1] = s.pos[0]; o[2] = s.pos[1]; o[3] = s.pos[2]; o[4] = e.x; o[5] = e.y; o[6] = e.z;
    o[7] = s.gyro[0]; o[8] = s.gyro[1]; o[9] = s.gyro[2];
    g2l(s.vel[0], s.vel[1], s.vel[2], o[10], o[11], o[12]);
    g2l(s.acc[0], s.acc[1], s.acc[2], o[13], o[14], o[15]);
    g2l(P_(s, S_X, S_X), P_(s, S_Y, S_Y), P_(s, S_Z, S_Z), o[16], o[17], o[18]);
    o[16] = std::fabs(o[16]); o[17] = std::fabs(o[17]); o[18] = std::fabs(o[18]);
    o[19] = std::sqrt(P_(s, S_X, S_X)); o[20] = std::sqrt(P_(s, S_Y, S_Y)); o[21] = std::sqrt(P_(s, S_Z, S_Z));
    o[22] = P_(s, S_ROLL, S_ROLL); o[23] = P_(s, S_PITCH, S_PITCH); o[24] = P_(s, S_YAW, S_YAW); o[25] = 0.0;
    std::memcpy(s.ego, o, 26 * sizeof(double));
    s.ego_prev_timestamp = ts;
}

}  // namespace orc
