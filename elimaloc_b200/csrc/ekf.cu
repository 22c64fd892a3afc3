// 27-state EKF of ekf_localization on the GPU: predict / update as small fixed-size single-block kernels, state and
// covariance resident in HBM so that a scan's registration result can be folded in without a host round trip.
//   RunPredictionImu          ekf_localization/src/ekf_algorithm.cpp:167-316
//   RunGnssUpdate (PCM, PCM_INIT sources)                          :318-432
//   UpdateEkfState<M,K>       ekf_localization/include/ekf_algorithm.hpp:116-145
//   ComplementaryKalmanFilter ekf_algorithm.cpp:597-701 (its function-static memory lives in the state blob)
//   Check*Initialized/Stabilized ekf_algorithm.hpp:148-209
// Layout: one block of 768 threads; thread (i, j) owns P(i, j) (27 x 27 = 729); thread 0 does the scalar state algebra.
// Out of scope (off in config/localization.ini): RunPrediction, RunCanUpdate, ZUPT, NavSat/BESTPOS, CalibrateVehicleToImu.
#include "ekf.cuh"

#include "icp_device.cuh"

namespace elm {

namespace {

constexpr int N = 27;
enum { S_X = 0, S_Y, S_Z, S_ROLL, S_PITCH, S_YAW, S_VX, S_VY, S_VZ, S_ROLL_RATE, S_PITCH_RATE, S_YAW_RATE, S_AX, S_AY, S_AZ,
       S_B_ROLL_RATE, S_B_PITCH_RATE, S_B_YAW_RATE, S_B_AX, S_B_AY, S_B_AZ, S_G_X, S_G_Y, S_G_Z, S_IMU_ROLL, S_IMU_PITCH, S_IMU_YAW };
constexpr double kInitCov = 100.0;
constexpr double kPi = 3.14159265358979323846;
constexpr double kDeg = kPi / 180.0;

struct Q4 { double w, x, y, z; };
__device__ Q4 q_mul(const Q4& a, const Q4& b) {
    return Q4{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
__device__ Q4 q_unit(const Q4& q) {
    const double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return Q4{q.w / n, q.x / n, q.y / n, q.z / n};
}
__device__ void q_to_R(const Q4& q, double* R) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ Q4 q_from_R(const double* m) {  // Shepperd, as Eigen's Quaterniond(Matrix3d)
    Q4 q;
    double t = m[0] + m[4] + m[8];
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t; t = 0.5 / t;
        q.x = (m[7] - m[5]) * t; q.y = (m[2] - m[6]) * t; q.z = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
        double v[3];
        v[i] = 0.5 * t; t = 0.5 / t;
        q.w = (m[3 * k + j] - m[3 * j + k]) * t;
        v[j] = (m[3 * j + i] + m[3 * i + j]) * t;
        v[k] = (m[3 * k + i] + m[3 * i + k]) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}
__device__ Q4 q_from_rotvec(double rx, double ry, double rz) {  // Quaterniond(AngleAxisd(|r|, r/|r|)); zero vector -> identity
    const double n2 = rx * rx + ry * ry + rz * rz;
    const double n = sqrt(n2);
    double ax = rx, ay = ry, az = rz;
    if (n2 > 0.0) { ax /= n; ay /= n; az /= n; }
    const double h = 0.5 * n, s = sin(h);
    return Q4{cos(h), s * ax, s * ay, s * az};
}
__device__ void q_rotate(const Q4& q, const double* v, double* o) {  // Quaterniond * Vector3d
    double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    o[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
    o[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
    o[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}
__device__ double norm_angle(double a) {
    while (a > kPi) a -= kPi * 2.;
    while (a < -kPi) a += kPi * 2.;
    return a;
}
__device__ void rot_to_vec(const double* R, double* a) {  // lfun.hpp RotToVec
    if (fabs(R[6]) > 0.998) {
        a[2] = atan2(-R[5], R[4]);
        a[1] = kPi / 2 * (R[6] >= 0 ? 1 : -1);
        a[0] = 0;
    } else {
        a[1] = asin(-R[6]);
        const double c = cos(a[1]);
        a[0] = atan2(R[7] / c, R[8] / c);
        a[2] = atan2(R[3] / c, R[0] / c);
    }
    for (int i = 0; i < 3; ++i) a[i] = fmod(a[i] + kPi, 2 * kPi) - kPi;
}
// Exp(omega) (lfun.hpp Exp) and d Exp / d gyro (PartialDerivativeRotWrtGyro)
__device__ void exp_so3(const double* w, double* R, double dt, double* dR) {
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    for (int i = 0; i < 9; ++i) { R[i] = (i % 4 == 0) ? 1.0 : 0.0; dR[i] = 0.0; }
    if (th < 1e-5) return;
    const double a[3] = {w[0] / th, w[1] / th, w[2] / th};
    const double K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    double KK[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) KK[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
    const double s = sin(th), c1 = 1 - cos(th);
    const double ca = c1 / (th * th), cb = (th - s) / (th * th * th);
    for (int i = 0; i < 9; ++i) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + s * K[i] + c1 * KK[i];
        dR[i] = dt * (I + ca * K[i] + cb * KK[i]);
    }
}
__device__ Q4 get_q(const double* r) { return Q4{r[0], r[1], r[2], r[3]}; }
__device__ void set_q(double* r, const Q4& q) { r[0] = q.w; r[1] = q.x; r[2] = q.y; r[3] = q.z; }

__device__ void check_flags(elm_ekf_state* s, bool yaw, bool init, bool rot, bool stab) {
    const double* P = s->P;
    const double sr = sqrt(P[S_ROLL * N + S_ROLL]), sp = sqrt(P[S_PITCH * N + S_PITCH]), sy = sqrt(P[S_YAW * N + S_YAW]);
    const double sx = sqrt(P[S_X * N + S_X]), syy = sqrt(P[S_Y * N + S_Y]);
    if (yaw) s->yaw_initialized = sy < 5.0 * kDeg;
    if (init) s->state_initialized = sr < 5.0 * kDeg && sp < 5.0 * kDeg && sy < 5.0 * kDeg && sx < 1.0 && syy < 1.0;
    if (rot) s->rotation_stabilized = sr < 0.2 * kDeg && sp < 0.2 * kDeg && sy < 0.2 * kDeg;
    if (stab) s->state_stabilized = sr < 0.2 * kDeg && sp < 0.2 * kDeg && sy < 0.2 * kDeg && sx < 0.5 && syy < 0.5;
}

// state part of UpdateEkfState (ekf_alg.hpp:124-141); du = K * Y
__device__ void apply_state_update(elm_ekf_state* s, const double* du) {
    for (int k = 0; k < 3; ++k) {
        s->pos[k] += du[S_X + k]; s->vel[k] += du[S_VX + k]; s->gyro[k] += du[S_ROLL_RATE + k]; s->acc[k] += du[S_AX + k];
        s->bg[k] += du[S_B_ROLL_RATE + k]; s->ba[k] += du[S_B_AX + k]; s->grav[k] += du[S_G_X + k];
    }
    set_q(s->rot, q_unit(q_mul(get_q(s->rot), q_from_rotvec(du[3], du[4], du[5]))));
    set_q(s->imu_rot, q_unit(q_mul(get_q(s->imu_rot), q_from_rotvec(du[24], du[25], du[26]))));
}

// Shared by both kernels: given K (N x M, shared), Y (M) and the selector rows of H, update state and P = P - K H P.
template <int M>
__device__ void update_block(elm_ekf_state* s, const double* sK, const double* sY, const int* hrow, double (*sHP)[N]) {
    const int tid = threadIdx.x;
    if (tid < M * N) sHP[tid / N][tid % N] = s->P[hrow[tid / N] * N + tid % N];
    __syncthreads();
    if (tid == 0) {
        double du[N];
        for (int i = 0; i < N; ++i) { double a = 0; for (int k = 0; k < M; ++k) a += sK[i * M + k] * sY[k]; du[i] = a; }
        apply_state_update(s, du);
    }
    if (tid < N * N) {
        const int i = tid / N, j = tid % N;
        double a = 0;
#pragma unroll
        for (int k = 0; k < M; ++k) a += sK[i * M + k] * sHP[k][j];
        s->P[tid] -= a;
    }
    __syncthreads();
}

// ComplementaryKalmanFilter (ekf_alg.cpp:597-701); every thread of the block calls it
__device__ void complementary_filter(elm_ekf_state* s, double timestamp, const double* acc_in, double* sK, double* sY, double (*sHP)[N],
                                     double* sSinv, int* sflag) {
    const int tid = threadIdx.x;
    if (tid == 0) {
        *sflag = 0;
        const double am[3] = {acc_in[0] - s->ba[0], acc_in[1] - s->ba[1], acc_in[2] - s->ba[2]};
        const Q4 rot = get_q(s->rot);
        const double n2 = rot.w * rot.w + rot.x * rot.x + rot.y * rot.y + rot.z * rot.z;
        const Q4 inv{rot.w / n2, -rot.x / n2, -rot.y / n2, -rot.z / n2};
        double vl[3];
        q_rotate(inv, s->vel, vl);
        const double centripetal = vl[0] * s->gyro[2];
        if (!s->ckf_has_prev) { s->ckf_prev_vel_local_x = vl[0]; s->ckf_prev_time = timestamp; s->ckf_has_prev = 1; }
        const double dt = timestamp - s->ckf_prev_time;
        if (!(dt < 1e-6)) {
            const double est_acc_x = (vl[0] - s->ckf_prev_vel_local_x) / dt;
            s->ckf_prev_vel_local_x = vl[0];
            s->ckf_prev_time = timestamp;
            double c[3] = {am[0], am[1] - centripetal, am[2]};
            if (s->rotation_stabilized) c[0] -= est_acc_x;
            const double acc_diff = sqrt(am[0] * am[0] + am[1] * am[1] + am[2] * am[2]) -
                                    sqrt(s->grav[0] * s->grav[0] + s->grav[1] * s->grav[1] + s->grav[2] * s->grav[2]);
            const double cn2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
            if (cn2 > 0.0) { const double cn = sqrt(cn2); c[0] /= cn; c[1] /= cn; c[2] /= cn; }
            const double z0 = atan2(c[1], c[2]), z1 = -asin(c[0]);
            double R[9], rpy[3];
            q_to_R(rot, R);
            rot_to_vec(R, rpy);
            sY[0] = norm_angle(z0 - rpy[0]);
            sY[1] = norm_angle(z1 - rpy[1]);
            double base = 1.0 * kDeg;
            if (!s->state_initialized) base = 10.0 * kDeg;
            const double cu = fabs(centripetal) / 9.81 * 10.0, lu = fabs(est_acc_x) / 9.81 * 10.0, au = fabs(acc_diff) / 9.81 * 10.0;
            const double lat = 1.0 + au + cu, lon = 1.0 + au + lu;
            const double R0 = fmax((base * lat) * (base * lat), kDeg * kDeg), R1 = fmax((base * lon) * (base * lon), kDeg * kDeg);
            const double* P = s->P;
            const double S00 = P[S_ROLL * N + S_ROLL] + R0, S01 = P[S_ROLL * N + S_PITCH], S10 = P[S_PITCH * N + S_ROLL], S11 = P[S_PITCH * N + S_PITCH] + R1;
            const double det = S00 * S11 - S01 * S10;
            sSinv[0] = S11 / det; sSinv[1] = -S01 / det; sSinv[2] = -S10 / det; sSinv[3] = S00 / det;
            *sflag = 1;
        }
    }
    __syncthreads();
    if (!*sflag) return;  // block-uniform
    if (tid < N) {
        const double a = s->P[tid * N + S_ROLL], b = s->P[tid * N + S_PITCH];
        sK[tid * 2] = a * sSinv[0] + b * sSinv[2];
        sK[tid * 2 + 1] = a * sSinv[1] + b * sSinv[3];
    }
    __syncthreads();
    __shared__ int hrow[2];
    if (tid == 0) { hrow[0] = S_ROLL; hrow[1] = S_PITCH; }
    __syncthreads();
    update_block<2>(s, sK, sY, hrow, sHP);
}

}  // namespace

__global__ void __launch_bounds__(768) ekf_predict_imu_kernel(elm_ekf_state* s, elm_ekf_config c, double t, double gx, double gy, double gz,
                                                              double ax, double ay, double az) {
    __shared__ double sF[N * N], sFP[N * N], sQ[N];
    __shared__ double sK[N * 6], sY[6], sHP[6][N], sSinv[4];
    __shared__ int s_mode, s_flag;
    const int tid = threadIdx.x;
    const double acc_in[3] = {ax, ay, az};
    if (tid < N * N) sF[tid] = (tid / N == tid % N) ? 1.0 : 0.0;
    if (tid < N) sQ[tid] = 0.0;
    __syncthreads();
    if (tid == 0) {
        int mode = 2;  // 0: return false; 1: complementary filter only; 2: full prediction
        if (s->reset_for_init_prediction) { s->prev_timestamp = t; s->reset_for_init_prediction = 0; mode = 0; }  // :182-187
        else if (s->pcm_init_on_going) { s->prev_timestamp = t; mode = 0; }                                      // :189-194
        else {
            check_flags(s, false, false, true, false);                                                            // :196
            if (!s->state_initialized) {                                                                          // :198-208
                s->prev_timestamp = t;
                mode = (s->yaw_initialized && c.use_complementary_filter) ? 1 : 0;
            } else if (fabs(t - s->prev_timestamp) < 1e-6) mode = 0;                                              // :210-213
        }
        if (mode == 2) {
            const double dt = t - s->prev_timestamp;
            const Q4 rot_prev = get_q(s->rot);
            double G[9];
            q_to_R(rot_prev, G);                                                                      // :231
            const double cg[3] = {gx - s->bg[0], gy - s->bg[1], gz - s->bg[2]};                        // :234
            const double om[3] = {cg[0] * dt, cg[1] * dt, cg[2] * dt};
            double E[9], dE[9];
            exp_so3(om, E, dt, dE);
            set_q(s->rot, q_unit(q_mul(rot_prev, q_from_R(E))));                                      // :235-236
            const double ca[3] = {ax - s->ba[0], ay - s->ba[1], az - s->ba[2]};                        // :239
            double ag[3];
            for (int k = 0; k < 3; ++k) ag[k] = G[3 * k] * ca[0] + G[3 * k + 1] * ca[1] + G[3 * k + 2] * ca[2] - s->grav[k];  // :240
            for (int k = 0; k < 3; ++k) s->pos[k] += s->vel[k] * dt + 0.5 * ag[k] * dt * dt;           // :243
            for (int k = 0; k < 3; ++k) { s->vel[k] += ag[k] * dt; s->gyro[k] = cg[k]; s->acc[k] = ag[k]; }  // :244-248
            for (int k = 0; k < 3; ++k) {                                                              // Q :256-272
                sQ[S_X + k] = c.state_std_pos_m * c.state_std_pos_m * dt * dt;
                sQ[S_ROLL + k] = (c.state_std_rot_deg * kDeg) * (c.state_std_rot_deg * kDeg) * dt * dt;
                sQ[S_VX + k] = c.state_std_vel_mps * c.state_std_vel_mps * dt * dt;
                sQ[S_ROLL_RATE + k] = (c.imu_std_gyro_dps * kDeg) * (c.imu_std_gyro_dps * kDeg) * dt * dt;
                sQ[S_AX + k] = c.imu_std_acc_mps * c.imu_std_acc_mps * dt * dt;
                sQ[S_B_ROLL_RATE + k] = c.imu_bias_cov_gyro * c.imu_bias_cov_gyro * dt * dt;
                sQ[S_B_AX + k] = c.imu_bias_cov_acc * c.imu_bias_cov_acc * dt * dt;
                sQ[S_G_X + k] = c.imu_bias_cov_acc * c.imu_bias_cov_acc * dt * dt;
                sQ[S_IMU_ROLL + k] = (c.state_std_rot_deg * kDeg) * (c.state_std_rot_deg * kDeg) * dt * dt;
            }
            for (int a = 0; a < 3; ++a) {                                                              // F :275-297
                sF[(S_X + a) * N + S_VX + a] = dt;
                sF[(S_ROLL_RATE + a) * N + S_B_ROLL_RATE + a] = -1.0;
                for (int b = 0; b < 3; ++b) {
                    sF[(S_X + a) * N + S_B_AX + b] = -0.5 * G[3 * a + b] * dt * dt;
                    sF[(S_ROLL + a) * N + S_B_ROLL_RATE + b] = -dE[3 * a + b];
                    sF[(S_VX + a) * N + S_B_AX + b] = -G[3 * a + b] * dt;
                    sF[(S_AX + a) * N + S_B_AX + b] = -G[3 * a + b];
                }
            }
            if (c.imu_estimate_gravity) {
                sF[S_Z * N + S_G_Z] = -0.5 * dt * dt;
                sF[S_VZ * N + S_G_Z] = -dt;
                sF[S_AZ * N + S_G_Z] = -1.0;
            }
            s->prev_timestamp = t;
            s->predictions += 1;
        }
        s_mode = mode;
    }
    __syncthreads();
    const int mode = s_mode;
    if (mode == 0) return;
    if (mode == 2) {  // P = F P F^T + Q (:300)
        if (tid < N * N) {
            const int i = tid / N, j = tid % N;
            double a = 0;
            for (int k = 0; k < N; ++k) a += sF[i * N + k] * s->P[k * N + j];
            sFP[tid] = a;
        }
        __syncthreads();
        if (tid < N * N) {
            const int i = tid / N, j = tid % N;
            double a = 0;
            for (int k = 0; k < N; ++k) a += sFP[i * N + k] * sF[j * N + k];
            s->P[tid] = a + (i == j ? sQ[i] : 0.0);
        }
        __syncthreads();
    }
    if (c.use_complementary_filter) complementary_filter(s, t, acc_in, sK, sY, sHP, sSinv, &s_flag);  // :204, :312
}

// RunGnssUpdate for the PCM / PCM_INIT sources (:318-432) by the whole block; m may live in shared memory.  skip != 0: no update.
struct UpdateShared {
    double sK[N * 6], sY[6], sHP[6][N], sS[6][12];
    int s_go, hrow[6];
};
__device__ void update_pose_block(elm_ekf_state* s, const elm_ekf_config& c, const elm_ekf_measurement& m, int skip, UpdateShared& sh) {
    double* const sK = sh.sK; double* const sY = sh.sY; double (*const sHP)[N] = sh.sHP; double (*const sS)[12] = sh.sS;
    int& s_go = sh.s_go; int* const hrow = sh.hrow;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_go = 1;
        if (skip) {
            s_go = 0;
        } else if (m.source == 4) {  // PCM_INIT: hard reset (:324-349)
            for (int k = 0; k < 3; ++k) { s->pos[k] = m.pos[k]; s->vel[k] = s->gyro[k] = s->acc[k] = s->bg[k] = s->ba[k] = s->grav[k] = 0.0; }
            for (int k = 0; k < 4; ++k) s->rot[k] = m.rot[k];
            s->grav[2] = c.imu_gravity;
            s->state_initialized = 1; s->yaw_initialized = 1; s->pcm_init_on_going = 1;
            s_go = 2;
        } else {
            check_flags(s, true, true, true, true);                                       // :351-354
            if (s->pcm_init_on_going && m.source == 3) {                                  // :357-364
                if (s->pcm_update_count > 10) s->pcm_init_on_going = 0;
                s->pcm_update_count++;
            }
            // S = H P H^T + R, H = [I6 0]; inverse by Gauss-Jordan with partial pivoting (stands in for Matrix6d::inverse, :400-403)
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { sS[i][j] = s->P[i * N + j]; sS[i][6 + j] = (i == j) ? 1.0 : 0.0; }
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { sS[i][j] += m.pos_cov[3 * i + j]; sS[3 + i][3 + j] += m.rot_cov[3 * i + j]; }
            for (int col = 0; col < 6; ++col) {
                int p = col;
                for (int r = col + 1; r < 6; ++r) if (fabs(sS[r][col]) > fabs(sS[p][col])) p = r;
                if (p != col) for (int j = 0; j < 12; ++j) { const double tmp = sS[p][j]; sS[p][j] = sS[col][j]; sS[col][j] = tmp; }
                const double d = 1.0 / sS[col][col];
                for (int j = 0; j < 12; ++j) sS[col][j] *= d;
                for (int r = 0; r < 6; ++r) {
                    if (r == col) continue;
                    const double f = sS[r][col];
                    if (f == 0.0) continue;
                    for (int j = 0; j < 12; ++j) sS[r][j] -= f * sS[col][j];
                }
            }
            // Y = [pos residual ; Euler residual] (:406-410, lfun.hpp CalEulerResidualFromQuat)
            double Rs[9], Rm[9], as[3], am[3];
            q_to_R(q_unit(get_q(s->rot)), Rs);
            q_to_R(q_unit(get_q(m.rot)), Rm);
            rot_to_vec(Rs, as);
            rot_to_vec(Rm, am);
            for (int k = 0; k < 3; ++k) { sY[k] = m.pos[k] - s->pos[k]; sY[3 + k] = norm_angle(am[k] - as[k]); }
            for (int k = 0; k < 6; ++k) hrow[k] = k;
            s->prev_gnss_timestamp = m.timestamp;
            s->updates += 1;
        }
    }
    __syncthreads();
    if (s_go == 0) return;
    if (s_go == 2) {  // covariance part of the PCM_INIT reset: top-left 15 x 15 = 100 I
        if (tid < N * N) {
            const int i = tid / N, j = tid % N;
            if (i <= S_AZ && j <= S_AZ) s->P[tid] = (i == j) ? kInitCov : 0.0;
        }
        return;
    }
    if (tid < N * 6) {  // K = P H^T S^-1
        const int i = tid / 6, j = tid % 6;
        double a = 0;
        for (int k = 0; k < 6; ++k) a += s->P[i * N + k] * sS[k][6 + j];
        sK[tid] = a;
    }
    __syncthreads();
    update_block<6>(s, sK, sY, hrow, sHP);  // :427
}

__global__ void __launch_bounds__(768) ekf_update_pose_kernel(elm_ekf_state* s, elm_ekf_config c, elm_ekf_measurement m) {
    __shared__ UpdateShared sh;
    update_pose_block(s, c, m, 0, sh);
}

// ---- the scan's registration result folded into the filter without a host round trip --------------------------------------
// What happens between RunRegister and RunGnssUpdate in the two ROS nodes, on the IcpState that still sits in HBM:
//   pcm_matching.cpp:283-299   success gate, icp_ego_pose = icp_lidar_pose * tf_ego_to_lidar^-1
//   PublishPcmOdom :1047-1101  position, Quaterniond(rotation), covariance shaping (NormalizeCovariance, pcm_matching.hpp:247-273)
//   CallbackPcmOdom            ekf_localization.cpp:147-179 (the two 3 x 3 blocks of the message covariance)
//   GnssTimeCompensation       ekf_localization.cpp:323-394 against the ring of EgoStates the prediction kernel's companion
//                              (ekf_ring_push_kernel = PublishInThread's deque, :398-410) keeps in HBM
//   RunGnssUpdate (PCM)        ekf_algorithm.cpp:318-432
__device__ double angle_diff(double ref, double rel) {  // lfun.hpp AngleDiffRad
    double d = rel - ref;
    while (d > kPi) d -= 2 * kPi;
    while (d < -kPi) d += 2 * kPi;
    return d;
}
__global__ void __launch_bounds__(768) ekf_update_from_icp_kernel(elm_ekf_state* s, elm_ekf_config c, const IcpState* icp, EkfIcpParams p, EkfRing ring) {
    __shared__ UpdateShared sh;
    __shared__ elm_ekf_measurement m;
    __shared__ int s_skip;
    if (threadIdx.x == 0) {
        int skip = 0;
        // elm_register_fetch's success rule (registration.cpp:352-356, 405-417)
        if (p.trivial || icp->overlap_fail || icp->fitness > p.max_fitness) skip = 1;
        if (!skip) {
            double E[12];  // icp_lidar_pose * tf_ego_to_lidar^-1, rows 0..2
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 4; ++j) {
                    double a = 0.0;
                    for (int k = 0; k < 4; ++k) a += icp->T[4 * i + k] * p.T_lidar_to_ego[4 * k + j];
                    E[4 * i + j] = a;
                }
            const double R[9] = {E[0], E[1], E[2], E[4], E[5], E[6], E[8], E[9], E[10]};
            const Q4 q = q_from_R(R);
            double pc[36];
            for (int i = 0; i < 36; ++i) pc[i] = 0.0;
            shape_pcm_covariance_hd(R, icp->local_cov, icp->fitness, pc);
            m.timestamp = p.stamp; m.source = 3; m.reserved = 0;
            m.pos[0] = E[3]; m.pos[1] = E[7]; m.pos[2] = E[11];
            m.rot[0] = q.w; m.rot[1] = q.x; m.rot[2] = q.y; m.rot[3] = q.z;
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { m.pos_cov[3 * i + j] = pc[6 * i + j]; m.rot_cov[3 * i + j] = pc[6 * (i + 3) + j + 3]; }
            // GnssTimeCompensation
            const int head = ring.meta[0], cnt = ring.meta[1];
            if (cnt < 1) skip = 1;
            else {
                const double* cur = ring.e + 8 * static_cast<size_t>((head + cnt - 1) % ring.cap);
                const double* front = ring.e + 8 * static_cast<size_t>(head);
                if (front[0] > m.timestamp) skip = 1;
                else {
                    const double* closest = front;
                    for (int k = 0; k < cnt; ++k) {
                        const double* st = ring.e + 8 * static_cast<size_t>((head + k) % ring.cap);
                        closest = st;
                        if (st[0] > m.timestamp) break;
                    }
                    const double gap = cur[0] - m.timestamp;
                    if (gap > 0.0) {
                        double d[6] = {0, 0, 0, 0, 0, 0};
                        if (fabs(cur[0] - closest[0]) > 1e-5) {
                            const double ratio = gap / (cur[0] - closest[0]);
                            for (int k = 0; k < 3; ++k) d[k] = (cur[1 + k] - closest[1 + k]) * ratio;
                            for (int k = 0; k < 3; ++k) d[3 + k] = angle_diff(closest[4 + k], cur[4 + k]) * ratio;
                        }
                        m.timestamp = cur[0];
                        for (int k = 0; k < 3; ++k) m.pos[k] += d[k];
                        // AngleAxis(yaw, Z) * AngleAxis(pitch, Y) * AngleAxis(roll, X)
                        const Q4 qz{cos(0.5 * d[5]), 0, 0, sin(0.5 * d[5])}, qy{cos(0.5 * d[4]), 0, sin(0.5 * d[4]), 0}, qx{cos(0.5 * d[3]), sin(0.5 * d[3]), 0, 0};
                        const Q4 qn = q_unit(q_mul(q, q_mul(q_mul(qz, qy), qx)));
                        m.rot[0] = qn.w; m.rot[1] = qn.x; m.rot[2] = qn.y; m.rot[3] = qn.z;
                    }
                }
            }
        }
        s_skip = skip;
    }
    __syncthreads();
    update_pose_block(s, c, m, s_skip, sh);
}

// PublishInThread's deque of EgoStates (ekf_localization.cpp:398-410) as a ring in HBM: after every RunPredictionImu the
// node takes GetCurrentState and appends it unless the stamp did not advance; a stamp that went backwards clears the queue;
// at most 1000 entries.  Entry = {timestamp, x, y, z, roll, pitch, yaw, 0}.
__global__ void ekf_ring_push_kernel(elm_ekf_state* s, EkfRing ring) {
    if (threadIdx.x != 0) return;
    double ego[26];
    ekf_current_state(*s, ego);
    int head = ring.meta[0], cnt = ring.meta[1];
    const double back_t = cnt > 0 ? ring.e[8 * static_cast<size_t>((head + cnt - 1) % ring.cap)] : 0.0;
    if (cnt < 1 || back_t + 1e-5 < ego[0]) {
        if (cnt == ring.cap) head = (head + 1) % ring.cap; else ++cnt;
        double* e = ring.e + 8 * static_cast<size_t>((head + cnt - 1) % ring.cap);
        for (int k = 0; k < 7; ++k) e[k] = ego[k];
        e[7] = 0.0;
    } else if (back_t > ego[0]) {
        cnt = 0;
    }
    ring.meta[0] = head; ring.meta[1] = cnt;
}

cudaError_t launch_ekf_predict_imu(elm_ekf_state* s, const elm_ekf_config& c, double t, const double g[3], const double a[3], cudaStream_t st) {
    ekf_predict_imu_kernel<<<1, 768, 0, st>>>(s, c, t, g[0], g[1], g[2], a[0], a[1], a[2]);
    return cudaGetLastError();
}
cudaError_t launch_ekf_update_pose(elm_ekf_state* s, const elm_ekf_config& c, const elm_ekf_measurement& m, cudaStream_t st) {
    ekf_update_pose_kernel<<<1, 768, 0, st>>>(s, c, m);
    return cudaGetLastError();
}

cudaError_t launch_ekf_update_from_icp(elm_ekf_state* s, const elm_ekf_config& c, const IcpState* icp, const EkfIcpParams& p, const EkfRing& ring, cudaStream_t st) {
    ekf_update_from_icp_kernel<<<1, 768, 0, st>>>(s, c, icp, p, ring);
    return cudaGetLastError();
}
cudaError_t launch_ekf_ring_push(elm_ekf_state* s, const EkfRing& ring, cudaStream_t st) {
    ekf_ring_push_kernel<<<1, 32, 0, st>>>(s, ring);
    return cudaGetLastError();
}

// ---- host-side pieces (no arithmetic on P) ---------------------------------------------------------------------------------
void ekf_init_state(const elm_ekf_config& c, elm_ekf_state& s) {  // EkfAlgorithm::Init, ekf_alg.cpp:22-66
    std::memset(&s, 0, sizeof s);
    s.pos[0] = c.ekf_init_x_m; s.pos[1] = c.ekf_init_y_m; s.pos[2] = c.ekf_init_z_m;
    // AngleAxis(yaw, Z) * AngleAxis(pitch, Y) * AngleAxis(roll, X)
    const double y = 0.5 * c.ekf_init_yaw_deg * kDeg, p = 0.5 * c.ekf_init_pitch_deg * kDeg, r = 0.5 * c.ekf_init_roll_deg * kDeg;
    const double qz[4] = {std::cos(y), 0, 0, std::sin(y)}, qy[4] = {std::cos(p), 0, std::sin(p), 0}, qx[4] = {std::cos(r), std::sin(r), 0, 0};
    auto mul = [](const double* a, const double* b, double* o) {
        o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
        o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
        o[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
        o[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
    };
    double zy[4];
    mul(qz, qy, zy);
    mul(zy, qx, s.rot);
    s.imu_rot[0] = 1.0;
    s.grav[2] = c.imu_gravity;
    for (int i = 0; i < N; ++i) s.P[i * N + i] = kInitCov;
    for (int k = 0; k < 3; ++k) {
        s.P[(S_B_ROLL_RATE + k) * N + S_B_ROLL_RATE + k] = c.imu_bias_cov_gyro;
        s.P[(S_B_AX + k) * N + S_B_AX + k] = c.imu_bias_cov_acc;
        s.P[(S_G_X + k) * N + S_G_X + k] = c.imu_bias_cov_acc;
        s.P[(S_IMU_ROLL + k) * N + S_IMU_ROLL + k] = c.imu_bias_cov_gyro;
    }
    s.reset_for_init_prediction = 1;
}

// EkfAlgorithm::GetCurrentState (ekf_alg.cpp:778-833) on a state blob fetched from the device; returns true when the
// cached previous EgoState was returned unchanged (delta_time < 1e-6)
__host__ __device__ bool ekf_current_state(elm_ekf_state& s, double o[26]) {
    const double ts = s.prev_timestamp;
    if (ts - s.ego_prev_timestamp < 1e-6) { for (int i = 0; i < 26; ++i) o[i] = s.ego[i]; return true; }
    const double* q = s.rot;
    const double tx = 2 * q[1], ty = 2 * q[2], tz = 2 * q[3];
    const double twx = tx * q[0], twy = ty * q[0], twz = tz * q[0], txx = tx * q[1], txy = ty * q[1], txz = tz * q[1], tyy = ty * q[2], tyz = tz * q[2], tzz = tz * q[3];
    const double R00 = 1 - (tyy + tzz), R10 = txy + twz, R11 = 1 - (txx + tzz), R12 = tyz - twx, R20 = txz - twy, R21 = tyz + twx, R22 = 1 - (txx + tyy);
    double e[3];
    if (std::fabs(R20) > 0.998) { e[2] = std::atan2(-R12, R11); e[1] = kPi / 2 * (R20 >= 0 ? 1 : -1); e[0] = 0; }
    else {
        e[1] = std::asin(-R20);
        const double c = std::cos(e[1]);
        e[0] = std::atan2(R21 / c, R22 / c);
        e[2] = std::atan2(R10 / c, R00 / c);
    }
    for (int i = 0; i < 3; ++i) e[i] = std::fmod(e[i] + kPi, 2 * kPi) - kPi;
    const double cy = std::cos(e[2]), sy = std::sin(e[2]), cp = std::cos(e[1]), sp = std::sin(e[1]), cr = std::cos(e[0]), sr = std::sin(e[0]);
    auto g2l = [&](double gx, double gy, double gz, double& lx, double& ly, double& lz) {
        lx = gx * (cy * cp) + gy * (sy * cp) + gz * (-sp);
        ly = gx * (cy * sp * sr - sy * cr) + gy * (sy * sp * sr + cy * cr) + gz * (cp * sr);
        lz = gx * (cy * sp * cr + sy * sr) + gy * (sy * sp * cr - cy * sr) + gz * (cp * cr);
    };
    const double Pxx = s.P[S_X * N + S_X], Pyy = s.P[S_Y * N + S_Y], Pzz = s.P[S_Z * N + S_Z];
    o[0] = ts; o[1] = s.pos[0]; o[2] = s.pos[1]; o[3] = s.pos[2]; o[4] = e[0]; o[5] = e[1]; o[6] = e[2];
    o[7] = s.gyro[0]; o[8] = s.gyro[1]; o[9] = s.gyro[2];
    g2l(s.vel[0], s.vel[1], s.vel[2], o[10], o[11], o[12]);
    g2l(s.acc[0], s.acc[1], s.acc[2], o[13], o[14], o[15]);
    g2l(Pxx, Pyy, Pzz, o[16], o[17], o[18]);
    o[16] = std::fabs(o[16]); o[17] = std::fabs(o[17]); o[18] = std::fabs(o[18]);
    o[19] = std::sqrt(Pxx); o[20] = std::sqrt(Pyy); o[21] = std::sqrt(Pzz);
    o[22] = s.P[S_ROLL * N + S_ROLL]; o[23] = s.P[S_PITCH * N + S_PITCH]; o[24] = s.P[S_YAW * N + S_YAW]; o[25] = 0.0;
    for (int i = 0; i < 26; ++i) s.ego[i] = o[i];
    s.ego_prev_timestamp = ts;
    return false;
}

}  // namespace elm
