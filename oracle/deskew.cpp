// ORACLE — TEST INFRASTRUCTURE ONLY (see deskew.hpp).
#include "deskew.hpp"

#include <cmath>

namespace orc {

namespace {
// pcl::getTransformation(x, y, z, roll, pitch, yaw) in float (pcm_matching.cpp:626,713,809)
struct Affine3f { float m[4][4]; };
Affine3f getTransformation(float x, float y, float z, float roll, float pitch, float yaw) {
    const float A = std::cos(yaw), B = std::sin(yaw), C = std::cos(pitch), D = std::sin(pitch), E = std::cos(roll), F = std::sin(roll);
    const float DE = D * E, DF = D * F;
    Affine3f t;
    t.m[0][0] = A * C; t.m[0][1] = A * DF - B * E; t.m[0][2] = B * F + A * DE; t.m[0][3] = x;
    t.m[1][0] = B * C; t.m[1][1] = A * E + B * DF; t.m[1][2] = B * DE - A * F; t.m[1][3] = y;
    t.m[2][0] = -D;    t.m[2][1] = C * F;          t.m[2][2] = C * E;          t.m[2][3] = z;
    t.m[3][0] = 0; t.m[3][1] = 0; t.m[3][2] = 0; t.m[3][3] = 1;
    return t;
}
}  // namespace

void ImuDeskewInfo(const double* stamp, const double* gyro, int n, DeskewTables& t) {
    t.imu_available = false;
    t.imu_time.assign(kImuQueueLength, 0.0);
    t.imu_rot_x.assign(kImuQueueLength, 0.0);
    t.imu_rot_y.assign(kImuQueueLength, 0.0);
    t.imu_rot_z.assign(kImuQueueLength, 0.0);
    int first = 0;
    while (first < n && stamp[first] < t.time_scan_cur - 0.01) ++first;  // :536-542 drop samples older than the scan
    if (first >= n) return;                                               // :544-547
    t.imu_pointer_cur = 0;
    for (int i = first; i < n; ++i) {
        const double cur = stamp[i];
        if (cur > t.time_scan_end + 0.01) break;  // :556
        if (t.imu_pointer_cur == 0) {             // :558-565
            t.imu_rot_x[0] = t.imu_rot_y[0] = t.imu_rot_z[0] = 0.0;
            t.imu_time[0] = cur;
            ++t.imu_pointer_cur;
            continue;
        }
        const int k = t.imu_pointer_cur;
        if (k >= kImuQueueLength) break;
        const double dt = cur - t.imu_time[k - 1];  // :572-576
        t.imu_rot_x[k] = t.imu_rot_x[k - 1] + gyro[3 * i] * dt;
        t.imu_rot_y[k] = t.imu_rot_y[k - 1] + gyro[3 * i + 1] * dt;
        t.imu_rot_z[k] = t.imu_rot_z[k - 1] + gyro[3 * i + 2] * dt;
        t.imu_time[k] = cur;
        ++t.imu_pointer_cur;
    }
    --t.imu_pointer_cur;  // :580
    if (t.imu_pointer_cur <= 0) return;
    t.imu_available = true;
}

void OdomDeskewInfo(const double sp[6], double start_stamp, const double ep[6], double end_stamp, DeskewTables& t) {
    t.odom_available = false;
    const Affine3f b = getTransformation((float)sp[0], (float)sp[1], (float)sp[2], (float)sp[3], (float)sp[4], (float)sp[5]);  // :625-627
    const Affine3f e = getTransformation((float)ep[0], (float)ep[1], (float)ep[2], (float)ep[3], (float)ep[4], (float)ep[5]);  // :712-714
    // affine_trans_begin.inverse() * affine_trans_end (:716): translation part = R_b^-1 (t_e - t_b); R_b is a rotation built
    // by getTransformation, Eigen's Affine inverse inverts the 3x3 linear part generally (cofactors, float)
    const float (*M)[4] = b.m;
    const float c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1], c01 = M[0][2] * M[2][1] - M[0][1] * M[2][2], c02 = M[0][1] * M[1][2] - M[0][2] * M[1][1];
    const float c10 = M[1][2] * M[2][0] - M[1][0] * M[2][2], c11 = M[0][0] * M[2][2] - M[0][2] * M[2][0], c12 = M[0][2] * M[1][0] - M[0][0] * M[1][2];
    const float c20 = M[1][0] * M[2][1] - M[1][1] * M[2][0], c21 = M[0][1] * M[2][0] - M[0][0] * M[2][1], c22 = M[0][0] * M[1][1] - M[0][1] * M[1][0];
    const float det = M[0][0] * c00 + M[0][1] * c10 + M[0][2] * c20;
    const float id = 1.0f / det;
    const float dx = e.m[0][3] - b.m[0][3], dy = e.m[1][3] - b.m[1][3], dz = e.m[2][3] - b.m[2][3];
    const float tx = (c00 * dx + c01 * dy + c02 * dz) * id, ty = (c10 * dx + c11 * dy + c12 * dz) * id, tz = (c20 * dx + c21 * dy + c22 * dz) * id;
    // InterpolateTfWithTime (localization_functions.hpp:216-241): translation * ratio, ratio narrowed to the vector's scalar
    const double dt_scan = t.time_scan_end - t.time_scan_cur;  // :719
    const double dt_trans = end_stamp - start_stamp;           // :720
    if (dt_trans == 0.0) { t.odom_incre_x = t.odom_incre_y = t.odom_incre_z = 0.f; }
    else {
        const float ratio = static_cast<float>(dt_scan / dt_trans);
        t.odom_incre_x = tx * ratio; t.odom_incre_y = ty * ratio; t.odom_incre_z = tz * ratio;
    }
    t.odom_available = true;
}

void DeskewPoint(const DeskewTables& t, const float in[3], double d_rel_time, float out[3]) {
    if (!t.imu_available) { out[0] = in[0]; out[1] = in[1]; out[2] = in[2]; return; }  // :781
    const double d_point_time = t.time_scan_cur + d_rel_time;                           // :783
    const float rx_end = t.imu_rot_x[t.imu_pointer_cur], ry_end = t.imu_rot_y[t.imu_pointer_cur], rz_end = t.imu_rot_z[t.imu_pointer_cur];  // :785-788
    // FindRotation :731-762
    float rx = 0, ry = 0, rz = 0;
    int front = 0;
    while (front < t.imu_pointer_cur) {
        if (d_point_time < t.imu_time[front]) break;
        ++front;
    }
    if (d_point_time > t.imu_time[front] || front == 0) {
        rx = t.imu_rot_x[front]; ry = t.imu_rot_y[front]; rz = t.imu_rot_z[front];
    } else {
        const int back = front - 1;
        const double rf = (d_point_time - t.imu_time[back]) / (t.imu_time[front] - t.imu_time[back]);
        const double rb = (t.imu_time[front] - d_point_time) / (t.imu_time[front] - t.imu_time[back]);
        rx = t.imu_rot_x[front] * rf + t.imu_rot_x[back] * rb;
        ry = t.imu_rot_y[front] * rf + t.imu_rot_y[back] * rb;
        rz = t.imu_rot_z[front] * rf + t.imu_rot_z[back] * rb;
    }
    // FindPosition :764-778
    float px = 0, py = 0, pz = 0;
    if (t.odom_available) {
        const float f_ratio = d_rel_time / (t.time_scan_end - t.time_scan_cur);
        px = f_ratio * t.odom_incre_x; py = f_ratio * t.odom_incre_y; pz = f_ratio * t.odom_incre_z;
    }
    (void)pz;
    const float rxe = rx - rx_end, rye = ry - ry_end, rze = rz - rz_end;  // :796-799
    const float pxe = px - t.odom_incre_x, pye = py - t.odom_incre_y;     // :801-803
    const float pze = rz - t.odom_incre_z;                                // :804 — uses f_rot_z_cur, not f_pos_z_cur (Q3)
    const Affine3f m = getTransformation(pxe, pye, pze, rxe, rye, rze);   // :809
    out[0] = m.m[0][0] * in[0] + m.m[0][1] * in[1] + m.m[0][2] * in[2] + m.m[0][3];  // :815-820
    out[1] = m.m[1][0] * in[0] + m.m[1][1] * in[1] + m.m[1][2] * in[2] + m.m[1][3];
    out[2] = m.m[2][0] * in[0] + m.m[2][1] * in[1] + m.m[2][2] * in[2] + m.m[2][3];
}

void DeskewPoints(const DeskewTables& t, const float* xyz, const float* rel_time, size_t n, float* out) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < static_cast<long>(n); ++i) DeskewPoint(t, xyz + 3 * i, static_cast<double>(rel_time[i]), out + 3 * i);
}

}  // namespace orc
