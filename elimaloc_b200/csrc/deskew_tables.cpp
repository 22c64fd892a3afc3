// Host-side table builders of the deskew stage: see deskew_tables.hpp.  Reference: pcm_matching/src/pcm_matching.cpp
// (ImuDeskewInfo :533-585, OdomDeskewInfo :587-729), localization_functions.hpp:216-241 (InterpolateTfWithTime).
// Third-party pieces restated from their published formulas: tf::Matrix3x3::getRPY / tf::Quaternion::setRPY (Bullet),
// pcl::getTransformation (pcl/common/eigen.h: Rz(yaw) Ry(pitch) Rx(roll), float).
#include "deskew_tables.hpp"

#include <cmath>

namespace elm {

namespace {

struct Rigid3f { float r[3][3]; float t[3]; };

// pcl::getTransformation(x, y, z, roll, pitch, yaw) — every argument narrowed to float at the call (pcm_matching.cpp:625-627)
Rigid3f get_transformation(double x, double y, double z, double roll, double pitch, double yaw) {
    const float fr = static_cast<float>(roll), fp = static_cast<float>(pitch), fy = static_cast<float>(yaw);
    const float A = std::cos(fy), B = std::sin(fy), C = std::cos(fp), D = std::sin(fp), E = std::cos(fr), F = std::sin(fr);
    const float DE = D * E, DF = D * F;
    Rigid3f m;
    m.r[0][0] = A * C; m.r[0][1] = A * DF - B * E; m.r[0][2] = B * F + A * DE;
    m.r[1][0] = B * C; m.r[1][1] = A * E + B * DF; m.r[1][2] = B * DE - A * F;
    m.r[2][0] = -D;    m.r[2][1] = C * F;          m.r[2][2] = C * E;
    m.t[0] = static_cast<float>(x); m.t[1] = static_cast<float>(y); m.t[2] = static_cast<float>(z);
    return m;
}

// index of the first message whose stamp is not older than t; the last message when all are older
// (the `for ... if (stamp < t) continue; else break;` loops of :611-618 and :641-648)
size_t first_not_older(const double* stamp, size_t begin, size_t n, double t) {
    size_t i = begin;
    for (; i < n; ++i)
        if (!(stamp[i] < t)) return i;
    return n - 1;
}

}  // namespace

void quat_to_rpy(const double q[4], double& roll, double& pitch, double& yaw) {
    // tf::Matrix3x3::setRotation
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double d = x * x + y * y + z * z + w * w;
    const double s = 2.0 / d;
    const double xs = x * s, ys = y * s, zs = z * s;
    const double wx = w * xs, wy = w * ys, wz = w * zs, xx = x * xs, xy = x * ys, xz = x * zs, yy = y * ys, yz = y * zs, zz = z * zs;
    const double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
    // tf::Matrix3x3::getEulerYPR, solution 1
    if (std::fabs(m20) >= 1.0) {
        yaw = 0.0;
        const double delta = std::atan2(m21, m22);
        const double kHalfPi = 1.57079632679489661923;
        pitch = (m20 < 0) ? kHalfPi : -kHalfPi;
        roll = delta;
    } else {
        pitch = -std::asin(m20);
        const double c = std::cos(pitch);
        roll = std::atan2(m21 / c, m22 / c);
        yaw = std::atan2(m10 / c, m00 / c);
    }
}

void rpy_to_quat(double roll, double pitch, double yaw, double q[4]) {
    const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
    const double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
    q[0] = sr * cp * cy - cr * sp * sy;
    q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy;
    q[3] = cr * cp * cy + sr * sp * sy;
}

void imu_deskew_info(const ImuQueueView& q, DeskewTableSet& t, size_t* drop_front) {
    t.imu_available = false;
    if (t.imu_time.size() != static_cast<size_t>(kImuQueueLength)) {
        t.imu_time.assign(kImuQueueLength, 0.0); t.imu_rot_x.assign(kImuQueueLength, 0.0);
        t.imu_rot_y.assign(kImuQueueLength, 0.0); t.imu_rot_z.assign(kImuQueueLength, 0.0);
    }
    size_t first = 0;
    while (first < q.n && q.stamp[first] < t.time_scan_cur - 0.01) ++first;  // :536-542
    if (drop_front) *drop_front = first;
    if (first >= q.n) return;                                                // :544-547
    t.imu_pointer_cur = 0;
    for (size_t i = first; i < q.n; ++i) {
        const double cur = q.stamp[i];
        if (cur > t.time_scan_end + 0.01) break;  // :556
        if (t.imu_pointer_cur == 0) {             // :558-565
            t.imu_rot_x[0] = t.imu_rot_y[0] = t.imu_rot_z[0] = 0.0;
            t.imu_time[0] = cur;
            ++t.imu_pointer_cur;
            continue;
        }
        const int k = t.imu_pointer_cur;
        if (k >= kImuQueueLength) break;  // (the reference would write past its tables)
        const double dt = cur - t.imu_time[k - 1];  // :572-576
        t.imu_rot_x[k] = t.imu_rot_x[k - 1] + q.gyro[3 * i] * dt;
        t.imu_rot_y[k] = t.imu_rot_y[k - 1] + q.gyro[3 * i + 1] * dt;
        t.imu_rot_z[k] = t.imu_rot_z[k - 1] + q.gyro[3 * i + 2] * dt;
        t.imu_time[k] = cur;
        ++t.imu_pointer_cur;
    }
    --t.imu_pointer_cur;  // :580
    if (t.imu_pointer_cur <= 0) return;
    t.imu_available = true;
}

void odom_deskew_info(const OdomQueueView& q, DeskewTableSet& t, size_t* drop_front) {
    t.odom_available = false;
    size_t first = 0;
    while (first < q.n && q.stamp[first] < t.time_scan_cur - 0.1) ++first;  // :591-596
    if (drop_front) *drop_front = first;
    if (first >= q.n) return;                         // :598-602 odometry too old
    if (q.stamp[first] > t.time_scan_cur) return;     // :604-607 nothing synced with the scan start
    // start of the sweep: first message not older than the scan start (:610-627)
    const size_t is = first_not_older(q.stamp, first, q.n, t.time_scan_cur);
    double roll, pitch, yaw;
    quat_to_rpy(&q.quat_xyzw[4 * is], roll, pitch, yaw);
    const Rigid3f b = get_transformation(q.pos[3 * is], q.pos[3 * is + 1], q.pos[3 * is + 2], roll, pitch, yaw);
    // end of the sweep (:630-714)
    double end_stamp, ex, ey, ez, eq[4];
    const size_t last = q.n - 1;
    if (q.stamp[last] > t.time_scan_end) {
        const size_t ie = first_not_older(q.stamp, first, q.n, t.time_scan_end);
        end_stamp = q.stamp[ie];
        ex = q.pos[3 * ie]; ey = q.pos[3 * ie + 1]; ez = q.pos[3 * ie + 2];
        for (int k = 0; k < 4; ++k) eq[k] = q.quat_xyzw[4 * ie + k];
    } else {
        // no odometry beyond the scan end: the latest message integrated forward with its own twist (:650-709)
        const double dt = t.time_scan_end - q.stamp[last];
        double r, p, y;
        quat_to_rpy(&q.quat_xyzw[4 * last], r, p, y);
        // Rz(yaw) Ry(pitch) Rx(roll) * local velocity
        const double cy = std::cos(y), sy = std::sin(y), cp = std::cos(p), sp = std::sin(p), cr = std::cos(r), sr = std::sin(r);
        const double R[3][3] = {{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr},
                                {sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr},
                                {-sp, cp * sr, cp * cr}};
        const double* v = &q.lin_vel[3 * last];
        const double gv[3] = {R[0][0] * v[0] + R[0][1] * v[1] + R[0][2] * v[2], R[1][0] * v[0] + R[1][1] * v[1] + R[1][2] * v[2],
                              R[2][0] * v[0] + R[2][1] * v[1] + R[2][2] * v[2]};
        ex = q.pos[3 * last] + gv[0] * dt; ey = q.pos[3 * last + 1] + gv[1] * dt; ez = q.pos[3 * last + 2] + gv[2] * dt;
        r += q.ang_vel[3 * last] * dt; p += q.ang_vel[3 * last + 1] * dt; y += q.ang_vel[3 * last + 2] * dt;
        rpy_to_quat(r, p, y, eq);  // setRPY, read back through getRPY below exactly as the node does
        end_stamp = t.time_scan_end;
    }
    quat_to_rpy(eq, roll, pitch, yaw);
    const Rigid3f e = get_transformation(ex, ey, ez, roll, pitch, yaw);
    // affine_trans_begin.inverse() * affine_trans_end (:716): only the translation is used (:724-726) = R_b^-1 (t_e - t_b), the
    // linear part inverted generally (cofactors, float) as Eigen's Affine inverse does
    const float (*M)[3] = b.r;
    const float c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1], c01 = M[0][2] * M[2][1] - M[0][1] * M[2][2], c02 = M[0][1] * M[1][2] - M[0][2] * M[1][1];
    const float c10 = M[1][2] * M[2][0] - M[1][0] * M[2][2], c11 = M[0][0] * M[2][2] - M[0][2] * M[2][0], c12 = M[0][2] * M[1][0] - M[0][0] * M[1][2];
    const float c20 = M[1][0] * M[2][1] - M[1][1] * M[2][0], c21 = M[0][1] * M[2][0] - M[0][0] * M[2][1], c22 = M[0][0] * M[1][1] - M[0][1] * M[1][0];
    const float det = M[0][0] * c00 + M[0][1] * c10 + M[0][2] * c20;
    const float id = 1.0f / det;
    const float dx = e.t[0] - b.t[0], dy = e.t[1] - b.t[1], dz = e.t[2] - b.t[2];
    const float tx = (c00 * dx + c01 * dy + c02 * dz) * id, ty = (c10 * dx + c11 * dy + c12 * dz) * id, tz = (c20 * dx + c21 * dy + c22 * dz) * id;
    // InterpolateTfWithTime (localization_functions.hpp:216-241): translation * ratio, the ratio narrowed to the vector's scalar
    const double dt_scan = t.time_scan_end - t.time_scan_cur;  // :719
    const double dt_trans = end_stamp - q.stamp[is];           // :720
    if (dt_trans == 0.0) { t.odom_incre[0] = t.odom_incre[1] = t.odom_incre[2] = 0.f; }
    else {
        const float ratio = static_cast<float>(dt_scan / dt_trans);
        t.odom_incre[0] = tx * ratio; t.odom_incre[1] = ty * ratio; t.odom_incre[2] = tz * ratio;
    }
    t.odom_available = true;
}

}  // namespace elm
