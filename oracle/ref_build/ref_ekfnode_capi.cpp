// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// C entry points over the REFERENCE's own EKF ROS node class EkfLocalization (ekf_localization/src/ekf_localization.cpp with
// ekf_algorithm.cpp and the vendored ini parser, compiled unmodified from /root/reference against node_stubs/ and stubs/).
// Like ref_node_capi.cpp the driver plays the middleware: messages in through the callbacks, publications out of the capture
// registry.  Together with libref_node.so this closes the localisation loop on the reference's own two nodes.
// Built by oracle/Makefile into oracle/_ref/libref_ekfnode.so (git-ignored).
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "node_world.hpp"
#include "ini_parser.h"
#include "localization_functions.hpp"
#include "ekf_localization_config.hpp"

#define private public
#include "ekf_algorithm.hpp"
#include "ekf_localization.hpp"
#undef private

namespace {
struct Quiet {
    Quiet() { std::cout.setstate(std::ios_base::failbit); }
    ~Quiet() { std::cout.clear(); }
};
nav_msgs::Odometry::ConstPtr make_odom(double t, const double* pos, const double* q_xyzw, const double* cov36) {
    auto m = std::make_shared<nav_msgs::Odometry>();
    m->header.stamp = ros::Time(t);
    m->header.frame_id = "world";
    m->pose.pose.position.x = pos[0]; m->pose.pose.position.y = pos[1]; m->pose.pose.position.z = pos[2];
    m->pose.pose.orientation.x = q_xyzw[0]; m->pose.pose.orientation.y = q_xyzw[1];
    m->pose.pose.orientation.z = q_xyzw[2]; m->pose.pose.orientation.w = q_xyzw[3];
    if (cov36) for (int i = 0; i < 36; ++i) m->pose.covariance[i] = cov36[i];
    return m;
}
}  // namespace

extern "C" {

void* ref_ekfnode_create(const char* config_dir) {
    Quiet q;
    setenv("PWD", config_dir, 1);
    ros::Capture::get().by_topic.clear();
    return new EkfLocalization("ekf_localization", 0.01);
}
void ref_ekfnode_destroy(void* h) { Quiet q; delete static_cast<EkfLocalization*>(h); }

void ref_ekfnode_imu(void* h, double t, const double* gyro, const double* acc) {
    Quiet q;
    auto m = std::make_shared<sensor_msgs::Imu>();
    m->header.stamp = ros::Time(t);
    m->angular_velocity.x = gyro[0]; m->angular_velocity.y = gyro[1]; m->angular_velocity.z = gyro[2];
    m->linear_acceleration.x = acc[0]; m->linear_acceleration.y = acc[1]; m->linear_acceleration.z = acc[2];
    ros::Capture::get().by_topic.clear();  // keep only the publications of this call
    static_cast<EkfLocalization*>(h)->CallbackImu(m);
}
void ref_ekfnode_pcm_odom(void* h, double t, const double* pos, const double* q_xyzw, const double* cov36) {
    Quiet q;
    static_cast<EkfLocalization*>(h)->CallbackPcmOdom(make_odom(t, pos, q_xyzw, cov36));
}
void ref_ekfnode_pcm_init_odom(void* h, double t, const double* pos, const double* q_xyzw) {
    Quiet q;
    static_cast<EkfLocalization*>(h)->CallbackPcmInitOdom(make_odom(t, pos, q_xyzw, nullptr));
}
// /app/loc/ekf_pose_odom of the last IMU callback: stamp, position, orientation (x, y, z, w), local linear velocity, angular
// velocity; 0 if the callback published none
int ref_ekfnode_last_odom(double* stamp, double* pos, double* q_xyzw, double* lin, double* ang) {
    auto& v = ros::Capture::get().by_topic["/app/loc/ekf_pose_odom"];
    if (v.empty()) return 0;
    const auto* m = static_cast<const nav_msgs::Odometry*>(v.back().get());
    *stamp = m->header.stamp.toSec();
    pos[0] = m->pose.pose.position.x; pos[1] = m->pose.pose.position.y; pos[2] = m->pose.pose.position.z;
    q_xyzw[0] = m->pose.pose.orientation.x; q_xyzw[1] = m->pose.pose.orientation.y;
    q_xyzw[2] = m->pose.pose.orientation.z; q_xyzw[3] = m->pose.pose.orientation.w;
    lin[0] = m->twist.twist.linear.x; lin[1] = m->twist.twist.linear.y; lin[2] = m->twist.twist.linear.z;
    ang[0] = m->twist.twist.angular.x; ang[1] = m->twist.twist.angular.y; ang[2] = m->twist.twist.angular.z;
    return 1;
}
// the filter's raw pose: position + quaternion (w, x, y, z)
void ref_ekfnode_filter_pose(void* h, double* pos, double* q_wxyz) {
    const auto& S = static_cast<EkfLocalization*>(h)->ptr_ekf_algorithm_->S_;
    pos[0] = S.pos.x(); pos[1] = S.pos.y(); pos[2] = S.pos.z();
    q_wxyz[0] = S.rot.w(); q_wxyz[1] = S.rot.x(); q_wxyz[2] = S.rot.y(); q_wxyz[3] = S.rot.z();
}
// GnssTimeCompensation (ekf_localization.cpp:323-394) on its own against the node's current state queue.
// in / out: t, pos[3], quat (w, x, y, z); returns 0 when the node refuses to compensate
int ref_ekfnode_time_compensate(void* h, double t, const double* pos, const double* q_wxyz, double* t_out, double* pos_out, double* q_out) {
    Quiet q;
    EkfGnssMeasurement in, out;
    in.timestamp = t;
    in.gnss_source = GnssSource::PCM;
    in.pos = Eigen::Vector3d(pos[0], pos[1], pos[2]);
    in.rot = Eigen::Quaterniond(q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]);
    if (!static_cast<EkfLocalization*>(h)->GnssTimeCompensation(in, out)) return 0;
    *t_out = out.timestamp;
    pos_out[0] = out.pos.x(); pos_out[1] = out.pos.y(); pos_out[2] = out.pos.z();
    q_out[0] = out.rot.w(); q_out[1] = out.rot.x(); q_out[2] = out.rot.y(); q_out[3] = out.rot.z();
    return 1;
}
// the node's EKF state queue (what GnssTimeCompensation interpolates over): n rows of {t, x, y, z, roll, pitch, yaw}
size_t ref_ekfnode_state_queue(void* h, double* rows, size_t capacity) {
    auto& d = static_cast<EkfLocalization*>(h)->deq_ekf_state_;
    size_t n = 0;
    for (const auto& e : d) {
        if (n >= capacity) break;
        const double r[7] = {e.timestamp, e.x_m, e.y_m, e.z_m, e.roll_rad, e.pitch_rad, e.yaw_rad};
        std::memcpy(rows + 7 * n, r, sizeof r);
        ++n;
    }
    return d.size();
}

}  // extern "C"
