// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// C entry points over the REFERENCE's own ROS node class PcmMatching (pcm_matching/src/pcm_matching.cpp + include/pcm_matching.hpp,
// with registration.cpp, voxel_hash_map.cpp and bsw/system/ini_parser/ini_parser.cpp, all compiled unmodified from
// /root/reference against the stand-in headers of node_stubs/ and stubs/).  The driver plays the middleware: it hands messages
// to the node's callbacks and reads back what the node "published" and the members of its deskew stage.
// Built by oracle/Makefile into oracle/_ref/libref_node.so (git-ignored).
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "node_world.hpp"
#include "ini_parser.h"
#include "localization_functions.hpp"
#include "pcm_matching_config.hpp"
#include "registration.hpp"

// The members of the deskew stage are private; the dumps below only READ them (and call the private stage functions the way
// CallbackPointCloud does).  Everything the class header includes is already included above.
#define private public
#include "pcm_matching.hpp"
#undef private

namespace {
struct Quiet {
    Quiet() { std::cout.setstate(std::ios_base::failbit); }
    ~Quiet() { std::cout.clear(); }
};
template <typename M>
const M* last_on(const std::string& topic) {
    auto& v = ros::Capture::get().by_topic[topic];
    return v.empty() ? nullptr : static_cast<const M*>(v.back().get());
}
pcl::PointCloud<PointXYZIT>::Ptr make_cloud(const float* xyz, const float* rel_time, size_t n) {
    pcl::PointCloud<PointXYZIT>::Ptr c(new pcl::PointCloud<PointXYZIT>());
    c->points.resize(n);
    for (size_t i = 0; i < n; ++i) {
        PointXYZIT p{};
        p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
        p.intensity = static_cast<float>(i);  // carries the input index through the filter
        p.time = rel_time ? rel_time[i] : 0.f;
        c->points[i] = p;
    }
    c->width = static_cast<uint32_t>(n);
    c->height = 1;
    return c;
}
}  // namespace

extern "C" {

// what Init() published about the map: the voxel-map cloud (visualisation) and, for VGICP / AVGICP, one marker per voxel
// covariance with more than 2 points (position = voxel mean, scale = 3 sqrt(eigenvalue + 0.01) by descending eigenvalue)
size_t ref_node_init_map_cloud(float* xyz, size_t capacity) {
    const auto* m = last_on<sensor_msgs::PointCloud2>("/app/loc/voxel_map_pc");
    if (!m) return 0;
    for (size_t i = 0; i < m->records.size() && i < capacity; ++i) { xyz[3 * i] = m->records[i].x; xyz[3 * i + 1] = m->records[i].y; xyz[3 * i + 2] = m->records[i].z; }
    return m->records.size();
}
size_t ref_node_init_markers(double* pos_scale6, size_t capacity) {
    const auto* m = last_on<visualization_msgs::MarkerArray>("/app/loc/voxel_map_cov");
    if (!m) return 0;
    for (size_t i = 0; i < m->markers.size() && i < capacity; ++i) {
        const auto& k = m->markers[i];
        const double v[6] = {k.pose.position.x, k.pose.position.y, k.pose.position.z, k.scale.x, k.scale.y, k.scale.z};
        std::memcpy(pos_scale6 + 6 * i, v, sizeof v);
    }
    return m->markers.size();
}

// CallbackInitialPose (pcm_matching.cpp:356-447): an rviz pose (x, y, yaw) -> ground height from the map -> RunRegister on
// the last lidar cloud -> /app/loc/pcm_init_odom.  Returns 1 and the published pose when the node published one.
int ref_node_initial_pose(void* h, double x, double y, double z, double yaw, double* pos, double* quat_xyzw) {
    Quiet q;
    auto m = std::make_shared<geometry_msgs::PoseWithCovarianceStamped>();
    m->pose.pose.position.x = x; m->pose.pose.position.y = y; m->pose.pose.position.z = z;
    m->pose.pose.orientation.w = std::cos(0.5 * yaw); m->pose.pose.orientation.z = std::sin(0.5 * yaw);
    m->pose.pose.orientation.x = 0.0; m->pose.pose.orientation.y = 0.0;
    const size_t before = ros::Capture::get().by_topic["/app/loc/pcm_init_odom"].size();
    static_cast<PcmMatching*>(h)->CallbackInitialPose(m);
    if (ros::Capture::get().by_topic["/app/loc/pcm_init_odom"].size() == before) return 0;
    const auto* o = last_on<nav_msgs::Odometry>("/app/loc/pcm_init_odom");
    pos[0] = o->pose.pose.position.x; pos[1] = o->pose.pose.position.y; pos[2] = o->pose.pose.position.z;
    quat_xyzw[0] = o->pose.pose.orientation.x; quat_xyzw[1] = o->pose.pose.orientation.y;
    quat_xyzw[2] = o->pose.pose.orientation.z; quat_xyzw[3] = o->pose.pose.orientation.w;
    return 1;
}

// config_dir must hold config/localization.ini and config/calibration.ini (the node reads $PWD/config/...); the map is what
// the node would have loaded from its .pcd file.
void* ref_node_create(const char* config_dir, const float* map_xyz, size_t n_map) {
    Quiet q;
    setenv("PWD", config_dir, 1);
    ros::Capture::get().by_topic.clear();
    ros::Capture::get().params["/pcm_matching/map_path"] = "map.pcd";
    pcl::io::MapRegistry::get().xyz.assign(map_xyz, map_xyz + 3 * n_map);
    auto* node = new PcmMatching("pcm_matching", 1.0);
    pcl::io::MapRegistry::get().xyz.clear();
    pcl::io::MapRegistry::get().xyz.shrink_to_fit();
    return node;
}
void ref_node_destroy(void* h) { delete static_cast<PcmMatching*>(h); }
size_t ref_node_map_points(void* h) { return static_cast<PcmMatching*>(h)->local_map_.Pointcloud().size(); }

void ref_node_imu(void* h, double t, const double* gyro, const double* acc) {
    auto m = std::make_shared<sensor_msgs::Imu>();
    m->header.stamp = ros::Time(t);
    m->angular_velocity.x = gyro[0]; m->angular_velocity.y = gyro[1]; m->angular_velocity.z = gyro[2];
    m->linear_acceleration.x = acc[0]; m->linear_acceleration.y = acc[1]; m->linear_acceleration.z = acc[2];
    static_cast<PcmMatching*>(h)->CallbackImu(m);
}
// EKF odometry: position, orientation (x, y, z, w), local linear velocity, angular velocity
void ref_node_odom(void* h, double t, const double* pos, const double* quat_xyzw, const double* lin, const double* ang) {
    Quiet q;
    auto m = std::make_shared<nav_msgs::Odometry>();
    m->header.stamp = ros::Time(t);
    m->header.frame_id = "world";
    m->pose.pose.position.x = pos[0]; m->pose.pose.position.y = pos[1]; m->pose.pose.position.z = pos[2];
    m->pose.pose.orientation.x = quat_xyzw[0]; m->pose.pose.orientation.y = quat_xyzw[1];
    m->pose.pose.orientation.z = quat_xyzw[2]; m->pose.pose.orientation.w = quat_xyzw[3];
    m->twist.twist.linear.x = lin[0]; m->twist.twist.linear.y = lin[1]; m->twist.twist.linear.z = lin[2];
    m->twist.twist.angular.x = ang[0]; m->twist.twist.angular.y = ang[1]; m->twist.twist.angular.z = ang[2];
    static_cast<PcmMatching*>(h)->CallbackEkfState(m);
}
// One lidar message through CallbackPointCloud; returns how many /app/loc/pcm_odom messages exist afterwards.
size_t ref_node_cloud(void* h, double stamp, const float* xyz, const float* rel_time, size_t n) {
    Quiet q;
    auto m = std::make_shared<sensor_msgs::PointCloud2>();
    m->header.stamp = ros::Time(stamp);
    m->header.frame_id = "lidar";
    m->records.resize(n);
    for (size_t i = 0; i < n; ++i) {
        sensor_msgs::PointRecord r;
        r.x = xyz[3 * i]; r.y = xyz[3 * i + 1]; r.z = xyz[3 * i + 2];
        r.intensity = static_cast<float>(i);
        r.time = rel_time[i];
        m->records[i] = r;
    }
    static_cast<PcmMatching*>(h)->CallbackPointCloud(m);
    return ros::Capture::get().by_topic["/app/loc/pcm_odom"].size();
}
// last /app/loc/pcm_odom: stamp, position, orientation (x, y, z, w), 6 x 6 covariance; 0 if none
int ref_node_last_pcm_odom(double* stamp, double* pos, double* quat_xyzw, double* cov36) {
    const auto* m = last_on<nav_msgs::Odometry>("/app/loc/pcm_odom");
    if (!m) return 0;
    *stamp = m->header.stamp.toSec();
    pos[0] = m->pose.pose.position.x; pos[1] = m->pose.pose.position.y; pos[2] = m->pose.pose.position.z;
    quat_xyzw[0] = m->pose.pose.orientation.x; quat_xyzw[1] = m->pose.pose.orientation.y;
    quat_xyzw[2] = m->pose.pose.orientation.z; quat_xyzw[3] = m->pose.pose.orientation.w;
    for (int i = 0; i < 36; ++i) cov36[i] = m->pose.covariance[i];
    return 1;
}
// last /app/loc/icp_map_pc: the down-sampled scan in the world frame (what RunRegister aligned); returns its size
size_t ref_node_last_registered_cloud(float* xyz, size_t capacity) {
    const auto* m = last_on<sensor_msgs::PointCloud2>("/app/loc/icp_map_pc");
    if (!m) return 0;
    for (size_t i = 0; i < m->records.size() && i < capacity; ++i) { xyz[3 * i] = m->records[i].x; xyz[3 * i + 1] = m->records[i].y; xyz[3 * i + 2] = m->records[i].z; }
    return m->records.size();
}

// FilterPointsByDistance (pcm_matching.cpp:451-465) on its own: surviving input indices
size_t ref_node_filter_by_distance(void* h, const float* xyz, size_t n, int32_t* index_out) {
    auto c = make_cloud(xyz, nullptr, n);
    static_cast<PcmMatching*>(h)->FilterPointsByDistance(c);
    for (size_t i = 0; i < c->points.size(); ++i) index_out[i] = static_cast<int32_t>(c->points[i].intensity);
    return c->points.size();
}

// DeskewPointCloud (pcm_matching.cpp:467-531) on its own, against the node's current IMU / odometry queues.
// Returns 1 on success; the undistorted cloud and the tables of the stage are read with the two functions below.
int ref_node_deskew(void* h, double stamp, const float* xyz, const float* rel_time, size_t n) {
    Quiet q;
    auto c = make_cloud(xyz, rel_time, n);
    return static_cast<PcmMatching*>(h)->DeskewPointCloud(c, stamp) ? 1 : 0;
}
size_t ref_node_undistorted(void* h, float* xyz, size_t capacity) {
    const auto& pts = static_cast<PcmMatching*>(h)->undistort_pcptr_->points;
    for (size_t i = 0; i < pts.size() && i < capacity; ++i) { xyz[3 * i] = pts[i].x; xyz[3 * i + 1] = pts[i].y; xyz[3 * i + 2] = pts[i].z; }
    return pts.size();
}
// members of the deskew stage: tables[4][2000] (time, rot x/y/z), meta_i = {imu_pointer_cur, imu_available, odom_available},
// meta_f = odom increments, times = {scan_cur, scan_end}
void ref_node_deskew_tables(void* h, double* imu_time, double* rot_x, double* rot_y, double* rot_z, int32_t* meta_i, float* meta_f, double* times) {
    auto* n = static_cast<PcmMatching*>(h);
    std::memcpy(imu_time, n->vec_d_imu_time_, i_queue_length_ * sizeof(double));
    std::memcpy(rot_x, n->vec_d_imu_rot_x_, i_queue_length_ * sizeof(double));
    std::memcpy(rot_y, n->vec_d_imu_rot_y_, i_queue_length_ * sizeof(double));
    std::memcpy(rot_z, n->vec_d_imu_rot_z_, i_queue_length_ * sizeof(double));
    meta_i[0] = n->i_imu_pointer_cur_; meta_i[1] = n->b_is_imu_available_; meta_i[2] = n->b_is_odom_available_;
    meta_f[0] = n->f_odom_incre_x_; meta_f[1] = n->f_odom_incre_y_; meta_f[2] = n->f_odom_incre_z_;
    times[0] = n->d_time_scan_cur_; times[1] = n->d_time_scan_end_;
}

// GetInterpolatedPose (pcm_matching.cpp:933-1045): 4 x 4 row-major float pose; returns 0 when no pose precedes t
int ref_node_interpolated_pose(void* h, double t, float* T16) {
    Quiet q;
    Eigen::Affine3f a;
    if (!static_cast<PcmMatching*>(h)->GetInterpolatedPose(t, a)) return 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T16[4 * i + j] = a.matrix()(i, j);
    return 1;
}

// PublishPcmOdom's covariance shaping (pcm_matching.cpp:1047-1101) on its own
void ref_node_shape_covariance(void* h, const double* pose16, const double* local_cov36, double icp_pose_std_m, double* cov36) {
    auto* n = static_cast<PcmMatching*>(h);
    Eigen::Matrix4d T;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T(i, j) = pose16[4 * i + j];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) n->icp_local_cov_(i, j) = local_cov36[6 * i + j];
    n->cfg_.d_icp_pose_std_m = icp_pose_std_m;
    n->PublishPcmOdom(T, ros::Time(0.0), "world");
    const auto* m = last_on<nav_msgs::Odometry>("/app/loc/pcm_odom");
    for (int i = 0; i < 36; ++i) cov36[i] = m->pose.covariance[i];
    ros::Capture::get().by_topic["/app/loc/pcm_odom"].pop_back();
}

}  // extern "C"
