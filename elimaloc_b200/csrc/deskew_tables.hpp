// Host-side table builders of the deskew stage: PcmMatching::ImuDeskewInfo / OdomDeskewInfo
// (pcm_matching/src/pcm_matching.cpp:533-729) and the time base of DeskewPointCloud (:467-489), on plain arrays instead
// of ROS message queues.  Host arithmetic on a few hundred samples per scan; the per-point work is deskew.cu.
#pragma once
#include <cstddef>
#include <vector>

namespace elm {

constexpr int kImuQueueLength = 2000;  // pcm_matching.hpp:113 (capacity of the four member tables)

struct ImuQueueView {   // deq_imu_: stamp[n], gyro[3 n] = angular velocity as ImuAngular2RosAngular returns it
    const double* stamp;
    const double* gyro;
    size_t n;
};
struct OdomQueueView {  // deq_odom_: stamp[n], position[3 n], orientation (x, y, z, w)[4 n], twist linear[3 n] (body frame), twist angular[3 n]
    const double* stamp;
    const double* pos;
    const double* quat_xyzw;
    const double* lin_vel;
    const double* ang_vel;
    size_t n;
};

struct DeskewTableSet {
    std::vector<double> imu_time, imu_rot_x, imu_rot_y, imu_rot_z;  // kImuQueueLength entries each
    int imu_pointer_cur = 0;
    bool imu_available = false, odom_available = false;
    float odom_incre[3] = {0.f, 0.f, 0.f};
    double time_scan_cur = 0.0, time_scan_end = 0.0;
};

// ImuDeskewInfo (:533-585).  drop_front (may be NULL): samples the node pops from the front of its queue (older than scan start - 0.01 s).
void imu_deskew_info(const ImuQueueView& q, DeskewTableSet& t, size_t* drop_front);
// OdomDeskewInfo (:587-729) incl. the integrate-forward branch when no odometry lies beyond the scan end.
void odom_deskew_info(const OdomQueueView& q, DeskewTableSet& t, size_t* drop_front);

// tf::Matrix3x3(q).getRPY / tf::Quaternion::setRPY (Bullet's published formulas; q = x, y, z, w)
void quat_to_rpy(const double q_xyzw[4], double& roll, double& pitch, double& yaw);
void rpy_to_quat(double roll, double pitch, double yaw, double q_xyzw[4]);

}  // namespace elm
