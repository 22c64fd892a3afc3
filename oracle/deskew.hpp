// ORACLE — TEST INFRASTRUCTURE ONLY (see smallmat.hpp header).
// Parity status: PINNED on oracle/_ref/libref_node.so — the reference's own ROS node class (pcm_matching.cpp) compiled
// unmodified against stand-in ROS / tf / PCL / Eigen headers (tests/test_reference_build_node.py: tables 1e-15, the per-point
// transform bit-equal); pcl::getTransformation and tf's RPY conversions stay restated from their published definitions.
// CPU restatement of the per-point deskew of pcm_matching (float32 arithmetic like the reference):
//   /root/reference/src/app/localization/pcm_matching/src/pcm_matching.cpp
//     DeskewPointCloud :467-531   ImuDeskewInfo :533-585   OdomDeskewInfo :587-729
//     FindRotation :731-762       FindPosition :764-778    DeskewPoint :780-824
// Third-party arithmetic absent from /root/reference: pcl::getTransformation (PCL, unpinned; noetic => 1.10,
// common/impl/eigen.hpp) — restated below from its published closed form  R = Rz(yaw) Ry(pitch) Rx(roll).
#pragma once
#include <vector>

namespace orc {

constexpr int kImuQueueLength = 2000;  // pcm_matching.hpp:113

struct DeskewTables {
    // ImuDeskewInfo: integrated gyro angles at the IMU stamps of this scan
    std::vector<double> imu_time, imu_rot_x, imu_rot_y, imu_rot_z;
    int imu_pointer_cur = 0;
    bool imu_available = false;
    // OdomDeskewInfo: translation of the scan-span odometry increment (only xyz is used, pcm_matching.cpp:724-726)
    bool odom_available = false;
    float odom_incre_x = 0.f, odom_incre_y = 0.f, odom_incre_z = 0.f;
    double time_scan_cur = 0.0, time_scan_end = 0.0;
};

// pcm_matching.cpp:533-585 on plain arrays (stamps ascending, gyro already in the ego frame)
void ImuDeskewInfo(const double* stamp, const double* gyro_xyz, int n, DeskewTables& t);
// pcm_matching.cpp:587-729 reduced to its arithmetic: start / end odometry poses (x y z roll pitch yaw) and stamps
void OdomDeskewInfo(const double start_pose[6], double start_stamp, const double end_pose[6], double end_stamp, DeskewTables& t);
// pcm_matching.cpp:780-824 (+ :731-778)
void DeskewPoint(const DeskewTables& t, const float in[3], double rel_time, float out[3]);
void DeskewPoints(const DeskewTables& t, const float* xyz, const float* rel_time, size_t n, float* out);

}  // namespace orc
