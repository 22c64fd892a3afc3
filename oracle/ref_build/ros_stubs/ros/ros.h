// Stand-in for the ROS headers the EKF sources include (ROS is absent from this image).  ekf_algorithm.{hpp,cpp} and
// localization_functions.hpp only need: the namespaces `ros` and `tf` to exist (ekf_algorithm.hpp:76-77 has using-directives
// for them), and a sensor_msgs::Imu with the fields the inline IMU converters of localization_functions.hpp:112-217 read.
#pragma once
#include <cstring>
#include <iomanip>

namespace ros {
struct Time {
    double sec = 0.0;
    double toSec() const { return sec; }
};
}  // namespace ros
namespace tf {}
namespace std_msgs {
struct Header { ros::Time stamp; };
}  // namespace std_msgs
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
}  // namespace geometry_msgs
namespace sensor_msgs {
struct Imu {
    std_msgs::Header header;
    geometry_msgs::Quaternion orientation;
    geometry_msgs::Vector3 angular_velocity, linear_acceleration;
};
}  // namespace sensor_msgs
