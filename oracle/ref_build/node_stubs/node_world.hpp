// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// Stand-in for everything third-party that the reference's ROS node pcm_matching.{hpp,cpp} includes — ROS (roscpp, tf, tf2_ros,
// the message packages), PCL, boost, OpenCV — none of which is in this image.  With it the node's own translation unit compiles
// UNMODIFIED from /root/reference into oracle/_ref/libref_node.so, so that its callbacks (CallbackImu, CallbackEkfState,
// CallbackPointCloud: distance filter -> deskew -> pose interpolation -> voxel down-sampling -> RunRegister -> covariance
// shaping -> publish) run as written.  The test driver plays the middleware: it calls the callbacks directly and reads what the
// node "publishes" from the capture registry below.
//
// What is restated here, from the published definitions, because the node calls it:
//   tf::Matrix3x3::getRPY / tf::Quaternion::setRPY (Bullet's formulas), pcl::getTransformation / getTranslationAndEulerAngles
//   (pcl/common/eigen.h), pcl::transformPointCloud, the float Affine3f algebra (in stubs/mini_eigen.hpp).
// Every forwarding header of this directory includes only this file.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "Eigen/Dense"

// ---------------------------------------------------------------------------------------------------------------- boost
namespace boost {
template <typename T, std::size_t N> using array = std::array<T, N>;
template <typename T> using shared_ptr = std::shared_ptr<T>;
struct thread {
    template <typename F> explicit thread(F) {}  // the node's MainLoop thread is never started by the driver
    void join() {}
};
template <typename F, typename O> std::function<void()> bind(F f, O o) { return [f, o] { (o->*f)(); }; }
}  // namespace boost

// ------------------------------------------------------------------------------------------------------------------ ros
namespace ros {
struct Duration {
    double sec = 0.0;
    Duration() {}
    explicit Duration(double s) : sec(s) {}
    double toSec() const { return sec; }
};
struct Time {  // real ROS stores integer sec / nsec, i.e. quantises a double stamp to 1 ns; the stand-in keeps the double
    double sec = 0.0;
    Time() {}
    explicit Time(double s) : sec(s) {}
    double toSec() const { return sec; }
    static Time now() { return Time(0.0); }
    Time& operator-=(const Duration& d) { sec -= d.sec; return *this; }
    Duration operator-(const Time& o) const { return Duration(sec - o.sec); }
};
struct Rate { explicit Rate(double) {} void sleep() {} };
struct AsyncSpinner { explicit AsyncSpinner(int) {} void start() {} };
inline void init(int&, char**, const std::string&) {}
inline bool ok() { return false; }
inline void shutdown() {}
inline void waitForShutdown() {}
namespace package { inline std::string getPath(const std::string&) { return "."; } }

// what the node publishes, by topic, kept for the driver
struct Capture {
    std::map<std::string, std::vector<std::shared_ptr<void>>> by_topic;
    std::map<std::string, std::string> params;
    static Capture& get() { static Capture c; return c; }
};
struct Publisher {
    std::string topic;
    template <typename M> void publish(const M& m) const { Capture::get().by_topic[topic].push_back(std::make_shared<M>(m)); }
};
struct Subscriber {};
struct NodeHandle {
    template <typename M, typename T> Subscriber subscribe(const std::string&, int, void (T::*)(const M&), T*) { return Subscriber(); }
    template <typename M> Publisher advertise(const std::string& topic, int) { return Publisher{topic}; }
    bool getParam(const std::string& key, std::string& out) const {
        auto it = Capture::get().params.find(key);
        if (it == Capture::get().params.end()) return false;
        out = it->second;
        return true;
    }
    bool getParam(const std::string& key, double& out) const {
        auto it = Capture::get().params.find(key);
        if (it == Capture::get().params.end()) return false;
        out = std::atof(it->second.c_str());
        return true;
    }
};
}  // namespace ros
#define ROS_WARN_STREAM(x) do { std::ostringstream ros_stub_oss; ros_stub_oss << x; } while (0)
#define ROS_INFO_STREAM(x) ROS_WARN_STREAM(x)
#define ROS_ERROR_STREAM(x) ROS_WARN_STREAM(x)
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { } while (0)

// ------------------------------------------------------------------------------------------------------------- messages
namespace std_msgs {
struct Float32 { float data = 0; };
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; };
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
}  // namespace std_msgs
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct Twist { Vector3 linear, angular; };
struct PoseWithCovariance { Pose pose; boost::array<double, 36> covariance{}; };
struct TwistWithCovariance { Twist twist; boost::array<double, 36> covariance{}; };
struct PoseWithCovarianceStamped {
    std_msgs::Header header;
    PoseWithCovariance pose;
    using ConstPtr = boost::shared_ptr<const PoseWithCovarianceStamped>;
};
struct TwistStamped { std_msgs::Header header; Twist twist; };
using TwistStampedConstPtr = boost::shared_ptr<const TwistStamped>;
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::Header header; std::string child_frame_id; Transform transform; };
}  // namespace geometry_msgs
namespace nav_msgs {
struct Odometry {
    std_msgs::Header header;
    std::string child_frame_id;
    geometry_msgs::PoseWithCovariance pose;
    geometry_msgs::TwistWithCovariance twist;
    using ConstPtr = boost::shared_ptr<const Odometry>;
};
}  // namespace nav_msgs
namespace sensor_msgs {
struct Imu {
    std_msgs::Header header;
    geometry_msgs::Quaternion orientation;
    geometry_msgs::Vector3 angular_velocity, linear_acceleration;
    using ConstPtr = boost::shared_ptr<const Imu>;
};
struct NavSatStatus { int8_t status = 0; uint16_t service = 0; };
struct NavSatFix {
    std_msgs::Header header;
    NavSatStatus status;
    double latitude = 0, longitude = 0, altitude = 0;
    boost::array<double, 9> position_covariance{};
    uint8_t position_covariance_type = 0;
    using ConstPtr = boost::shared_ptr<const NavSatFix>;
};
// The wire format is not modelled: a cloud message carries typed records (x, y, z, intensity, per-point time and the Ouster
// fields); pcl::fromROSMsg below copies the fields the target point type has.
struct PointRecord { float x = 0, y = 0, z = 0, intensity = 0, time = 0; uint32_t t = 0; uint16_t reflectivity = 0, ring = 0, ambient = 0; uint32_t range = 0; };
struct PointCloud2 {
    std_msgs::Header header;
    std::vector<PointRecord> records;
    bool is_dense = true;
    using ConstPtr = boost::shared_ptr<const PointCloud2>;
};
}  // namespace sensor_msgs
namespace visualization_msgs {
struct Marker {
    enum { CUBE = 1, CYLINDER = 3, ADD = 0 };
    std_msgs::Header header;
    std::string ns;
    int id = 0, type = 0, action = 0;
    geometry_msgs::Pose pose;
    geometry_msgs::Vector3 scale;
    std_msgs::ColorRGBA color;
};
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs

// ------------------------------------------------------------------------------------------------------------------- tf
namespace tf {
struct Quaternion {  // Bullet's tf::Quaternion: (x, y, z, w)
    double x_ = 0, y_ = 0, z_ = 0, w_ = 1;
    Quaternion() {}
    Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
    double x() const { return x_; }
    double y() const { return y_; }
    double z() const { return z_; }
    double w() const { return w_; }
    void setRPY(double roll, double pitch, double yaw) {  // tf/LinearMath/Quaternion.h setRPY
        const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
        const double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
        x_ = sr * cp * cy - cr * sp * sy;
        y_ = cr * sp * cy + sr * cp * sy;
        z_ = cr * cp * sy - sr * sp * cy;
        w_ = cr * cp * cy + sr * sp * sy;
    }
};
inline void quaternionMsgToTF(const geometry_msgs::Quaternion& m, Quaternion& q) { q = Quaternion(m.x, m.y, m.z, m.w); }
struct Matrix3x3 {  // tf/LinearMath/Matrix3x3.h: setRotation + getEulerYPR (solution 1)
    double m[3][3];
    explicit Matrix3x3(const Quaternion& q) {
        const double d = q.x() * q.x() + q.y() * q.y() + q.z() * q.z() + q.w() * q.w();
        const double s = 2.0 / d;
        const double xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
        const double wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
        const double xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs;
        const double yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
        m[0][0] = 1.0 - (yy + zz); m[0][1] = xy - wz; m[0][2] = xz + wy;
        m[1][0] = xy + wz; m[1][1] = 1.0 - (xx + zz); m[1][2] = yz - wx;
        m[2][0] = xz - wy; m[2][1] = yz + wx; m[2][2] = 1.0 - (xx + yy);
    }
    void getRPY(double& roll, double& pitch, double& yaw) const {
        if (std::fabs(m[2][0]) >= 1.0) {  // gimbal lock branch of getEulerYPR
            yaw = 0.0;
            const double delta = std::atan2(m[2][1], m[2][2]);
            if (m[2][0] < 0) { pitch = M_PI / 2.0; roll = delta; }
            else { pitch = -M_PI / 2.0; roll = delta; }
        } else {
            pitch = -std::asin(m[2][0]);
            roll = std::atan2(m[2][1] / std::cos(pitch), m[2][2] / std::cos(pitch));
            yaw = std::atan2(m[1][0] / std::cos(pitch), m[0][0] / std::cos(pitch));
        }
    }
};
struct Vector3 { double x, y, z; Vector3(double a = 0, double b = 0, double c = 0) : x(a), y(b), z(c) {} };
struct Transform {
    Vector3 origin;
    Quaternion rotation;
    void setOrigin(const Vector3& v) { origin = v; }
    void setRotation(const Quaternion& q) { rotation = q; }
};
struct StampedTransform {
    Transform transform;
    ros::Time stamp;
    std::string frame_id, child_frame_id;
    StampedTransform(const Transform& t, const ros::Time& s, const std::string& f, const std::string& c) : transform(t), stamp(s), frame_id(f), child_frame_id(c) {}
};
struct TransformBroadcaster { void sendTransform(const StampedTransform&) {} };
}  // namespace tf
namespace jsk_rviz_plugins {
struct OverlayText {
    int action = 0, width = 0, height = 0, left = 0, top = 0, text_size = 0, line_width = 0;
    std_msgs::ColorRGBA bg_color, fg_color;
    std::string font, text;
};
}  // namespace jsk_rviz_plugins
// GeographicLib::LocalCartesian is only used to fill the latitude / longitude / height fields of the published state
// (ekf_localization.cpp:412-418, 643-648); geodesy is not restated: a flat-earth stand-in keeps the calls well defined.
namespace GeographicLib {
class LocalCartesian {
public:
    LocalCartesian(double lat0, double lon0, double h0) : lat0_(lat0), lon0_(lon0), h0_(h0) {}
    void Reverse(double x, double y, double z, double& lat, double& lon, double& h) const {
        lat = lat0_ + y / 111320.0; lon = lon0_ + x / (111320.0 * std::cos(lat0_ * M_PI / 180.0)); h = h0_ + z;
    }
    void Forward(double lat, double lon, double h, double& x, double& y, double& z) const {
        y = (lat - lat0_) * 111320.0; x = (lon - lon0_) * 111320.0 * std::cos(lat0_ * M_PI / 180.0); z = h - h0_;
    }
private:
    double lat0_, lon0_, h0_;
};
}  // namespace GeographicLib
namespace tf2_ros {
struct StaticTransformBroadcaster { void sendTransform(const geometry_msgs::TransformStamped&) {} };
}  // namespace tf2_ros

// ------------------------------------------------------------------------------------------------------------------ pcl
#define PCL_ADD_POINT4D float x, y, z, pcl_stub_pad
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fields)

namespace pcl {
struct PointXYZINormal {
    float x = 0, y = 0, z = 0, intensity = 0, normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0;
};
template <typename PointT>
struct PointCloud {
    using Ptr = boost::shared_ptr<PointCloud<PointT>>;
    using ConstPtr = boost::shared_ptr<const PointCloud<PointT>>;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = height = 0; }
    PointT& operator[](std::size_t i) { return points[i]; }
    const PointT& operator[](std::size_t i) const { return points[i]; }
    void push_back(const PointT& p) { points.push_back(p); width = static_cast<uint32_t>(points.size()); height = 1; }
};

namespace stub {
// field-wise copy between point types by the names they share with the message record / with each other
template <typename P> auto set_intensity(P& p, float v, int) -> decltype(p.intensity = v, void()) { p.intensity = v; }
template <typename P> void set_intensity(P&, float, long) {}
template <typename P> auto set_time(P& p, float v, int) -> decltype(p.time = v, void()) { p.time = v; }
template <typename P> void set_time(P&, float, long) {}
template <typename P> auto set_ouster(P& p, const sensor_msgs::PointRecord& r, int) -> decltype(p.reflectivity = r.reflectivity, void()) {
    p.t = r.t; p.reflectivity = r.reflectivity; p.ring = r.ring; p.ambient = r.ambient; p.range = r.range;
}
template <typename P> void set_ouster(P&, const sensor_msgs::PointRecord&, long) {}
template <typename P> auto get_intensity(const P& p, int) -> decltype(static_cast<float>(p.intensity)) { return p.intensity; }
template <typename P> float get_intensity(const P&, long) { return 0.f; }
}  // namespace stub

template <typename PointT>
void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointT>& cloud) {
    cloud.points.resize(msg.records.size());
    for (std::size_t i = 0; i < msg.records.size(); ++i) {
        PointT p{};
        p.x = msg.records[i].x; p.y = msg.records[i].y; p.z = msg.records[i].z;
        stub::set_intensity(p, msg.records[i].intensity, 0);
        stub::set_time(p, msg.records[i].time, 0);
        stub::set_ouster(p, msg.records[i], 0);
        cloud.points[i] = p;
    }
    cloud.width = static_cast<uint32_t>(cloud.points.size());
    cloud.height = 1;
    cloud.is_dense = msg.is_dense;
}
template <typename PointT> void moveFromROSMsg(sensor_msgs::PointCloud2& msg, PointCloud<PointT>& cloud) { fromROSMsg(msg, cloud); }
template <typename PointT>
void toROSMsg(const PointCloud<PointT>& cloud, sensor_msgs::PointCloud2& msg) {
    msg.records.resize(cloud.points.size());
    for (std::size_t i = 0; i < cloud.points.size(); ++i) {
        sensor_msgs::PointRecord r;
        r.x = cloud.points[i].x; r.y = cloud.points[i].y; r.z = cloud.points[i].z;
        r.intensity = stub::get_intensity(cloud.points[i], 0);
        msg.records[i] = r;
    }
    msg.is_dense = cloud.is_dense;
}
template <typename A, typename B>
void copyPointCloud(const PointCloud<A>& in, PointCloud<B>& out) {
    out.points.resize(in.points.size());
    for (std::size_t i = 0; i < in.points.size(); ++i) {
        B p{};
        p.x = in.points[i].x; p.y = in.points[i].y; p.z = in.points[i].z;
        stub::set_intensity(p, stub::get_intensity(in.points[i], 0), 0);
        out.points[i] = p;
    }
    out.width = static_cast<uint32_t>(out.points.size());
    out.height = 1;
    out.is_dense = in.is_dense;
}

// pcl/common/eigen.h getTransformation(x, y, z, roll, pitch, yaw): the float matrix written out term by term
inline Eigen::Affine3f getTransformation(float x, float y, float z, float roll, float pitch, float yaw) {
    const float A = std::cos(yaw), B = std::sin(yaw), C = std::cos(pitch), D = std::sin(pitch), E = std::cos(roll), F = std::sin(roll), DE = D * E, DF = D * F;
    Eigen::Affine3f t;
    t(0, 0) = A * C; t(0, 1) = A * DF - B * E; t(0, 2) = B * F + A * DE; t(0, 3) = x;
    t(1, 0) = B * C; t(1, 1) = A * E + B * DF; t(1, 2) = B * DE - A * F; t(1, 3) = y;
    t(2, 0) = -D;    t(2, 1) = C * F;          t(2, 2) = C * E;          t(2, 3) = z;
    return t;
}
// pcl/common/eigen.h getTranslationAndEulerAngles
inline void getTranslationAndEulerAngles(const Eigen::Affine3f& t, float& x, float& y, float& z, float& roll, float& pitch, float& yaw) {
    x = t(0, 3); y = t(1, 3); z = t(2, 3);
    roll = std::atan2(t(2, 1), t(2, 2));
    pitch = std::asin(-t(2, 0));
    yaw = std::atan2(t(1, 0), t(0, 0));
}
template <typename PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix<double, 4, 4>& T) {
    PointCloud<PointT> r = in;
    for (auto& p : r.points) {
        const double x = p.x, y = p.y, z = p.z;
        p.x = static_cast<float>(T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3));
        p.y = static_cast<float>(T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3));
        p.z = static_cast<float>(T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3));
    }
    out = r;
}
namespace io {
// the "PCD file" of the driver: a registry of raw xyz arrays by path suffix (no file format is modelled)
struct MapRegistry {
    std::vector<float> xyz;
    static MapRegistry& get() { static MapRegistry r; return r; }
};
template <typename PointT>
int loadPCDFile(const std::string&, PointCloud<PointT>& cloud) {
    const std::vector<float>& v = MapRegistry::get().xyz;
    cloud.points.resize(v.size() / 3);
    for (std::size_t i = 0; i < cloud.points.size(); ++i) {
        PointT p{};
        p.x = v[3 * i]; p.y = v[3 * i + 1]; p.z = v[3 * i + 2];
        cloud.points[i] = p;
    }
    cloud.width = static_cast<uint32_t>(cloud.points.size());
    cloud.height = 1;
    return 0;
}
}  // namespace io
}  // namespace pcl
