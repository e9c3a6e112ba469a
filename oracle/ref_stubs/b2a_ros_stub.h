// b2a_ros_stub.h -- ROS / tf2 / message types named by /root/reference/src/aruco_slam.cpp and
// src/map_loader.cpp, so that the UNMODIFIED reference sources compile without ROS (oracle/_ref build).
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  Messages are plain structs with the fields the reference
// fills; logging macros type-check their arguments and print nothing; ros::Time::now() reads an injectable
// clock (b2a_ref_clock, set by the harness) instead of the wall clock (aruco_slam.cpp:26,31-32).
// tf2::Quaternion::setRPY, tf2::Matrix3x3::getRotation, tf2::toMsg and tf2::doTransform(Pose) follow tf2's
// published formulas (setRPY: fixed axes X-Y-Z half-angle products; getRotation: Shepperd's trace method;
// doTransform: out = T * in on position and orientation).
#ifndef B2A_ROS_STUB_H
#define B2A_ROS_STUB_H
#include <array>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

extern "C" double b2a_ref_clock;      // seconds; defined in oracle/ref_harness.cpp

namespace ros {
struct Duration {
    double sec;
    Duration() : sec(0) {}
    Duration(double s) : sec(s) {}
    double toSec() const { return sec; }
};
struct Time {
    double sec;
    Time() : sec(0) {}
    static Time now() { Time t; t.sec = b2a_ref_clock; return t; }
    Duration operator-(const Time &o) const { return Duration(sec - o.sec); }
};
}  // namespace ros

#define B2A_ROS_STREAM_NOOP(args) do { if (false) { std::ostringstream b2a_os__; b2a_os__ << args; } } while (0)
#define ROS_INFO_STREAM(args) B2A_ROS_STREAM_NOOP(args)
#define ROS_ERROR_STREAM(args) B2A_ROS_STREAM_NOOP(args)
#define ROS_WARN_STREAM(args) B2A_ROS_STREAM_NOOP(args)
#define ROS_INFO_STREAM_ONCE(args) B2A_ROS_STREAM_NOOP(args)
#define B2A_ROS_PRINTF_NOOP(...) do { if (false) std::printf(__VA_ARGS__); } while (0)
#define ROS_INFO(...) B2A_ROS_PRINTF_NOOP(__VA_ARGS__)
#define ROS_ERROR(...) B2A_ROS_PRINTF_NOOP(__VA_ARGS__)
#define ROS_DEBUG(...) B2A_ROS_PRINTF_NOOP(__VA_ARGS__)
#define ROS_WARN(...) B2A_ROS_PRINTF_NOOP(__VA_ARGS__)

namespace std_msgs {
struct Header { std::string frame_id; };
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
}  // namespace std_msgs

namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::Header header; std::string child_frame_id; Transform transform; };
struct PoseWithCovariance { Pose pose; std::array<double, 36> covariance{}; };
struct PoseWithCovarianceStamped { std_msgs::Header header; PoseWithCovariance pose; };
}  // namespace geometry_msgs

namespace visualization_msgs {
struct Marker {
    enum { ARROW = 0, CUBE = 1, SPHERE = 2 };
    std_msgs::Header header;
    int id = 0, type = 0;
    geometry_msgs::Pose pose;
    geometry_msgs::Vector3 scale;
    std_msgs::ColorRGBA color;
    ros::Duration lifetime;
};
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs

namespace tf2 {
struct Vector3 {
    double v[3];
    Vector3() : v{0, 0, 0} {}
    Vector3(double x, double y, double z) : v{x, y, z} {}
};
class Quaternion {
public:
    Quaternion() : q_{0, 0, 0, 1} {}
    Quaternion(double x, double y, double z, double w) : q_{x, y, z, w} {}
    void setRPY(double roll, double pitch, double yaw)
    {
        const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
        const double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
        q_[0] = sr * cp * cy - cr * sp * sy;
        q_[1] = cr * sp * cy + sr * cp * sy;
        q_[2] = cr * cp * sy - sr * sp * cy;
        q_[3] = cr * cp * cy + sr * sp * sy;
    }
    double x() const { return q_[0]; }
    double y() const { return q_[1]; }
    double z() const { return q_[2]; }
    double w() const { return q_[3]; }
private:
    double q_[4];
};
inline Quaternion operator*(const Quaternion &a, const Quaternion &b)
{
    return Quaternion(a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x(),
                      a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z());
}
class Matrix3x3 {
public:
    Matrix3x3() : m_{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}} {}
    Matrix3x3(double xx, double xy, double xz, double yx, double yy, double yz, double zx, double zy, double zz) : m_{{xx, xy, xz}, {yx, yy, yz}, {zx, zy, zz}} {}
    void getRotation(Quaternion &q) const
    {
        const double trace = m_[0][0] + m_[1][1] + m_[2][2];
        double t[4];
        if (trace > 0.0) {
            double s = std::sqrt(trace + 1.0);
            t[3] = s * 0.5;
            s = 0.5 / s;
            t[0] = (m_[2][1] - m_[1][2]) * s;
            t[1] = (m_[0][2] - m_[2][0]) * s;
            t[2] = (m_[1][0] - m_[0][1]) * s;
        } else {
            const int i = m_[0][0] < m_[1][1] ? (m_[1][1] < m_[2][2] ? 2 : 1) : (m_[0][0] < m_[2][2] ? 2 : 0);
            const int j = (i + 1) % 3, k = (i + 2) % 3;
            double s = std::sqrt(m_[i][i] - m_[j][j] - m_[k][k] + 1.0);
            t[i] = s * 0.5;
            s = 0.5 / s;
            t[3] = (m_[k][j] - m_[j][k]) * s;
            t[j] = (m_[j][i] + m_[i][j]) * s;
            t[k] = (m_[k][i] + m_[i][k]) * s;
        }
        q = Quaternion(t[0], t[1], t[2], t[3]);
    }
    double m_[3][3];
};
class Transform {
public:
    Transform() {}
    Transform(const Matrix3x3 &b, const Vector3 &c) : basis_(b), origin_(c) {}
    Quaternion getRotation() const { Quaternion q; basis_.getRotation(q); return q; }
    const Vector3 &getOrigin() const { return origin_; }
private:
    Matrix3x3 basis_;
    Vector3 origin_;
};
inline geometry_msgs::Quaternion toMsg(const Quaternion &q)
{
    geometry_msgs::Quaternion m;
    m.x = q.x(); m.y = q.y(); m.z = q.z(); m.w = q.w();
    return m;
}
// tf2_geometry_msgs: pose_out = transform * pose_in
inline void doTransform(const geometry_msgs::Pose &in, geometry_msgs::Pose &out, const geometry_msgs::TransformStamped &t)
{
    const Quaternion qt(t.transform.rotation.x, t.transform.rotation.y, t.transform.rotation.z, t.transform.rotation.w);
    const Quaternion qi(in.orientation.x, in.orientation.y, in.orientation.z, in.orientation.w);
    // rotate the position by qt: p' = qt * (p, 0) * qt^-1
    const Quaternion p(in.position.x, in.position.y, in.position.z, 0.0);
    const Quaternion qc(-qt.x(), -qt.y(), -qt.z(), qt.w());
    const Quaternion r = qt * p * qc;
    const Quaternion qo = qt * qi;
    geometry_msgs::Pose o;
    o.position.x = r.x() + t.transform.translation.x;
    o.position.y = r.y() + t.transform.translation.y;
    o.position.z = r.z() + t.transform.translation.z;
    o.orientation = toMsg(qo);
    out = o;
}
}  // namespace tf2
#endif
