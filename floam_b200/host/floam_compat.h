// Types the reference's class API is written in.  With -DFLOAM_B200_WITH_PCL the real PCL / Eigen / ROS headers are used (the
// drop-in build inside the catkin workspace); otherwise minimal stand-ins with the same names and members keep the shims
// compilable on a box that has none of them (this repo's build container).  Only what the L2 class API touches is modelled.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#ifdef FLOAM_B200_WITH_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <sensor_msgs/Imu.h>
#include <sensor_msgs/PointCloud2.h>
#else
namespace Eigen {
struct Vector3d {
  double v[3] = {0, 0, 0};
  Vector3d() {}
  Vector3d(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
struct Quaterniond {  // coefficient order (x, y, z, w) like Eigen's coeffs()
  double c[4] = {0, 0, 0, 1};
  Quaterniond() {}
  Quaterniond(double w, double x, double y, double z) { c[0] = x; c[1] = y; c[2] = z; c[3] = w; }
  double x() const { return c[0]; }
  double y() const { return c[1]; }
  double z() const { return c[2]; }
  double w() const { return c[3]; }
  const double* coeffs_data() const { return c; }
};
struct Isometry3d {  // row-major 4x4
  double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  static Isometry3d Identity() { return Isometry3d(); }
  Vector3d translation() const { return Vector3d(m[3], m[7], m[11]); }
  double operator()(int r, int c) const { return m[r * 4 + c]; }
  double& operator()(int r, int c) { return m[r * 4 + c]; }
};
}  // namespace Eigen

namespace pcl {
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, data3 = 1.0f;
  float intensity = 0, pad[3] = {0, 0, 0};
};
struct PCLHeader {
  std::uint32_t seq = 0;
  std::uint64_t stamp = 0;  // microseconds
  std::string frame_id;
};
template <class T>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<T>> Ptr;  // boost::shared_ptr in PCL 1.8
  typedef std::shared_ptr<const PointCloud<T>> ConstPtr;
  PCLHeader header;
  std::vector<T> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  void resize(std::size_t n) { points.resize(n); width = (std::uint32_t)n; height = 1; }
  void clear() { points.clear(); width = 0; }
  void push_back(const T& p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = (std::uint32_t)points.size(); height = 1; is_dense = is_dense && o.is_dense;
    return *this;
  }
};
}  // namespace pcl

namespace sensor_msgs {
struct PointField {
  enum { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };
  std::string name;
  std::uint32_t offset = 0;
  std::uint8_t datatype = 0;
  std::uint32_t count = 1;
};
struct PointCloud2 {
  struct { struct { double t = 0; double toSec() const { return t; } } stamp; std::string frame_id; } header;
  std::uint32_t height = 1, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  std::uint32_t point_step = 0, row_step = 0;
  std::vector<std::uint8_t> data;
  bool is_dense = true;
  typedef std::shared_ptr<const PointCloud2> ConstPtr;
};
struct Imu {
  struct { struct { double t = 0; double toSec() const { return t; } } stamp; } header;
  struct { double x = 0, y = 0, z = 0, w = 0; } orientation;  // default message: all-zero quaternion
  struct { double x = 0, y = 0, z = 0; } angular_velocity, linear_acceleration;
  typedef std::shared_ptr<const Imu> ConstPtr;
};
}  // namespace sensor_msgs
#endif

namespace floam_b200_host {
inline void isometry_to_rowmajor(const Eigen::Isometry3d& T, double out[16]) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out[r * 4 + c] = r < 3 ? T(r, c) : (c == 3 ? 1.0 : 0.0);
}
inline void rowmajor_to_isometry(const double in[16], Eigen::Isometry3d& T) {
#ifdef FLOAM_B200_WITH_PCL
  T = Eigen::Isometry3d::Identity();
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T.linear()(r, c) = in[r * 4 + c];
    T.translation()(r) = in[r * 4 + 3];
  }
#else
  for (int i = 0; i < 16; ++i) T.m[i] = in[i];
#endif
}
inline void quaternion_to_xyzw(const Eigen::Quaterniond& q, double out[4]) { out[0] = q.x(); out[1] = q.y(); out[2] = q.z(); out[3] = q.w(); }
}  // namespace floam_b200_host
