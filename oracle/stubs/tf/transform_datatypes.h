// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <tf/transform_datatypes.h> (node sources only: syntax check).
#pragma once
#include <geometry_msgs/types.h>
#include <ros/time.h>
#include <cmath>
#include <string>
namespace tf {
struct Vector3 { double x_, y_, z_; Vector3(double x = 0, double y = 0, double z = 0) : x_(x), y_(y), z_(z) {} };
struct Quaternion {
  double x_, y_, z_, w_;
  Quaternion(double x = 0, double y = 0, double z = 0, double w = 1) : x_(x), y_(y), z_(z), w_(w) {}
};
inline void quaternionMsgToTF(const geometry_msgs::Quaternion& m, Quaternion& q) { q = Quaternion(m.x, m.y, m.z, m.w); }
struct Matrix3x3 {
  double m[3][3];
  explicit Matrix3x3(const Quaternion& q) {
    const double d = q.x_ * q.x_ + q.y_ * q.y_ + q.z_ * q.z_ + q.w_ * q.w_, s = 2.0 / d;
    const double xs = q.x_ * s, ys = q.y_ * s, zs = q.z_ * s, wx = q.w_ * xs, wy = q.w_ * ys, wz = q.w_ * zs;
    const double xx = q.x_ * xs, xy = q.x_ * ys, xz = q.x_ * zs, yy = q.y_ * ys, yz = q.y_ * zs, zz = q.z_ * zs;
    m[0][0] = 1 - (yy + zz); m[0][1] = xy - wz; m[0][2] = xz + wy;
    m[1][0] = xy + wz; m[1][1] = 1 - (xx + zz); m[1][2] = yz - wx;
    m[2][0] = xz - wy; m[2][1] = yz + wx; m[2][2] = 1 - (xx + yy);
  }
  void getRPY(double& roll, double& pitch, double& yaw) const {
    pitch = std::asin(-m[2][0]); roll = std::atan2(m[2][1], m[2][2]); yaw = std::atan2(m[1][0], m[0][0]);
  }
};
struct Transform {
  Vector3 origin; Quaternion rotation;
  void setOrigin(const Vector3& v) { origin = v; }
  void setRotation(const Quaternion& q) { rotation = q; }
};
struct StampedTransform : Transform {
  ros::Time stamp_; std::string frame_id_, child_frame_id_;
  StampedTransform(const Transform& t, const ros::Time& s, const std::string& f, const std::string& c) : Transform(t), stamp_(s), frame_id_(f), child_frame_id_(c) {}
};
}  // namespace tf
