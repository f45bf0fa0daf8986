// ORACLE — TEST INFRASTRUCTURE ONLY: stand-ins for the geometry_msgs types the reference touches.
#pragma once
namespace geometry_msgs {
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };  // default message: all zero
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; double covariance[36] = {0}; };
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; double covariance[36] = {0}; };
}  // namespace geometry_msgs
