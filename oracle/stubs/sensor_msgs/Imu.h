// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <sensor_msgs/Imu.h>.
#pragma once
#include <boost/shared_ptr.hpp>
#include <geometry_msgs/types.h>
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct Imu {
  std_msgs::Header header;
  geometry_msgs::Quaternion orientation;
  double orientation_covariance[9] = {0};
  geometry_msgs::Vector3 angular_velocity;
  double angular_velocity_covariance[9] = {0};
  geometry_msgs::Vector3 linear_acceleration;
  double linear_acceleration_covariance[9] = {0};
  typedef boost::shared_ptr<Imu> Ptr;
  typedef boost::shared_ptr<Imu const> ConstPtr;
};
typedef boost::shared_ptr<Imu> ImuPtr;
typedef boost::shared_ptr<Imu const> ImuConstPtr;
}  // namespace sensor_msgs
