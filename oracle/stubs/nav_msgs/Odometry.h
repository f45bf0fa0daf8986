// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <nav_msgs/Odometry.h>.
#pragma once
#include <boost/shared_ptr.hpp>
#include <geometry_msgs/types.h>
#include <std_msgs/Header.h>
namespace nav_msgs {
struct Odometry {
  std_msgs::Header header;
  std::string child_frame_id;
  geometry_msgs::PoseWithCovariance pose;
  geometry_msgs::TwistWithCovariance twist;
  typedef boost::shared_ptr<Odometry> Ptr;
  typedef boost::shared_ptr<Odometry const> ConstPtr;
};
typedef boost::shared_ptr<Odometry> OdometryPtr;
typedef boost::shared_ptr<Odometry const> OdometryConstPtr;
}  // namespace nav_msgs
