// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl_conversions/pcl_conversions.h> (ROS Melodic): stamp conversions as the
// real header does them (pcl stamps are microseconds), and declarations of from/toROSMsg for the node sources.
#pragma once
#include <pcl/point_cloud.h>
#include <ros/time.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl_conversions {
inline void fromPCL(const std::uint64_t& pcl_stamp, ros::Time& stamp) { stamp.fromNSec(pcl_stamp * 1000ull); }
inline void toPCL(const ros::Time& stamp, std::uint64_t& pcl_stamp) { pcl_stamp = stamp.toNSec() / 1000ull; }
inline ros::Time fromPCL(const std::uint64_t& pcl_stamp) { ros::Time stamp; fromPCL(pcl_stamp, stamp); return stamp; }
inline std::uint64_t toPCL(const ros::Time& stamp) { std::uint64_t pcl_stamp; toPCL(stamp, pcl_stamp); return pcl_stamp; }
}  // namespace pcl_conversions
namespace pcl {
template <typename T> void fromROSMsg(const sensor_msgs::PointCloud2& cloud, pcl::PointCloud<T>& pcl_cloud);
template <typename T> void toROSMsg(const pcl::PointCloud<T>& pcl_cloud, sensor_msgs::PointCloud2& cloud);
}  // namespace pcl
