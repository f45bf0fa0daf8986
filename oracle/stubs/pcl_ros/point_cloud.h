// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl_ros/point_cloud.h> (include/lidar.h:11): pcl clouds as ROS messages.
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl_conversions/pcl_conversions.h>
#include <ros/ros.h>
