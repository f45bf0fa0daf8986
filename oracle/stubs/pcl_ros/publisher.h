// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl_ros/publisher.h> (include/lidar.h:12).
#pragma once
#include <pcl_ros/point_cloud.h>
