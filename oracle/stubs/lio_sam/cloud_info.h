// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <lio_sam/cloud_info.h> (LIO-SAM message; src/odomEstimationNode.cpp:122-164, syntax check).
#pragma once
#include <sensor_msgs/PointCloud2.h>
#include <std_msgs/Header.h>
namespace lio_sam {
struct cloud_info {
  std_msgs::Header header;
  std::int64_t imuAvailable = 0, odomAvailable = 0;
  float imuRollInit = 0, imuPitchInit = 0, imuYawInit = 0;
  float initialGuessX = 0, initialGuessY = 0, initialGuessZ = 0, initialGuessRoll = 0, initialGuessPitch = 0, initialGuessYaw = 0;
  sensor_msgs::PointCloud2 cloud_deskewed, cloud_corner, cloud_surface;
};
}  // namespace lio_sam
