// Drop-in for include/dataHandler.h: dmapping::ImuHandler, Compensate, CompensateVelocity (src/dataHandler.cpp:24-122) plus the
// fused CenterTime + Compensate + alignment call the laserProcessing node can use instead of its three separate steps.
#ifndef FLOAM_B200_HOST_DATA_HANDLER_H_
#define FLOAM_B200_HOST_DATA_HANDLER_H_
#include <cstdio>
#ifdef FLOAM_B200_WITH_PCL   // include/dataHandler.h:4-12
#include "ros/ros.h"
#include <map>
#include <algorithm>
#include "sensor_msgs/Imu.h"
#include "pcl_conversions/pcl_conversions.h"
#include "math.h"
using std::cout;
using std::endl;
#endif
#include <lidar.h>       // through the include path (this directory first), so that lidar.h's #include_next finds the reference's

#define SCAN_RATE 10.0
namespace dmapping {

inline Eigen::Quaterniond Imu2Orientation(const sensor_msgs::Imu& data) {  // :6-8
  return Eigen::Quaterniond(data.orientation.w, data.orientation.x, data.orientation.y, data.orientation.z);
}
inline Eigen::Vector3d Imu2AngularVelocity(const sensor_msgs::Imu& data) {  // :10-12
  return Eigen::Vector3d(data.angular_velocity.x, data.angular_velocity.y, data.angular_velocity.z);
}
inline Eigen::Vector3d Imu2LinearAcceleration(const sensor_msgs::Imu& data) {  // :14-16
  return Eigen::Vector3d(data.linear_acceleration.x, data.linear_acceleration.y, data.linear_acceleration.z);
}

class ImuHandler {
 public:
  ImuHandler() : owned_(new floam_b200_host::FloamContext()), fc_(owned_.get()) { small_capacities(); }
  explicit ImuHandler(floam_b200_host::FloamContext* shared) : fc_(shared) {}
  void AddMsg(sensor_msgs::Imu::ConstPtr msg) {  // :24-40
    if (fc_->ensure()) return;
    const double q[4] = {msg->orientation.x, msg->orientation.y, msg->orientation.z, msg->orientation.w};
    floam_imu_push(fc_->ctx, msg->header.stamp.toSec(), q);
  }
  bool Get(const double& tStamp, sensor_msgs::Imu& data) const {  // :51-69 (zero-order hold)
    if (fc_->ensure()) return false;
    double q[4]; int valid = 0;
    floam_imu_get(fc_->ctx, tStamp, q, &valid);
    if (valid) { data.orientation.x = q[0]; data.orientation.y = q[1]; data.orientation.z = q[2]; data.orientation.w = q[3]; }
    return valid != 0;
  }
  sensor_msgs::Imu Get(const double& tStamp) const { sensor_msgs::Imu data; Get(tStamp, data); return data; }  // :71-75
  bool TimeContained(const double t) const {  // :76-81
    if (fc_->ensure()) return false;
    int contained = 0;
    floam_imu_time_contained(fc_->ctx, t, &contained);
    return contained != 0;
  }
  std::size_t size() { int n = 0; if (!fc_->ensure()) floam_imu_size(fc_->ctx, &n); return (std::size_t)n; }
  floam_b200_host::FloamContext* context() const { return fc_; }

 private:
  void small_capacities() { fc_->prm.max_map_points = 1 << 16; fc_->prm.max_global_map_points = 0; fc_->prm.max_grid_cells = 1 << 16; }
  std::unique_ptr<floam_b200_host::FloamContext> owned_;
  floam_b200_host::FloamContext* fc_;
};

// :93-122. `compensated` receives the rotation-deskewed cloud; returns false ("no imu data") when the scan is not covered.
inline bool Compensate(pcl::PointCloud<vel_point::PointXYZIRT>& input, pcl::PointCloud<vel_point::PointXYZIRT>& compensated, ImuHandler& handler,
                       Eigen::Quaterniond& extrinsics) {
  compensated.points = input.points;
  compensated.width = input.width; compensated.height = input.height;
  if (handler.context()->ensure()) return false;
  double q[4];
  floam_b200_host::quaternion_to_xyzw(extrinsics, q);
  std::uint64_t stamp = input.header.stamp;
  const int rc = floam_deskew_align_ex(handler.context()->ctx, reinterpret_cast<floam_point_xyzirt*>(compensated.points.data()), (int)compensated.points.size(),
                                       &stamp, q, FLOAM_DESKEW_COMPENSATE);
  if (rc == FLOAM_NO_IMU) { std::printf("no imu data\n"); return false; }
  floam_b200_host::report(rc, "dmapping::Compensate");
  return rc == FLOAM_OK;
}

// CenterTime (src/laserProcessingNode.cpp:65-78) + Compensate + IMU alignment (:113-116) in one device pass, in place.
inline bool CenterCompensateAlign(pcl::PointCloud<vel_point::PointXYZIRT>& cloud, ImuHandler& handler, Eigen::Quaterniond& extrinsics) {
  if (handler.context()->ensure()) return false;
  double q[4];
  floam_b200_host::quaternion_to_xyzw(extrinsics, q);
  std::uint64_t stamp = cloud.header.stamp;
  const int rc = floam_deskew_align(handler.context()->ctx, reinterpret_cast<floam_point_xyzirt*>(cloud.points.data()), (int)cloud.points.size(), &stamp, q);
  cloud.header.stamp = stamp;
  return rc == FLOAM_OK;
}

inline void CompensateVelocity(pcl::PointCloud<vel_point::PointXYZIRT>::Ptr input, const Eigen::Vector3d& velocity, floam_b200_host::FloamContext* fc) {  // :82-91
  if (fc->ensure()) return;
  const double v[3] = {velocity(0), velocity(1), velocity(2)};
  floam_b200_host::report(floam_compensate_velocity(fc->ctx, reinterpret_cast<floam_point_xyzirt*>(input->points.data()), (int)input->points.size(), v),
                          "dmapping::CompensateVelocity");
}

// the reference's two-argument signature (include/dataHandler.h:64); served by one small process-wide context
inline void CompensateVelocity(pcl::PointCloud<vel_point::PointXYZIRT>::Ptr input, const Eigen::Vector3d& velocity) {
  static floam_b200_host::FloamContext shared;
  static bool sized = false;
  if (!sized) { shared.prm.max_map_points = 1 << 16; shared.prm.max_global_map_points = 0; shared.prm.max_grid_cells = 1 << 16; sized = true; }
  CompensateVelocity(input, velocity, &shared);
}

}  // namespace dmapping
#endif
