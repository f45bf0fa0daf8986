// Node-side adapter: sensor_msgs::PointCloud2 -> the fused device pipeline, without the host-side
// pcl::fromROSMsg copy of src/laserProcessingNode.cpp:98 and without the two TCPROS hops between the three reference nodes.
// The message bytes go to the GPU as they are; the field lookup below is what pcl::fromROSMsg's FieldMatches does for
// vel_point::PointXYZIRT (include/lidar.h:25-31): a field matches by NAME and DATATYPE, an unmatched field stays zero.
#pragma once
#include <cstdio>

#include "floam_b200.h"
#include "floam_compat.h"

namespace floam_b200_host {

inline floam_pc2_layout LayoutFromMsg(const sensor_msgs::PointCloud2& msg) {
  floam_pc2_layout L;
  L.width = msg.width; L.height = msg.height; L.point_step = msg.point_step; L.row_step = msg.row_step;
  L.off_x = L.off_y = L.off_z = L.off_intensity = L.off_ring = L.off_time = -1;
  L.is_bigendian = msg.is_bigendian ? 1 : 0;
  for (const auto& f : msg.fields) {
    const bool f32 = f.datatype == sensor_msgs::PointField::FLOAT32, u16 = f.datatype == sensor_msgs::PointField::UINT16;
    if (f.name == "x" && f32) L.off_x = (int32_t)f.offset;
    else if (f.name == "y" && f32) L.off_y = (int32_t)f.offset;
    else if (f.name == "z" && f32) L.off_z = (int32_t)f.offset;
    else if (f.name == "intensity" && f32) L.off_intensity = (int32_t)f.offset;
    else if (f.name == "ring" && u16) L.off_ring = (int32_t)f.offset;
    else if (f.name == "time" && f32) L.off_time = (int32_t)f.offset;
  }
  const int32_t* offs[6] = {&L.off_x, &L.off_y, &L.off_z, &L.off_intensity, &L.off_ring, &L.off_time};
  static const char* names[6] = {"x", "y", "z", "intensity", "ring", "time"};
  for (int k = 0; k < 6; ++k)
    if (*offs[k] < 0) std::fprintf(stderr, "Failed to find match for field '%s'.\n", names[k]);   // PCL's warning, same wording
  return L;
}

// velodyneHandler + laser_processing + odom_estimation of the reference nodes in one call: submit the message, get the pose of the
// previous one back (two frames in flight). Returns the C-ABI status; FLOAM_NO_IMU means "cannot compensate - no IMU data" (the
// reference skips such scans, src/laserProcessingNode.cpp:108-112).
class FusedOdometryNode {
 public:
  FusedOdometryNode(floam_ctx* ctx, bool use_imu, bool deskew, const Eigen::Quaterniond& extrinsics) : ctx_(ctx), use_imu_(use_imu), deskew_(deskew) {
    quaternion_to_xyzw(extrinsics, extr_);
  }
  void imuHandler(const sensor_msgs::Imu& m) {
    const double q[4] = {m.orientation.x, m.orientation.y, m.orientation.z, m.orientation.w};
    floam_imu_push(ctx_, m.header.stamp.toSec(), q);
  }
  // msg must stay alive until the next call (its bytes are read by an asynchronous upload)
  int velodyneHandler(const sensor_msgs::PointCloud2& msg, double pose_prev[7], bool* have_prev) {
    const floam_pc2_layout L = LayoutFromMsg(msg);
    uint64_t stamp_us = (uint64_t)(msg.header.stamp.toSec() * 1e6);   // pcl_conversions::toPCL: microseconds
    int rc = floam_process_submit_pc2(ctx_, msg.data.data(), &L, use_imu_ ? &stamp_us : nullptr, use_imu_ ? extr_ : nullptr, deskew_ ? 1 : 0);
    if (rc != FLOAM_OK) return rc;
    ++inflight_;
    *have_prev = false;
    if (inflight_ == 2) {
      rc = floam_process_wait(ctx_, pose_prev);
      --inflight_;
      *have_prev = true;
    }
    return rc;
  }
  int flush(double pose_last[7]) {
    if (inflight_ == 0) return FLOAM_ERR_ARG;
    --inflight_;
    return floam_process_wait(ctx_, pose_last);
  }

 private:
  floam_ctx* ctx_;
  bool use_imu_, deskew_;
  double extr_[4];
  int inflight_ = 0;
};

}  // namespace floam_b200_host
