// sensor_msgs/PointCloud2 ingestion on the device: the raw message bytes are uploaded as they are and re-packed into the 32-byte
// vel_point::PointXYZIRT layout by a kernel — the job pcl::fromROSMsg does on the host in the reference
// (src/laserProcessingNode.cpp:98; field matching by name AND datatype, pcl/conversions.h FieldMatches [ext]).
#pragma once
#include "common.cuh"
#include "floam_b200.h"

namespace floam {

// raw: device copy of msg.data; out: n = width * height points. Fields with offset < 0 stay 0 (what fromROSMsg leaves when the
// message has no field of that name and datatype). Multi-byte values are read byte by byte: point_step need not be a multiple of 4
// (the Velodyne driver's XYZIRT layout is 22 bytes).
void unpack_pointcloud2_device(const unsigned char* d_raw, const floam_pc2_layout& layout, PointIRT* d_out, cudaStream_t s);

}  // namespace floam
