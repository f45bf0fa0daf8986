// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl/filters/voxel_grid.h>.  pcl::VoxelGrid<PointXYZI>::applyFilter is the
// restatement in oracle/pcl_filters.cpp (PCL 1.8.1) behind PCL's own class interface.
#pragma once
#include <pcl/filters/filter.h>
#include <cstring>
#include "../../../floam_oracle.h"
namespace pcl {
namespace floam_stub {
// test knob (not in PCL): stable order inside a voxel = the total-order contract of the CUDA path; default = std::sort like PCL
inline bool& voxel_total_order() { static bool v = false; return v; }
}  // namespace floam_stub
template <typename PointT>
class VoxelGrid : public Filter<PointT> {
  static_assert(sizeof(PointT) == sizeof(fo::PointXYZI), "stand-in handles the 32-byte XYZI layout only");
 public:
  VoxelGrid() { leaf_size_[0] = leaf_size_[1] = leaf_size_[2] = 0.f; }
  inline void setLeafSize(float lx, float ly, float lz) { leaf_size_[0] = lx; leaf_size_[1] = ly; leaf_size_[2] = lz; }
 protected:
  void applyFilter(pcl::PointCloud<PointT>& output) {
    fo::CloudI out;
    const pcl::PointCloud<PointT>& in = *this->input_;
    fo::voxel_grid_filter(reinterpret_cast<const fo::PointXYZI*>(in.points.data()), in.points.size(), leaf_size_[0], out, floam_stub::voxel_total_order());
    const pcl::PCLHeader header = in.header;  // `in` may alias `output`
    output.points.resize(out.size());
    if (!out.empty()) std::memcpy(static_cast<void*>(output.points.data()), out.data(), out.size() * sizeof(PointT));
    output.header = header;
    output.height = 1;
    output.is_dense = true;
    output.width = static_cast<std::uint32_t>(output.points.size());
  }
  float leaf_size_[3];
};
}  // namespace pcl
