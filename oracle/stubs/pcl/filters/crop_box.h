// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl/filters/crop_box.h>; body = oracle/pcl_filters.cpp (PCL 1.8.1, Appendix A.2).
#pragma once
#include <pcl/filters/filter.h>
#include <pcl/common/transforms.h>  // the real crop_box.h includes it (src/laserProcessingNode.cpp:116 gets transformPointCloud this way)
#include <cstring>
#include "../../../floam_oracle.h"
namespace pcl {
template <typename PointT>
class CropBox : public Filter<PointT> {
 public:
  CropBox() : negative_(false) { for (int i = 0; i < 3; ++i) { min_[i] = -1.f; max_[i] = 1.f; } }
  inline void setMin(const Eigen::Vector4f& min_pt) { for (int i = 0; i < 3; ++i) min_[i] = min_pt(i); }
  inline void setMax(const Eigen::Vector4f& max_pt) { for (int i = 0; i < 3; ++i) max_[i] = max_pt(i); }
  inline void setNegative(bool negative) { negative_ = negative; }
 protected:
  void applyFilter(pcl::PointCloud<PointT>& output) {
    if (negative_) { std::fprintf(stderr, "CropBox stand-in: negative=true is not on the path\n"); std::abort(); }
    fo::CloudI out;
    const pcl::PointCloud<PointT>& in = *this->input_;
    fo::crop_box_filter(reinterpret_cast<const fo::PointXYZI*>(in.points.data()), in.points.size(), min_, max_, out);
    const pcl::PCLHeader header = in.header;
    output.points.resize(out.size());
    if (!out.empty()) std::memcpy(static_cast<void*>(output.points.data()), out.data(), out.size() * sizeof(PointT));
    output.header = header;
    output.height = 1;
    output.is_dense = true;
    output.width = static_cast<std::uint32_t>(output.points.size());
  }
  float min_[3], max_[3];
  bool negative_;
};
}  // namespace pcl
