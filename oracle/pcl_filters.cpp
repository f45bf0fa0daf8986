// ORACLE — TEST INFRASTRUCTURE ONLY (see types.h).  Restates PCL 1.8.1 filters (un-vendored, absent here): parity unpinned for these library internals.
// Restates pcl::VoxelGrid<PointXYZI>::applyFilter and pcl::CropBox<PointXYZI>::applyFilter of PCL 1.8.1
// (un-vendored dependency, pinned by README.md:31-33 -> ROS Melodic), as configured at
// src/odomEstimationClass.cpp:13-14,137-142,270-292 and src/laserMappingClass.cpp:31,175-184
// (downsample_all_data_=true, min_points_per_voxel_=0, no filter-field limits; CropBox: identity transform, negative=false).
#include "floam_oracle.h"
#include <algorithm>
#include <cfloat>
#include <cstdint>

namespace fo {

namespace {
struct cloud_point_index_idx {  // pcl/filters/voxel_grid.h
  unsigned int idx;
  unsigned int cloud_point_index;
  bool operator<(const cloud_point_index_idx& p) const { return idx < p.idx; }
};
}  // namespace

void voxel_grid_filter(const PointXYZI* in, size_t n_in, float leaf, CloudI& out, bool total_order, bool* passthrough) {
  if (passthrough) *passthrough = false;
  CloudI result;
  if (n_in == 0) { out.swap(result); return; }
  const float inverse_leaf_size = 1.0f / leaf;  // Eigen::Array4f::Ones()/leaf_size_.array()

  // getMinMax3D
  float min_p[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, max_p[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (size_t pi = 0; pi < n_in; ++pi) {
    const PointXYZI& p = in[pi];
    min_p[0] = std::min(min_p[0], p.x); min_p[1] = std::min(min_p[1], p.y); min_p[2] = std::min(min_p[2], p.z);
    max_p[0] = std::max(max_p[0], p.x); max_p[1] = std::max(max_p[1], p.y); max_p[2] = std::max(max_p[2], p.z);
  }
  std::int64_t dx = static_cast<std::int64_t>((max_p[0] - min_p[0]) * inverse_leaf_size) + 1;
  std::int64_t dy = static_cast<std::int64_t>((max_p[1] - min_p[1]) * inverse_leaf_size) + 1;
  std::int64_t dz = static_cast<std::int64_t>((max_p[2] - min_p[2]) * inverse_leaf_size) + 1;
  if ((dx * dy * dz) > static_cast<std::int64_t>(std::numeric_limits<std::int32_t>::max())) {
    // "Leaf size is too small for the input dataset. Integer indices would overflow." -> output = *input_ (Q13)
    if (passthrough) *passthrough = true;
    result.assign(in, in + n_in);
    out.swap(result);
    return;
  }
  int min_b[3], max_b[3], div_b[3], divb_mul[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = static_cast<int>(std::floor(min_p[a] * inverse_leaf_size));
    max_b[a] = static_cast<int>(std::floor(max_p[a] * inverse_leaf_size));
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  divb_mul[0] = 1; divb_mul[1] = div_b[0]; divb_mul[2] = div_b[0] * div_b[1];

  std::vector<cloud_point_index_idx> index_vector;
  index_vector.reserve(n_in);
  for (size_t cp = 0; cp < n_in; ++cp) {
    int ijk0 = static_cast<int>(std::floor(in[cp].x * inverse_leaf_size) - static_cast<float>(min_b[0]));
    int ijk1 = static_cast<int>(std::floor(in[cp].y * inverse_leaf_size) - static_cast<float>(min_b[1]));
    int ijk2 = static_cast<int>(std::floor(in[cp].z * inverse_leaf_size) - static_cast<float>(min_b[2]));
    int idx = ijk0 * divb_mul[0] + ijk1 * divb_mul[1] + ijk2 * divb_mul[2];
    index_vector.push_back(cloud_point_index_idx{static_cast<unsigned int>(idx), static_cast<unsigned int>(cp)});
  }
  if (total_order) std::stable_sort(index_vector.begin(), index_vector.end(), std::less<cloud_point_index_idx>());
  else std::sort(index_vector.begin(), index_vector.end(), std::less<cloud_point_index_idx>());

  // one output per run of equal idx; CentroidPoint<PointXYZI>: Vector3f xyz sum + float intensity sum, then / n
  size_t index = 0;
  while (index < index_vector.size()) {
    size_t i = index + 1;
    while (i < index_vector.size() && index_vector[i].idx == index_vector[index].idx) ++i;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (size_t li = index; li < i; ++li) {
      const PointXYZI& p = in[index_vector[li].cloud_point_index];
      sx += p.x; sy += p.y; sz += p.z; si += p.intensity;
    }
    const float n = static_cast<float>(i - index);
    result.push_back(make_xyzi(sx / n, sy / n, sz / n, si / n));
    index = i;
  }
  out.swap(result);
}

void crop_box_filter(const PointXYZI* in, size_t n_in, const float mn[3], const float mx[3], CloudI& out) {
  CloudI result;
  result.reserve(n_in);
  for (size_t pi = 0; pi < n_in; ++pi) {
    const PointXYZI& p = in[pi];
    if ((p.x < mn[0] || p.y < mn[1] || p.z < mn[2]) || (p.x > mx[0] || p.y > mx[1] || p.z > mx[2])) continue;
    result.push_back(p);
  }
  out.swap(result);
}

}  // namespace fo
