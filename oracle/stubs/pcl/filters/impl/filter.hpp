// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl/filters/impl/filter.hpp> (PCL 1.8.1).
#pragma once
#include <pcl/filters/filter.h>
#include <cmath>
// index-only overload: the cloud is NOT modified (src/laserProcessingClass.cpp:74-75)
template <typename PointT> void pcl::removeNaNFromPointCloud(const pcl::PointCloud<PointT>& cloud_in, std::vector<int>& index) {
  index.resize(cloud_in.points.size());
  if (cloud_in.is_dense) {
    for (int j = 0; j < static_cast<int>(cloud_in.points.size()); ++j) index[j] = j;
  } else {
    int j = 0;
    for (int i = 0; i < static_cast<int>(cloud_in.points.size()); ++i) {
      if (!std::isfinite(cloud_in.points[i].x) || !std::isfinite(cloud_in.points[i].y) || !std::isfinite(cloud_in.points[i].z)) continue;
      index[j] = i;
      j++;
    }
    if (j != static_cast<int>(cloud_in.points.size())) index.resize(j);
  }
}
template <typename PointT> void pcl::removeNaNFromPointCloud(const pcl::PointCloud<PointT>& cloud_in, pcl::PointCloud<PointT>& cloud_out, std::vector<int>& index) {
  if (&cloud_in != &cloud_out) { cloud_out.header = cloud_in.header; cloud_out.points.resize(cloud_in.points.size()); }
  index.resize(cloud_in.points.size());
  size_t j = 0;
  if (cloud_in.is_dense) {
    cloud_out = cloud_in;
    for (j = 0; j < cloud_out.points.size(); ++j) index[j] = static_cast<int>(j);
  } else {
    for (size_t i = 0; i < cloud_in.points.size(); ++i) {
      if (!std::isfinite(cloud_in.points[i].x) || !std::isfinite(cloud_in.points[i].y) || !std::isfinite(cloud_in.points[i].z)) continue;
      cloud_out.points[j] = cloud_in.points[i];
      index[j] = static_cast<int>(i);
      j++;
    }
    if (j != cloud_in.points.size()) { cloud_out.points.resize(j); index.resize(j); }
    cloud_out.height = 1;
    cloud_out.width = static_cast<std::uint32_t>(j);
    cloud_out.is_dense = true;
  }
}
