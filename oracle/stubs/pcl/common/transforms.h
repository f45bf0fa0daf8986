// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl/common/transforms.h> (PCL 1.8.1): transformPointCloud with an affine transform.
#pragma once
#include <pcl/point_cloud.h>
#include <Eigen/Geometry>
namespace pcl {
// pcl/common/impl/transforms.hpp (1.8.1): xyz = M(0..2,0..2)*p + M(0..2,3) written out per coefficient in Scalar, cast to float;
// all other fields copied.
template <typename PointT, typename Scalar>
void transformPointCloud(const pcl::PointCloud<PointT>& cloud_in, pcl::PointCloud<PointT>& cloud_out, const Eigen::Transform<Scalar, 3, Eigen::Affine>& transform,
                         bool copy_all_fields = true) {
  if (&cloud_in != &cloud_out) {
    cloud_out.header = cloud_in.header;
    cloud_out.is_dense = cloud_in.is_dense;
    cloud_out.width = cloud_in.width;
    cloud_out.height = cloud_in.height;
    cloud_out.points.reserve(cloud_in.points.size());
    if (copy_all_fields) cloud_out.points.assign(cloud_in.points.begin(), cloud_in.points.end());
    else cloud_out.points.resize(cloud_in.points.size());
  }
  for (size_t i = 0; i < cloud_out.points.size(); ++i) {
    if (!cloud_in.is_dense && (!std::isfinite(cloud_in.points[i].x) || !std::isfinite(cloud_in.points[i].y) || !std::isfinite(cloud_in.points[i].z))) continue;
    Eigen::Matrix<Scalar, 3, 1> pt(cloud_in[i].x, cloud_in[i].y, cloud_in[i].z);
    cloud_out[i].x = static_cast<float>(transform(0, 0) * pt.coeffRef(0) + transform(0, 1) * pt.coeffRef(1) + transform(0, 2) * pt.coeffRef(2) + transform(0, 3));
    cloud_out[i].y = static_cast<float>(transform(1, 0) * pt.coeffRef(0) + transform(1, 1) * pt.coeffRef(1) + transform(1, 2) * pt.coeffRef(2) + transform(1, 3));
    cloud_out[i].z = static_cast<float>(transform(2, 0) * pt.coeffRef(0) + transform(2, 1) * pt.coeffRef(1) + transform(2, 2) * pt.coeffRef(2) + transform(2, 3));
  }
}
template <typename PointT>
void transformPointCloud(const pcl::PointCloud<PointT>& cloud_in, pcl::PointCloud<PointT>& cloud_out, const Eigen::Affine3f& transform, bool copy_all_fields = true) {
  return (transformPointCloud<PointT, float>(cloud_in, cloud_out, transform, copy_all_fields));
}
}  // namespace pcl
