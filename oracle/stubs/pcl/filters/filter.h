// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl/filters/filter.h> (PCL 1.8.1): removeNaNFromPointCloud + Filter base.
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <boost/algorithm/string.hpp>  // the real PCL headers drag Boost.Algorithm in (src/odomEstimationClass.cpp:23 relies on it)
#include <vector>
namespace pcl {
template <typename PointT> void removeNaNFromPointCloud(const pcl::PointCloud<PointT>& cloud_in, pcl::PointCloud<PointT>& cloud_out, std::vector<int>& index);
template <typename PointT> void removeNaNFromPointCloud(const pcl::PointCloud<PointT>& cloud_in, std::vector<int>& index);

template <typename PointT>
class Filter {
 public:
  typedef pcl::PointCloud<PointT> PointCloud;
  typedef typename PointCloud::Ptr PointCloudPtr;
  typedef typename PointCloud::ConstPtr PointCloudConstPtr;
  virtual ~Filter() {}
  inline void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  inline void filter(PointCloud& output) { applyFilter(output); }  // applyFilter below tolerates &output == input_.get()
 protected:
  virtual void applyFilter(PointCloud& output) = 0;
  PointCloudConstPtr input_;
};
}  // namespace pcl
#include <pcl/filters/impl/filter.hpp>
