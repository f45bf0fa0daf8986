// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl/kdtree/kdtree_flann.h>.  The tree is oracle/kdtree.cpp: FLANN 1.9.1's
// KDTreeSingleIndex as pcl::KdTreeFLANN builds it (pinned against the FLANN copy bundled with OpenCV, tests/test_oracle.py).
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <vector>
#include "../../../floam_oracle.h"
namespace pcl {
namespace floam_stub {
// test knob (not in PCL): brute-force search with (distance, index) tie order instead of the kd-tree traversal order
inline bool& knn_bruteforce() { static bool v = false; return v; }
inline long& knn_query_count() { static long v = 0; return v; }
}  // namespace floam_stub
template <typename PointT>
class KdTreeFLANN {
 public:
  typedef boost::shared_ptr<KdTreeFLANN<PointT> > Ptr;
  typedef boost::shared_ptr<const KdTreeFLANN<PointT> > ConstPtr;
  typedef typename pcl::PointCloud<PointT>::ConstPtr PointCloudConstPtr;
  KdTreeFLANN(bool sorted = true) { (void)sorted; }
  void setInputCloud(const PointCloudConstPtr& cloud) {
    input_ = cloud;
    if (!floam_stub::knn_bruteforce()) tree_.setInputCloud(reinterpret_cast<const fo::PointXYZI*>(cloud->points.data()), cloud->points.size());
  }
  int nearestKSearch(const PointT& point, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    const int total = static_cast<int>(input_->points.size());
    if (k > total) k = total;
    k_indices.resize(k);
    k_sqr_distances.resize(k);
    floam_stub::knn_query_count()++;
    if (k == 0) return 0;
    fo::PointXYZI q;
    q.x = point.x; q.y = point.y; q.z = point.z;
    if (floam_stub::knn_bruteforce())
      return fo::knn_bruteforce(reinterpret_cast<const fo::PointXYZI*>(input_->points.data()), (size_t)total, q, k, k_indices.data(), k_sqr_distances.data());
    return tree_.nearestKSearch(q, k, k_indices.data(), k_sqr_distances.data());
  }
 private:
  PointCloudConstPtr input_;
  fo::KdTreeFlann tree_;
};
}  // namespace pcl
