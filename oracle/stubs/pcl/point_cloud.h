// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl/point_cloud.h> (PCL 1.8.1): the container members the reference uses.
#pragma once
#include <pcl/pcl_macros.h>
#include <string>
#include <vector>
namespace pcl {
struct PCLHeader {  // pcl/PCLHeader.h
  PCLHeader() : seq(0), stamp(0) {}
  std::uint32_t seq;
  std::uint64_t stamp;  // microseconds since epoch
  std::string frame_id;
};
template <typename PointT>
class PointCloud {
 public:
  typedef PointT PointType;
  typedef std::vector<PointT, Eigen::aligned_allocator<PointT> > VectorType;
  typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT> > ConstPtr;
  typedef typename VectorType::iterator iterator;
  typedef typename VectorType::const_iterator const_iterator;
  PointCloud() : width(0), height(0), is_dense(true) {}
  inline PointCloud& operator+=(const PointCloud& rhs) {  // pcl/point_cloud.h (1.8.1)
    if (rhs.header.stamp > header.stamp) header.stamp = rhs.header.stamp;
    size_t nr_points = points.size();
    points.resize(nr_points + rhs.points.size());
    for (size_t i = nr_points; i < points.size(); ++i) points[i] = rhs.points[i - nr_points];
    width = static_cast<std::uint32_t>(points.size());
    height = 1;
    if (rhs.is_dense && is_dense) is_dense = true;
    else is_dense = false;
    return *this;
  }
  inline const PointCloud operator+(const PointCloud& rhs) { return (PointCloud(*this) += rhs); }
  PCLHeader header;
  VectorType points;
  std::uint32_t width, height;
  bool is_dense;
  inline iterator begin() { return points.begin(); }
  inline iterator end() { return points.end(); }
  inline const_iterator begin() const { return points.begin(); }
  inline const_iterator end() const { return points.end(); }
  inline size_t size() const { return points.size(); }
  inline void reserve(size_t n) { points.reserve(n); }
  inline bool empty() const { return points.empty(); }
  inline void resize(size_t n) {
    points.resize(n);
    if (width * height != n) { width = static_cast<std::uint32_t>(n); height = 1; }
  }
  inline const PointT& operator[](size_t n) const { return points[n]; }
  inline PointT& operator[](size_t n) { return points[n]; }
  inline const PointT& at(size_t n) const { return points.at(n); }
  inline PointT& at(size_t n) { return points.at(n); }
  inline const PointT& front() const { return points.front(); }
  inline PointT& front() { return points.front(); }
  inline const PointT& back() const { return points.back(); }
  inline PointT& back() { return points.back(); }
  inline void push_back(const PointT& pt) {
    points.push_back(pt);
    width = static_cast<std::uint32_t>(points.size());
    height = 1;
  }
  inline void clear() { points.clear(); width = 0; height = 0; }
  inline Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};
}  // namespace pcl
