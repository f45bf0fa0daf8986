// Drop-in for include/laserMappingClass.h: LaserMappingClass::{init, updateCurrentPointsToMap, getMap} (src/laserMappingClass.cpp).
#ifndef FLOAM_B200_HOST_LASER_MAPPING_CLASS_H_
#define FLOAM_B200_HOST_LASER_MAPPING_CLASS_H_
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>
#ifdef FLOAM_B200_WITH_PCL   // include/laserMappingClass.h:8-23
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/filters/filter.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/filters/passthrough.h>
#include <pcl_ros/impl/transforms.hpp>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <string>
#include <math.h>
#include <vector>
#endif
#include <lidar.h>       // through the include path (this directory first), so that lidar.h's #include_next finds the reference's

class LaserMappingClass {
 public:
  LaserMappingClass() : owned_(new floam_b200_host::FloamContext()), fc_(owned_.get()) {}
  explicit LaserMappingClass(floam_b200_host::FloamContext* shared) : fc_(shared) {}
  void init(double map_resolution) {
    fc_->prm.map_resolution = map_resolution;
    floam_b200_host::report(fc_->ensure(), "LaserMappingClass::init");
  }
  void updateCurrentPointsToMap(const pcl::PointCloud<pcl::PointXYZI>::Ptr& pc_in, const Eigen::Isometry3d& pose_current) {
    if (fc_->ensure()) return;
    double T[16];
    floam_b200_host::isometry_to_rowmajor(pose_current, T);
    floam_b200_host::report(floam_mapping_update(fc_->ctx, reinterpret_cast<const floam_point_xyzi*>(pc_in->points.data()), (int)pc_in->points.size(), T),
                            "LaserMappingClass::updateCurrentPointsToMap");
  }
  // getMap() (:188-200) is called after every frame by the node (src/laserMappingNode.cpp:87) and the whole map would cross PCIe every
  // time.  Only the 50 m cells that changed since the last call are fetched (floam_mapping_get_changed_cells); the host keeps one
  // cloud per cell and concatenates them in the reference's (x, y, z) loop order: the same cloud, a bounded download per frame.
  pcl::PointCloud<pcl::PointXYZI>::Ptr getMap(void) {
    pcl::PointCloud<pcl::PointXYZI>::Ptr laserCloudMap(new pcl::PointCloud<pcl::PointXYZI>());
    if (fc_->ensure()) return laserCloudMap;
    int n = 0;
    if (floam_mapping_get_changed_cells(fc_->ctx, nullptr, nullptr, 0, &n) != FLOAM_OK) return laserCloudMap;
    if (n > 0) {
      std::vector<floam_point_xyzi> pts((std::size_t)n);
      std::vector<std::int32_t> cells((std::size_t)3 * n);
      const int rc = floam_mapping_get_changed_cells(fc_->ctx, pts.data(), cells.data(), n, &n);
      floam_b200_host::report(rc, "LaserMappingClass::getMap");
      if (rc != FLOAM_OK) return laserCloudMap;
      std::map<CellKey, std::vector<floam_point_xyzi>> fresh;
      for (int i = 0; i < n; ++i) fresh[CellKey{cells[3 * i], cells[3 * i + 1], cells[3 * i + 2]}].push_back(pts[i]);
      for (auto& kv : fresh) cells_[kv.first].swap(kv.second);
    }
    std::size_t total = 0;
    for (const auto& kv : cells_) total += kv.second.size();
    laserCloudMap->points.resize(total);
    std::size_t at = 0;
    for (const auto& kv : cells_) {   // std::map order of (x, y, z) = the reference's triple loop
      if (!kv.second.empty()) std::memcpy(static_cast<void*>(laserCloudMap->points.data() + at), kv.second.data(), kv.second.size() * sizeof(floam_point_xyzi));
      at += kv.second.size();
    }
    laserCloudMap->width = (std::uint32_t)total; laserCloudMap->height = 1;
    return laserCloudMap;
  }
  // the whole map straight from the device (what getMap() did before it became incremental); both return identical clouds
  pcl::PointCloud<pcl::PointXYZI>::Ptr getMapFull(void) {
    pcl::PointCloud<pcl::PointXYZI>::Ptr laserCloudMap(new pcl::PointCloud<pcl::PointXYZI>());
    if (fc_->ensure()) return laserCloudMap;
    int n = 0;
    if (floam_mapping_get_map(fc_->ctx, nullptr, 0, &n) != FLOAM_OK) return laserCloudMap;
    laserCloudMap->points.resize(n);
    floam_b200_host::report(floam_mapping_get_map(fc_->ctx, reinterpret_cast<floam_point_xyzi*>(laserCloudMap->points.data()), n, &n), "LaserMappingClass::getMapFull");
    laserCloudMap->width = n; laserCloudMap->height = 1;
    return laserCloudMap;
  }

 private:
  struct CellKey {
    std::int32_t x, y, z;
    bool operator<(const CellKey& o) const { return x != o.x ? x < o.x : (y != o.y ? y < o.y : z < o.z); }
  };
  std::map<CellKey, std::vector<floam_point_xyzi>> cells_;   // host copy of the map, one cloud per 50 m cell
  std::unique_ptr<floam_b200_host::FloamContext> owned_;
  floam_b200_host::FloamContext* fc_;
};
#endif
