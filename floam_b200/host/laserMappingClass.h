// Drop-in for include/laserMappingClass.h: LaserMappingClass::{init, updateCurrentPointsToMap, getMap} (src/laserMappingClass.cpp).
#ifndef FLOAM_B200_HOST_LASER_MAPPING_CLASS_H_
#define FLOAM_B200_HOST_LASER_MAPPING_CLASS_H_
#include <cstdio>
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
  pcl::PointCloud<pcl::PointXYZI>::Ptr getMap(void) {
    pcl::PointCloud<pcl::PointXYZI>::Ptr laserCloudMap(new pcl::PointCloud<pcl::PointXYZI>());
    if (fc_->ensure()) return laserCloudMap;
    int n = 0;
    if (floam_mapping_get_map(fc_->ctx, nullptr, 0, &n) != FLOAM_OK) return laserCloudMap;
    laserCloudMap->points.resize(n);
    floam_b200_host::report(floam_mapping_get_map(fc_->ctx, reinterpret_cast<floam_point_xyzi*>(laserCloudMap->points.data()), n, &n), "LaserMappingClass::getMap");
    laserCloudMap->width = n; laserCloudMap->height = 1;
    return laserCloudMap;
  }

 private:
  std::unique_ptr<floam_b200_host::FloamContext> owned_;
  floam_b200_host::FloamContext* fc_;
};
#endif
