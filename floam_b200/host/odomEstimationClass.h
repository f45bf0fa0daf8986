// Drop-in for include/odomEstimationClass.h (src/odomEstimationClass.cpp:7-343): same public methods and members.  `odom`,
// `laserCloudCornerMap` and `laserCloudSurfMap` are public data members in the reference; here `odom` is refreshed after every
// update and the two map clouds are refreshed on demand by syncMaps() (the node only reads them in getMap / at exit).
#ifndef FLOAM_B200_HOST_ODOM_ESTIMATION_CLASS_H_
#define FLOAM_B200_HOST_ODOM_ESTIMATION_CLASS_H_
#include <cstdio>
#include <string>
#ifdef FLOAM_B200_WITH_PCL   // the includes of the reference's header the node depends on (include/odomEstimationClass.h:8-40): PCL filters and
#include <math.h>            // io for its dump-on-exit code, Dump / SavePosegraph / SaveOdom from utils.h; Ceres is no longer needed
#include <vector>
#include <sstream>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/filters/filter.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/filters/passthrough.h>
#include <pcl/filters/extract_indices.h>
#include <pcl/filters/crop_box.h>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <ros/ros.h>
#endif
#include <lidar.h>       // through the include path (this directory first), so that lidar.h's #include_next finds the reference's
#include <dataHandler.h>
#ifdef FLOAM_B200_WITH_PCL
#include "utils.h"
using std::cout;
using std::endl;
#endif

inline pcl::PointCloud<pcl::PointXYZI>::Ptr VelToIntensityCopy(const pcl::PointCloud<vel_point::PointXYZIRT>::Ptr VelCloud) {  // :308-318
  pcl::PointCloud<pcl::PointXYZI>::Ptr converted(new pcl::PointCloud<pcl::PointXYZI>());
  converted->points.resize(VelCloud->points.size());
  for (std::size_t i = 0; i < VelCloud->points.size(); i++) {
    converted->points[i].x = VelCloud->points[i].x; converted->points[i].y = VelCloud->points[i].y; converted->points[i].z = VelCloud->points[i].z;
    converted->points[i].intensity = VelCloud->points[i].intensity;
  }
  converted->width = (std::uint32_t)converted->points.size(); converted->height = 1;
  return converted;
}

class OdomEstimationClass {
 public:
  typedef enum { VANILLA, INITIAL_ITERATION, REFINEMENT_AND_UPDATE } UpdateType;

  OdomEstimationClass() : owned_(new floam_b200_host::FloamContext()), fc_(owned_.get()) { reset_members(); }
  explicit OdomEstimationClass(floam_b200_host::FloamContext* shared) : fc_(shared) { reset_members(); }

  void init(lidar::Lidar lidar_param, double map_resolution, const std::string& loss_function) {
    lidar_param_ = lidar_param;
    fc_->set_lidar(lidar_param);
    fc_->prm.map_resolution = map_resolution;
    fc_->prm.loss = floam_loss_from_string(loss_function.c_str());
    floam_b200_host::report(fc_->ensure(), "OdomEstimationClass::init");
  }
  void initMapWithPoints(const pcl::PointCloud<pcl::PointXYZI>::Ptr& edge_in, const pcl::PointCloud<pcl::PointXYZI>::Ptr& surf_in) {
    if (fc_->ensure()) return;
    floam_b200_host::report(floam_odom_init_map(fc_->ctx, reinterpret_cast<const floam_point_xyzi*>(edge_in->points.data()), (int)edge_in->points.size(),
                                                reinterpret_cast<const floam_point_xyzi*>(surf_in->points.data()), (int)surf_in->points.size()),
                            "OdomEstimationClass::initMapWithPoints");
  }
  // mutates edge_in / surf_in in deskew mode like the reference (:42-43)
  void UpdatePointsToMapSelector(pcl::PointCloud<vel_point::PointXYZIRT>::Ptr& edge_in, pcl::PointCloud<vel_point::PointXYZIRT>::Ptr& surf_in, bool deskew) {
    if (fc_->ensure()) return;
    double pose[7];
    floam_b200_host::report(floam_odom_update(fc_->ctx, reinterpret_cast<floam_point_xyzirt*>(edge_in->points.data()), (int)edge_in->points.size(),
                                              reinterpret_cast<floam_point_xyzirt*>(surf_in->points.data()), (int)surf_in->points.size(), deskew ? 1 : 0, pose),
                            "OdomEstimationClass::UpdatePointsToMapSelector");
    refresh_odom();
  }
  void updatePointsToMap(const pcl::PointCloud<pcl::PointXYZI>::Ptr& edge_in, const pcl::PointCloud<pcl::PointXYZI>::Ptr& surf_in,
                         const UpdateType update_type = UpdateType::VANILLA) {
    if (fc_->ensure()) return;
    double pose[7];
    floam_b200_host::report(floam_odom_update_xyzi(fc_->ctx, reinterpret_cast<const floam_point_xyzi*>(edge_in->points.data()), (int)edge_in->points.size(),
                                                   reinterpret_cast<const floam_point_xyzi*>(surf_in->points.data()), (int)surf_in->points.size(),
                                                   (int)update_type, pose),
                            "OdomEstimationClass::updatePointsToMap");
    refresh_odom();
  }
  void updatePointsToMap(const pcl::PointCloud<vel_point::PointXYZIRT>::Ptr& edge_in, const pcl::PointCloud<vel_point::PointXYZIRT>::Ptr& surf_in,
                         const UpdateType update_type = UpdateType::VANILLA) {
    updatePointsToMap(VelToIntensityCopy(edge_in), VelToIntensityCopy(surf_in), update_type);  // :52-56
  }
  void getMap(pcl::PointCloud<pcl::PointXYZI>::Ptr& laserCloudMap) {  // :296-300: surf then corner
    syncMaps();
    *laserCloudMap += *laserCloudSurfMap;
    *laserCloudMap += *laserCloudCornerMap;
  }
  Eigen::Vector3d GetVelocity() {  // include/odomEstimationClass.h:78
    double v[3] = {0, 0, 0};
    if (!fc_->ensure()) floam_odom_get(fc_->ctx, nullptr, v);
    return Eigen::Vector3d(v[0], v[1], v[2]);
  }
  // downloads the device-resident local maps into the public clouds
  void syncMaps() {
    if (fc_->ensure()) return;
    int ne = 0, ns = 0;
    floam_odom_map_sizes(fc_->ctx, &ne, &ns);
    laserCloudCornerMap->points.resize(ne); laserCloudSurfMap->points.resize(ns);
    floam_b200_host::report(floam_odom_get_map(fc_->ctx, reinterpret_cast<floam_point_xyzi*>(laserCloudCornerMap->points.data()), ne,
                                               reinterpret_cast<floam_point_xyzi*>(laserCloudSurfMap->points.data()), ns), "OdomEstimationClass::syncMaps");
    laserCloudCornerMap->width = ne; laserCloudCornerMap->height = 1; laserCloudSurfMap->width = ns; laserCloudSurfMap->height = 1;
  }
  floam_b200_host::FloamContext* context() { return fc_; }

  Eigen::Isometry3d odom;
  pcl::PointCloud<pcl::PointXYZI>::Ptr laserCloudCornerMap;
  pcl::PointCloud<pcl::PointXYZI>::Ptr laserCloudSurfMap;

 private:
  void reset_members() {
    odom = Eigen::Isometry3d::Identity();
    laserCloudCornerMap.reset(new pcl::PointCloud<pcl::PointXYZI>());
    laserCloudSurfMap.reset(new pcl::PointCloud<pcl::PointXYZI>());
  }
  void refresh_odom() {
    double T[16];
    if (floam_odom_get(fc_->ctx, T, nullptr) == FLOAM_OK) floam_b200_host::rowmajor_to_isometry(T, odom);
  }
  lidar::Lidar lidar_param_;
  std::unique_ptr<floam_b200_host::FloamContext> owned_;
  floam_b200_host::FloamContext* fc_;
};
#endif
