// Compile/link check of the reference-API shims over the C ABI, plus (when a B200 is present) a tiny end-to-end pass written the
// way the reference's node code calls the classes (src/laserProcessingNode.cpp:129, src/odomEstimationNode.cpp:219-228,
// src/laserMappingNode.cpp:85-87).  Exit code 0 = ok, 77 = no device (skipped).
#include <cmath>
#include <cstdio>
#include "dataHandler.h"
#include "laserMappingClass.h"
#include "laserProcessingClass.h"
#include "odomEstimationClass.h"

static pcl::PointCloud<vel_point::PointXYZIRT>::Ptr make_scan(double shift) {
  pcl::PointCloud<vel_point::PointXYZIRT>::Ptr c(new pcl::PointCloud<vel_point::PointXYZIRT>());
  for (int a = 0; a < 900; ++a)
    for (int r = 0; r < 16; ++r) {
      vel_point::PointXYZIRT p;
      const double az = -M_PI + 2 * M_PI * a / 900.0, el = (-15.0 + 2.0 * r) * M_PI / 180.0;
      // a 20 m x 12 m room, sensor at (shift, 0): range to the nearest wall along the ray
      const double cx = std::cos(az), sy = std::sin(az);
      const double tx = cx > 0 ? (10.0 - shift) / cx : (-10.0 - shift) / cx, ty = sy > 0 ? 6.0 / sy : -6.0 / sy;
      const double rng = std::fmin(std::fabs(tx), std::fabs(ty)) * (1.0 + 0.001 * std::sin(37.0 * a + r));
      p.x = (float)(rng * cx); p.y = (float)(rng * sy); p.z = (float)(rng * std::tan(el));
      p.intensity = 0.5f; p.ring = (std::uint16_t)r; p.time = (float)(0.1 * a / 900.0);
      c->push_back(p);
    }
  return c;
}

int main() {
  lidar::Lidar lidar_param;
  lidar_param.setLines(16); lidar_param.setScanPeriod(0.1); lidar_param.setMaxDistance(60.0); lidar_param.setMinDistance(0.5);
  floam_b200_host::FloamContext shared;          // features, odometry and mapping share one device context
  shared.prm.max_scan_points = 20000; shared.prm.max_map_points = 1 << 18; shared.prm.max_global_map_points = 1 << 18; shared.prm.max_grid_cells = 1 << 20;
  shared.set_lidar(lidar_param);
  LaserProcessingClass laserProcessing(&shared);
  OdomEstimationClass odomEstimation(&shared);
  LaserMappingClass laserMapping(&shared);
  laserProcessing.init(lidar_param);
  odomEstimation.init(lidar_param, 0.4, "Cauchy");
  if (!shared.ctx) { std::printf("shim_selftest: no CUDA device, compile/link check only\n"); return 77; }
  laserMapping.init(0.4);
  bool is_odom_inited = false;
  for (int f = 0; f < 5; ++f) {
    pcl::PointCloud<vel_point::PointXYZIRT>::Ptr pointcloud_in = make_scan(0.05 * f);
    pcl::PointCloud<vel_point::PointXYZIRT>::Ptr pointcloud_edge(new pcl::PointCloud<vel_point::PointXYZIRT>());
    pcl::PointCloud<vel_point::PointXYZIRT>::Ptr pointcloud_surf(new pcl::PointCloud<vel_point::PointXYZIRT>());
    laserProcessing.featureExtraction(pointcloud_in, pointcloud_edge, pointcloud_surf);
    if (!is_odom_inited) {
      odomEstimation.initMapWithPoints(VelToIntensityCopy(pointcloud_edge), VelToIntensityCopy(pointcloud_surf));
      is_odom_inited = true;
    } else {
      odomEstimation.UpdatePointsToMapSelector(pointcloud_edge, pointcloud_surf, false);
    }
    laserMapping.updateCurrentPointsToMap(VelToIntensityCopy(pointcloud_surf), odomEstimation.odom);
    std::printf("frame %d edge %zu surf %zu  t = %.4f %.4f %.4f  |v| = %.3f\n", f, pointcloud_edge->size(), pointcloud_surf->size(),
                odomEstimation.odom.translation().x(), odomEstimation.odom.translation().y(), odomEstimation.odom.translation().z(),
                odomEstimation.GetVelocity().norm());
  }
  pcl::PointCloud<pcl::PointXYZI>::Ptr local(new pcl::PointCloud<pcl::PointXYZI>());
  odomEstimation.getMap(local);
  pcl::PointCloud<pcl::PointXYZI>::Ptr global = laserMapping.getMap();
  std::printf("local map %zu points, global map %zu points\n", local->size(), global->size());
  const double x = odomEstimation.odom.translation().x();
  if (!(std::fabs(x - 0.2) < 0.05) || local->size() == 0 || global->size() == 0) { std::printf("shim_selftest: FAILED\n"); return 1; }
  std::printf("shim_selftest: ok\n");
  return 0;
}
