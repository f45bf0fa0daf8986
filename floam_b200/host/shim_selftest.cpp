// Compile/link check of the reference-API shims over the C ABI, plus (when a B200 is present) a tiny end-to-end pass written the
// way the reference's node code calls the classes (src/laserProcessingNode.cpp:129, src/odomEstimationNode.cpp:219-228,
// src/laserMappingNode.cpp:85-87).  Exit code 0 = ok, 77 = no device (skipped).
#include <cmath>
#include <cstdio>
#include "dataHandler.h"
#include "laserMappingClass.h"
#include "laserProcessingClass.h"
#include "odomEstimationClass.h"
#include "pointCloud2Adapter.h"
#include <cstring>

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

// the message the Velodyne driver would publish for a scan: XYZIRT, 22 bytes per point
static sensor_msgs::PointCloud2 to_msg(const pcl::PointCloud<vel_point::PointXYZIRT>& c, double stamp) {
  sensor_msgs::PointCloud2 m;
  m.header.stamp.t = stamp;
  m.width = (std::uint32_t)c.size(); m.height = 1; m.point_step = 22; m.row_step = m.width * 22; m.is_bigendian = false;
  const char* names[6] = {"x", "y", "z", "intensity", "ring", "time"};
  const std::uint32_t offs[6] = {0, 4, 8, 12, 16, 18};
  for (int k = 0; k < 6; ++k) {
    sensor_msgs::PointField f;
    f.name = names[k]; f.offset = offs[k]; f.datatype = k == 4 ? sensor_msgs::PointField::UINT16 : sensor_msgs::PointField::FLOAT32;
    m.fields.push_back(f);
  }
  m.data.resize((size_t)m.row_step);
  for (size_t i = 0; i < c.size(); ++i) {
    std::uint8_t* p = m.data.data() + i * 22;
    std::memcpy(p, &c.points[i].x, 4); std::memcpy(p + 4, &c.points[i].y, 4); std::memcpy(p + 8, &c.points[i].z, 4);
    std::memcpy(p + 12, &c.points[i].intensity, 4); std::memcpy(p + 16, &c.points[i].ring, 2); std::memcpy(p + 18, &c.points[i].time, 4);
  }
  return m;
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
  bool is_odom_inited = false, incremental_ok = true;
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
    if (f != 2) {   // the node's per-frame getMap(): only changed cells cross PCIe; one frame skipped so that a hand-out spans two updates
      pcl::PointCloud<pcl::PointXYZI>::Ptr inc = laserMapping.getMap(), full = laserMapping.getMapFull();
      incremental_ok = incremental_ok && inc->points.size() == full->points.size() && !full->points.empty() &&
                       std::memcmp(inc->points.data(), full->points.data(), full->points.size() * sizeof(pcl::PointXYZI)) == 0;
    }
    std::printf("frame %d edge %zu surf %zu  t = %.4f %.4f %.4f  |v| = %.3f\n", f, pointcloud_edge->size(), pointcloud_surf->size(),
                odomEstimation.odom.translation().x(), odomEstimation.odom.translation().y(), odomEstimation.odom.translation().z(),
                odomEstimation.GetVelocity().norm());
  }
  pcl::PointCloud<pcl::PointXYZI>::Ptr local(new pcl::PointCloud<pcl::PointXYZI>());
  odomEstimation.getMap(local);
  pcl::PointCloud<pcl::PointXYZI>::Ptr global = laserMapping.getMap();
  std::printf("local map %zu points, global map %zu points\n", local->size(), global->size());
  const double x = odomEstimation.odom.translation().x();
  std::printf("incremental getMap identical to the full download on every frame: %s\n", incremental_ok ? "yes" : "NO");
  if (!(std::fabs(x - 0.2) < 0.05) || local->size() == 0 || global->size() == 0 || !incremental_ok) { std::printf("shim_selftest: FAILED\n"); return 1; }
  // the same five scans as PointCloud2 messages through the fused node adapter (own context): same kernels, same pose
  floam_b200_host::FloamContext fused;
  fused.prm = shared.prm;
  fused.set_lidar(lidar_param);
  fused.ensure();
  if (!fused.ctx) { std::printf("shim_selftest: FAILED (second context)\n"); return 1; }
  floam_b200_host::FusedOdometryNode node(fused.ctx, false, false, Eigen::Quaterniond(1, 0, 0, 0));
  std::vector<sensor_msgs::PointCloud2> msgs;
  for (int f = 0; f < 5; ++f) msgs.push_back(to_msg(*make_scan(0.05 * f), 0.1 * f));
  double pose[7] = {0, 0, 0, 1, 0, 0, 0};
  bool have = false;
  for (int f = 0; f < 5; ++f)
    if (node.velodyneHandler(msgs[f], pose, &have) != FLOAM_OK) { std::printf("shim_selftest: FAILED (fused submit %d)\n", f); return 1; }
  while (node.flush(pose) == FLOAM_OK) {}
  std::printf("fused PointCloud2 path: t = %.4f %.4f %.4f\n", pose[4], pose[5], pose[6]);
  if (!(std::fabs(pose[4] - x) < 1e-9)) { std::printf("shim_selftest: FAILED (fused path differs: %.17g vs %.17g)\n", pose[4], x); return 1; }
  std::printf("shim_selftest: ok\n");
  return 0;
}
