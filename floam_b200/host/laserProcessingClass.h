// Drop-in for include/laserProcessingClass.h: LaserProcessingClass::{init, featureExtraction} (src/laserProcessingClass.cpp:6,72-118)
// on the CUDA path.  featureExtraction appends to the caller's clouds like the reference.
#ifndef FLOAM_B200_HOST_LASER_PROCESSING_CLASS_H_
#define FLOAM_B200_HOST_LASER_PROCESSING_CLASS_H_
#include <cstdio>
#ifdef FLOAM_B200_WITH_PCL   // what the reference's header includes (include/laserProcessingClass.h:6-16); the node relies on them
#define PCL_NO_PRECOMPILE
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/filters/filter.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/filters/passthrough.h>
#include <pcl/filters/extract_indices.h>
#include <pcl/filters/crop_box.h>
#endif
#include <lidar.h>       // through the include path (this directory first), so that lidar.h's #include_next finds the reference's

class LaserProcessingClass {
 public:
  LaserProcessingClass() : owned_(new floam_b200_host::FloamContext()), fc_(owned_.get()) {}
  explicit LaserProcessingClass(floam_b200_host::FloamContext* shared) : fc_(shared) {}
  void init(lidar::Lidar lidar_param_in) {
    lidar_param = lidar_param_in;
    fc_->set_lidar(lidar_param_in);
  }
  void featureExtraction(const pcl::PointCloud<vel_point::PointXYZIRT>::Ptr& pc_in, pcl::PointCloud<vel_point::PointXYZIRT>::Ptr& pc_out_edge,
                         pcl::PointCloud<vel_point::PointXYZIRT>::Ptr& pc_out_surf) {
    int rc = fc_->ensure();
    if (rc) return floam_b200_host::report(rc, "LaserProcessingClass::featureExtraction");
    const int n = (int)pc_in->points.size();
    const std::size_t e0 = pc_out_edge->points.size(), s0 = pc_out_surf->points.size();
    pc_out_edge->points.resize(e0 + n);
    pc_out_surf->points.resize(s0 + n);
    int ne = 0, ns = 0;
    rc = floam_feature_extract(fc_->ctx, reinterpret_cast<const floam_point_xyzirt*>(pc_in->points.data()), n,
                               reinterpret_cast<floam_point_xyzirt*>(pc_out_edge->points.data() + e0), n, &ne,
                               reinterpret_cast<floam_point_xyzirt*>(pc_out_surf->points.data() + s0), n, &ns);
    if (rc) { ne = 0; ns = 0; floam_b200_host::report(rc, "LaserProcessingClass::featureExtraction"); }
    pc_out_edge->points.resize(e0 + ne);
    pc_out_surf->points.resize(s0 + ns);
    pc_out_edge->width = (std::uint32_t)pc_out_edge->points.size(); pc_out_edge->height = 1;
    pc_out_surf->width = (std::uint32_t)pc_out_surf->points.size(); pc_out_surf->height = 1;
  }
  floam_b200_host::FloamContext* context() { return fc_; }

 private:
  lidar::Lidar lidar_param;
  std::unique_ptr<floam_b200_host::FloamContext> owned_;
  floam_b200_host::FloamContext* fc_;
};
#endif
