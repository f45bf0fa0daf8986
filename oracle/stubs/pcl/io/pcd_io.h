// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl/io/pcd_io.h>: savePCDFileBinary as PCL 1.8.1 writes XYZI clouds (PCD v0.7, binary).
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <cstdio>
#include <fstream>   // the real pcd_io.h brings <fstream> in (src/utils.cpp uses std::ofstream without including it)
#include <string>
namespace pcl {
namespace io {
template <typename PointT>
int savePCDFileBinary(const std::string& file_name, const pcl::PointCloud<PointT>& cloud) {
  std::FILE* f = std::fopen(file_name.c_str(), "wb");
  if (!f) return -1;
  std::fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n");
  std::fprintf(f, "WIDTH %u\nHEIGHT %u\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n", cloud.width, cloud.height, cloud.points.size());
  for (size_t i = 0; i < cloud.points.size(); ++i) {
    const float rec[4] = {cloud.points[i].x, cloud.points[i].y, cloud.points[i].z, cloud.points[i].intensity};
    std::fwrite(rec, sizeof(float), 4, f);
  }
  std::fclose(f);
  return 0;
}
}  // namespace io
}  // namespace pcl
