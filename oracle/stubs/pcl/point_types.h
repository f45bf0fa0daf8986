// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl/point_types.h> (PCL 1.8.1): pcl::PointXYZI, 32 bytes, intensity @16.
#pragma once
#include <pcl/pcl_macros.h>
namespace pcl {
struct EIGEN_ALIGN16 _PointXYZI {
  PCL_ADD_POINT4D;
  union {
    struct {
      float intensity;
    };
    float data_c[4];
  };
};
struct PointXYZI : public _PointXYZI {  // pcl/impl/point_types.hpp: x=y=z=0, data[3]=1, intensity=0
  inline PointXYZI() { x = y = z = 0.0f; data[3] = 1.0f; intensity = 0.0f; }
  inline PointXYZI(float _intensity) { x = y = z = 0.0f; data[3] = 1.0f; intensity = _intensity; }
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes");
struct EIGEN_ALIGN16 PointXYZ {
  PCL_ADD_POINT4D;
  inline PointXYZ() { x = y = z = 0.0f; data[3] = 1.0f; }
};
}  // namespace pcl
