// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl/pcl_macros.h> (PCL 1.8.1, un-vendored dependency of the reference).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <stdarg.h>
#include <stdio.h>
// PCL 1.8.1's pcl_macros.h includes the C header <math.h> (after `#define _USE_MATH_DEFINES`).  With libstdc++ that header is
// the C++ wrapper that pulls std::sqrt's float/long double overloads into the global namespace, which decides what the
// unqualified `sqrt(x*x + y*y)` on float operands at /root/reference/src/laserProcessingClass.cpp:14 resolves to (float overload).
#define _USE_MATH_DEFINES
#include <math.h>
#include <Eigen/Core>
#include <boost/shared_ptr.hpp>

#define PCL_ADD_UNION_POINT4D \
  union EIGEN_ALIGN16 {       \
    float data[4];            \
    struct {                  \
      float x;                \
      float y;                \
      float z;                \
    };                        \
  };
#define PCL_ADD_POINT4D PCL_ADD_UNION_POINT4D
// the field list (a Boost.PP sequence in real PCL) only feeds the PointCloud2 <-> PointT field mapping: not needed here
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fseq)
#define PCL_EXPORTS
