// Drop-in for the reference's include/lidar.h.
//  * Inside the catkin workspace (-DFLOAM_B200_WITH_PCL, this directory FIRST on the include path): the reference's own lidar.h is
//    pulled in through #include_next, so vel_point::PointXYZIRT (:14-32), lidar::Lidar (:53-86), PublishCloud (:36-49) and
//    GetParamFromRos stay the reference's (and src/lidar.cpp keeps providing euler2Quaternion and the setters); this header only
//    adds the context plumbing the shims share.  tests/test_abi.py parses the three unmodified node sources against this mode.
//  * Stand-alone (this repo's build container: no PCL / Eigen / ROS): the same types are defined here.
#ifndef FLOAM_B200_HOST_LIDAR_H_
#define FLOAM_B200_HOST_LIDAR_H_
#include "floam_b200.h"
#include "floam_compat.h"

#ifdef FLOAM_B200_WITH_PCL
#include_next <lidar.h>
#else
namespace vel_point {
struct alignas(16) PointXYZIRT {
  float x = 0, y = 0, z = 0, data3 = 1.0f;
  float intensity = 0;
  std::uint16_t ring = 0;
  float time = 0;
};
}  // namespace vel_point

namespace lidar {
class Lidar {  // include/lidar.h:53-86, src/lidar.cpp:18-50
 public:
  Lidar() {}
  void setScanPeriod(double v) { scan_period = v; }
  void setLines(double v) { num_lines = (int)v; }
  void setVerticalAngle(double v) { vertical_angle = v; }
  void setVerticalResolution(double v) { vertical_angle_resolution = v; }
  void setMaxDistance(double v) { max_distance = v; }
  void setMinDistance(double v) { min_distance = v; }
  double max_distance = 60.0, min_distance = 2.0;
  int num_lines = 64;
  double scan_period = 0.1;
  int points_per_line = 0;
  double horizontal_angle_resolution = 0, horizontal_angle = 0, vertical_angle_resolution = 0, vertical_angle = 2.0;
};
}  // namespace lidar

inline Eigen::Quaterniond euler2Quaternion(const double roll, const double pitch, const double yaw) {  // degrees; q = roll * yaw * pitch
  const double d = M_PI / 180.0;
  const double r[4] = {std::sin(0.5 * roll * d), 0, 0, std::cos(0.5 * roll * d)};
  const double p[4] = {0, std::sin(0.5 * pitch * d), 0, std::cos(0.5 * pitch * d)};
  const double y[4] = {0, 0, std::sin(0.5 * yaw * d), std::cos(0.5 * yaw * d)};
  auto mul = [](const double* a, const double* b, double* o) {
    const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1], yy = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0], w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = yy; o[2] = z; o[3] = w;
  };
  double ry[4], q[4];
  mul(r, y, ry);
  mul(ry, p, q);
  return Eigen::Quaterniond(q[3], q[0], q[1], q[2]);
}
#endif  // FLOAM_B200_WITH_PCL

static_assert(sizeof(vel_point::PointXYZIRT) == sizeof(floam_point_xyzirt), "PointXYZIRT must be byte-identical to the C ABI point");
static_assert(sizeof(pcl::PointXYZI) == sizeof(floam_point_xyzi), "pcl::PointXYZI must be byte-identical to the C ABI point");

namespace floam_b200_host {
// One floam_ctx per class instance by default; FloamContext::share() lets LaserProcessingClass and OdomEstimationClass of one
// process use the same context so that features never leave the device.
struct FloamContext {
  floam_ctx* ctx = nullptr;
  floam_params prm;
  FloamContext() { floam_params_default(&prm); }
  ~FloamContext() { floam_destroy(ctx); }
  FloamContext(const FloamContext&) = delete;
  FloamContext& operator=(const FloamContext&) = delete;
  int ensure(int device = 0) { return ctx ? FLOAM_OK : floam_create(&prm, device, &ctx); }
  void set_lidar(const lidar::Lidar& l) {
    prm.num_lines = l.num_lines; prm.scan_period = l.scan_period; prm.vertical_angle = l.vertical_angle;
    prm.max_distance = l.max_distance; prm.min_distance = l.min_distance;
  }
};
inline void report(int status, const char* where) {  // the reference's classes have no error returns: failures are prints
  if (status != FLOAM_OK) std::fprintf(stderr, "[floam_b200] %s: %s\n", where, floam_status_string(status));
}
}  // namespace floam_b200_host
#endif
