// ORACLE — TEST INFRASTRUCTURE ONLY.  The reference ships no golden vectors; this restatement is pinned against the
// reference's own class sources compiled unmodified into oracle/_ref (tests/test_reference_pin.py).  The PCL / FLANN / Eigen / Ceres
// internals it restates are un-vendored and absent from this image: for those, parity is unpinned.
//
// CPU restatement of the dan11003/floam per-frame odometry hot path, dependency-free C++17.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this
// directory.  The product (floam_b200/, include/) never includes, links or calls anything in here.
//
// Point layouts follow include/lidar.h:14-32 (vel_point::PointXYZIRT, 32 B, EIGEN_ALIGN16) and
// pcl::PointXYZI (32 B, intensity @16).
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <cmath>

namespace fo {

struct PointXYZIRT {  // include/lidar.h:14-22
  float x, y, z, _pad0;
  float intensity;
  std::uint16_t ring;
  std::uint16_t _pad1;
  float time;
  float _pad2;
};
static_assert(sizeof(PointXYZIRT) == 32, "PointXYZIRT must be 32 bytes");
static_assert(offsetof(PointXYZIRT, intensity) == 16, "intensity @16");
static_assert(offsetof(PointXYZIRT, ring) == 20, "ring @20");
static_assert(offsetof(PointXYZIRT, time) == 24, "time @24");

struct PointXYZI {  // pcl::PointXYZI
  float x, y, z, _pad0;
  float intensity;
  float _pad1[3];
};
static_assert(sizeof(PointXYZI) == 32, "PointXYZI must be 32 bytes");

// pcl::PointXYZI default constructor: x=y=z=0, data[3]=1, intensity=0.
inline PointXYZI make_xyzi(float x, float y, float z, float intensity) {
  PointXYZI p;
  p.x = x; p.y = y; p.z = z; p._pad0 = 1.0f;
  p.intensity = intensity; p._pad1[0] = p._pad1[1] = p._pad1[2] = 0.0f;
  return p;
}

// ---- tiny fixed-size linear algebra (double), standing in for the Eigen types the reference uses ----
struct Vec3 { double x, y, z; };
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator/(Vec3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(Vec3 a) { return std::sqrt(dot(a, a)); }

struct Quat { double x, y, z, w; };  // Eigen coefficient order (x,y,z,w), as in parameters[0..3]

struct Mat3 { double m[3][3]; };

struct Iso3 {  // Eigen::Isometry3d
  Mat3 R;
  Vec3 t;
};

}  // namespace fo
