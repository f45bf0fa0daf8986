// ORACLE — TEST INFRASTRUCTURE ONLY (see types.h).  Pinned: tests/test_reference_pin.py runs this restatement against the reference's own
// sources compiled unmodified (oracle/_ref, `make ref`) on identical inputs — identical selections, bytes and poses.
// Restates src/dataHandler.cpp:24-122 (ImuHandler, CompensateVelocity, Compensate) and the IMU folding done by the
// caller, src/laserProcessingNode.cpp:65-78 (CenterTime) and :113-116 (IMU alignment via pcl::transformPointCloud).
#include "floam_oracle.h"
#include <algorithm>
#include <limits>

namespace fo {

// ros::Time(double) / toSec() / pcl_conversions round trip: stamp held as microseconds in pcl headers.
static inline double stamp_to_sec(std::uint64_t stamp_us) {
  // pcl_conversions::fromPCL: nsec = stamp*1000 ; ros::Time::toSec() = sec + 1e-9*nsec
  std::uint64_t ns = stamp_us * 1000ull;
  return (double)(ns / 1000000000ull) + 1e-9 * (double)(ns % 1000000000ull);
}
static inline std::uint64_t sec_to_stamp(double t) {
  // ros::Time(double t): sec = floor(t), nsec = round((t-sec)*1e9), normalised ; toPCL: nsec/1000 (truncating) + sec*1e6
  std::uint64_t sec = (std::uint64_t)std::floor(t);
  std::uint64_t nsec = (std::uint64_t)std::llround((t - (double)sec) * 1e9);
  sec += nsec / 1000000000ull;
  nsec %= 1000000000ull;
  return nsec / 1000ull + sec * 1000000ull;
}

void ImuHandler::AddMsg(double stamp, const Quat& orientation) {
  if (data_.empty()) { data_.push_back(std::make_pair(stamp, orientation)); return; }
  const double tdiff = stamp - data_.back().first;
  if (tdiff > 0.00001) data_.push_back(std::make_pair(stamp, orientation));
}

bool ImuHandler::Get(double tStamp, Quat& data) const {
  auto first = data_.begin();
  auto last = data_.end();
  auto itr_after = std::lower_bound(first, last, tStamp, [](const std::pair<double, Quat>& a, double t) { return a.first < t; });
  if (itr_after == first) return false;  // std::prev(first) is undefined in the reference; the condition below rejects it anyway
  auto itr_before = std::prev(itr_after, 1);
  if (itr_after != last && itr_after != first && itr_before != first) {
    data = itr_before->second;  // Interpolate() returns data1 (zero-order hold), :48-50
    if (slerp) {  // opt-in fix, not the reference: Eigen's QuaternionBase::slerp at the tSlerp the reference computes (:61)
      const double t = (tStamp - itr_before->first) / (itr_after->first - itr_before->first);
      const Quat &a = itr_before->second, &b = itr_after->second;
      const double one = 1.0 - std::numeric_limits<double>::epsilon();
      const double d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w, absD = std::fabs(d);
      double scale0, scale1;
      if (absD >= one) { scale0 = 1.0 - t; scale1 = t; }
      else { const double theta = std::acos(absD), sinTheta = std::sin(theta); scale0 = std::sin((1.0 - t) * theta) / sinTheta; scale1 = std::sin(t * theta) / sinTheta; }
      if (d < 0) scale1 = -scale1;
      data = Quat{scale0 * a.x + scale1 * b.x, scale0 * a.y + scale1 * b.y, scale0 * a.z + scale1 * b.z, scale0 * a.w + scale1 * b.w};
    }
    return true;
  }
  return false;
}

Quat ImuHandler::Get(double tStamp) const {
  Quat data{0, 0, 0, 0};  // default sensor_msgs::Imu: all-zero orientation
  Get(tStamp, data);
  return data;
}

bool ImuHandler::TimeContained(double t) const {
  return !data_.empty() && t >= data_.front().first && t <= data_.back().first;
}

void CenterTime(CloudIRT& cloud, std::uint64_t& stamp_us) {
  if (cloud.empty()) return;
  const double tScan = stamp_to_sec(stamp_us);
  const double tEnd = tScan + cloud.back().time;
  const double tBegin = tScan + cloud.front().time;
  const double tCenter = tBegin + (tEnd - tBegin) / 2.0;
  stamp_us = sec_to_stamp(tCenter);
  for (PointXYZIRT& pnt : cloud) pnt.time = pnt.time + tScan - tCenter;  // float + double - double -> float store
}

bool Compensate(const CloudIRT& input, std::uint64_t stamp_us, CloudIRT& compensated, const ImuHandler& handler, const Quat& extrinsics) {
  compensated.assign(input.size(), PointXYZIRT{});
  if (input.empty()) return false;  // front()/back() on an empty cloud is undefined in the reference
  const double tScan = stamp_to_sec(stamp_us);
  const double t0 = input.front().time + tScan;
  const double t1 = input.back().time + tScan;
  if (!handler.TimeContained(t0) || !handler.TimeContained(t1)) return false;
  const Quat qInit = quat_mul(handler.Get(tScan), extrinsics);
  const Quat qInitInv = quat_inverse(qInit);
  for (size_t i = 0; i < input.size(); i++) {
    const double timeCurrent = tScan + input[i].time;
    const Quat qNow = quat_mul(handler.Get(timeCurrent), extrinsics);
    const Quat qDiff = quat_mul(qInitInv, qNow);
    const Vec3 pT = quat_rotate(qDiff, Vec3{input[i].x, input[i].y, input[i].z});
    compensated[i].x = (float)pT.x; compensated[i].y = (float)pT.y; compensated[i].z = (float)pT.z;
    compensated[i]._pad0 = 1.0f;
    compensated[i].ring = input[i].ring; compensated[i].time = input[i].time; compensated[i].intensity = input[i].intensity;
  }
  return true;
}

void ImuAlign(const CloudIRT& compensated, std::uint64_t stamp_us, const ImuHandler& handler, const Quat& extrinsics, CloudIRT& aligned) {
  const Quat q = quat_mul(handler.Get(stamp_to_sec(stamp_us)), extrinsics);
  const Mat3 R = quat_to_matrix(q);  // Eigen::Affine3d ImuNowT(q)
  aligned = compensated;
  for (size_t i = 0; i < compensated.size(); ++i) {
    // pcl::transformPointCloud(Affine3d): double R*p (+0), cast to float
    const double x = compensated[i].x, y = compensated[i].y, z = compensated[i].z;
    aligned[i].x = static_cast<float>(R.m[0][0] * x + R.m[0][1] * y + R.m[0][2] * z + 0.0);
    aligned[i].y = static_cast<float>(R.m[1][0] * x + R.m[1][1] * y + R.m[1][2] * z + 0.0);
    aligned[i].z = static_cast<float>(R.m[2][0] * x + R.m[2][1] * y + R.m[2][2] * z + 0.0);
  }
}

void CompensateVelocity(CloudIRT& input, Vec3 velocity) {
  for (PointXYZIRT& pnt : input) {
    const double tPoint = pnt.time;
    const Vec3 pntPosition{pnt.x, pnt.y, pnt.z};
    const Vec3 pntError = tPoint * velocity;
    const Vec3 c = pntPosition + pntError;
    pnt.x = (float)c.x; pnt.y = (float)c.y; pnt.z = (float)c.z;
  }
}

void CompensateVelocityRotated(CloudIRT& input, Vec3 velocity, const Mat3& R) {
  const Vec3 v = mat3_apply(mat3_transpose(R), velocity);
  CompensateVelocity(input, v);
}

}  // namespace fo
