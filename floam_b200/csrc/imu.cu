// IMU-aided deskew folded into the scan path: CenterTime (src/laserProcessingNode.cpp:65-78), dmapping::Compensate
// (src/dataHandler.cpp:93-122) and the IMU alignment transform (src/laserProcessingNode.cpp:113-116) as ONE per-point kernel.
// The per-point ImuHandler::Get (std::lower_bound over the whole history, :51-69) becomes a binary search over the device copy
// of the samples restricted to the scan's time window.
#include "imu.cuh"

#include <algorithm>
#include <cmath>

#include "odom_math.cuh"

namespace floam {

void imu_push(ImuDevice& imu, double stamp, const double q_xyzw[4]) {
  ImuSample s{stamp, {q_xyzw[0], q_xyzw[1], q_xyzw[2], q_xyzw[3]}};
  if (imu.total() == 0) { imu.first_stamp = stamp; imu.host.push_back(s); return; }
  const double tdiff = stamp - imu.host.back().stamp;
  if (!(tdiff > 0.00001)) return;
  imu.host.push_back(s);
  // sliding window: never hold more than dev_cap / 2 + a slack of dev_cap / 8 samples; drop the oldest in one go (amortised O(1))
  const size_t keep = (size_t)imu.dev_cap / 2, slack = (size_t)imu.dev_cap / 8;
  if (imu.dev_cap > 0 && imu.host.size() > keep + slack) {
    const size_t drop = imu.host.size() - keep;
    imu.host.erase(imu.host.begin(), imu.host.begin() + (long)drop);
    imu.base += (long long)drop;
  }
}

// global number of the sample ImuHandler::Get returns for tStamp, or -1 when the validity rule (:57) fails
static long long imu_lookup(const ImuDevice& imu, double t) {
  const std::vector<ImuSample>& v = imu.host;
  const long long n = (long long)v.size();
  long long lo = 0, hi = n;  // lower_bound: first stamp >= t
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (v[(size_t)mid].stamp < t) lo = mid + 1; else hi = mid;
  }
  const long long after = imu.base + lo;          // global numbers from here on
  if (after == 0 || after == imu.total()) return -1;
  if (lo == 0) return -1;                         // the sample before it left the window long ago (only for stamps > 20 min in the past)
  const long long before = after - 1;
  if (before == 0) return -1;
  return before;
}

bool imu_get(const ImuDevice& imu, double stamp, double q[4]) {
  const long long g = imu_lookup(imu, stamp);
  if (g < 0) return false;
  const ImuSample& before = imu.host[(size_t)(g - imu.base)];
  if (imu.slerp) {   // opt-in: what Interpolate(tSlerp, before, after) was meant to do (:48-50, :61-62)
    const ImuSample& after = imu.host[(size_t)(g + 1 - imu.base)];   // exists: the validity rule requires it
    m::quat_slerp((stamp - before.stamp) / (after.stamp - before.stamp), before.q, after.q, q);
    return true;
  }
  for (int k = 0; k < 4; ++k) q[k] = before.q[k];
  return true;
}

bool imu_time_contained(const ImuDevice& imu, double t) {
  return imu.total() > 0 && t >= imu.first_stamp && t <= imu.host.back().stamp;
}

static double stamp_to_sec(uint64_t stamp_us) {  // pcl_conversions::fromPCL + ros::Time::toSec
  const uint64_t ns = stamp_us * 1000ull;
  return (double)(ns / 1000000000ull) + 1e-9 * (double)(ns % 1000000000ull);
}
static uint64_t sec_to_stamp(double t) {         // ros::Time(double) + pcl_conversions::toPCL (truncating to microseconds)
  uint64_t sec = (uint64_t)std::floor(t);
  uint64_t nsec = (uint64_t)std::llround((t - (double)sec) * 1e9);
  sec += nsec / 1000000000ull;
  nsec %= 1000000000ull;
  return nsec / 1000ull + sec * 1000000ull;
}

void deskew_plan(const ImuDevice& imu, uint64_t stamp_us, float time_front, float time_back, const double extr[4], int flags, DeskewPlan* p) {
  p->do_center = (flags & FLOAM_DESKEW_CENTER_TIME) ? 1 : 0;
  p->do_compensate = (flags & FLOAM_DESKEW_COMPENSATE) ? 1 : 0;
  p->do_align = (flags & FLOAM_DESKEW_ALIGN) ? 1 : 0;
  const double tScan = stamp_to_sec(stamp_us);
  const double tEnd = tScan + time_back;
  const double tBegin = tScan + time_front;
  const double tCenter = tBegin + (tEnd - tBegin) / 2.0;
  p->t_scan_old = tScan;
  p->t_center = tCenter;
  p->stamp_us_new = p->do_center ? sec_to_stamp(tCenter) : stamp_us;
  p->t_scan_new = stamp_to_sec(p->stamp_us_new);
  for (int k = 0; k < 4; ++k) p->extr[k] = extr[k];
  // Compensate works on the re-centred times (float store of pnt.time + tScan - tCenter)
  const float tf = p->do_center ? (float)((double)time_front + tScan - tCenter) : time_front;
  const float tb = p->do_center ? (float)((double)time_back + tScan - tCenter) : time_back;
  const double t0 = (double)tf + p->t_scan_new, t1 = (double)tb + p->t_scan_new;
  p->can_compensate = (imu_time_contained(imu, t0) && imu_time_contained(imu, t1)) ? 1 : 0;
  double qi[4] = {0, 0, 0, 0};  // default sensor_msgs::Imu orientation when Get fails (:71-75)
  imu_get(imu, p->t_scan_new, qi);
  double q[4];
  m::quat_mul(qi, extr, q);
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n2 > 0) { p->q_init_inv[0] = -q[0] / n2; p->q_init_inv[1] = -q[1] / n2; p->q_init_inv[2] = -q[2] / n2; p->q_init_inv[3] = q[3] / n2; }
  else { p->q_init_inv[0] = p->q_init_inv[1] = p->q_init_inv[2] = p->q_init_inv[3] = 0.0; }
  m::quat_to_matrix(q, p->R_align);  // Eigen::Affine3d ImuNowT(q), q = Imu(stamp) * extrinsics
}

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) deskew_align_kernel(PointIRT* __restrict__ pts, const int* __restrict__ d_n, const DeskewPlan* __restrict__ d_plan,
                                                                 const ImuSample* __restrict__ samples) {
  pdl_prologue();
  const int n = *d_n;
  const DeskewPlan plan = *d_plan;
  const long long g_lo = plan.g_lo, g_hi = plan.g_hi;
  const int mask = plan.ring_mask;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    float4* raw = reinterpret_cast<float4*>(pts + i);
    float4 a = raw[0], b = raw[1];  // b = intensity, ring|pad, time, pad
    // CenterTime: pnt.time = pnt.time + tScan - tCenter (float + double - double, float store)
    const float t_new = plan.do_center ? (float)(((double)b.z + plan.t_scan_old) - plan.t_center) : b.z;
    b.z = t_new;
    if (plan.can_compensate && plan.do_compensate) {
      const double t_cur = plan.t_scan_new + (double)t_new;
      // ImuHandler::Get: sample strictly before lower_bound(t_cur); invalid -> zero quaternion
      long long lo = g_lo, hi = g_hi;   // global sample numbers; slot in the device ring = number & mask
      while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(&samples[(int)(mid & mask)].stamp) < t_cur) lo = mid + 1; else hi = mid;
      }
      double qi[4] = {0.0, 0.0, 0.0, 0.0};
      if (lo != 0 && lo != g_hi && lo - 1 != 0 && lo > g_lo) {
        const ImuSample* before = samples + (int)((lo - 1) & mask);
#pragma unroll
        for (int k = 0; k < 4; ++k) qi[k] = __ldg(&before->q[k]);
        if (plan.slerp) {
          const ImuSample* after = samples + (int)(lo & mask);
          double qa[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) qa[k] = __ldg(&after->q[k]);
          const double t0 = __ldg(&before->stamp), t1 = __ldg(&after->stamp);
          double qs[4];
          m::quat_slerp((t_cur - t0) / (t1 - t0), qi, qa, qs);
#pragma unroll
          for (int k = 0; k < 4; ++k) qi[k] = qs[k];
        }
      }
      double q_now[4], q_diff[4];
      m::quat_mul(qi, plan.extr, q_now);
      m::quat_mul(plan.q_init_inv, q_now, q_diff);
      const m::V3 pt = m::quat_rotate(q_diff, m::V3{(double)a.x, (double)a.y, (double)a.z});
      // compensated cloud is stored as float, then pcl::transformPointCloud(Affine3d): double R*p + 0, float store
      a.x = (float)pt.x; a.y = (float)pt.y; a.z = (float)pt.z;
      a.w = 1.0f;
    }
    if (plan.can_compensate && plan.do_align) {
      const double x = (double)a.x, y = (double)a.y, z = (double)a.z;
      const double* R = plan.R_align;
      a.x = (float)(R[0] * x + R[1] * y + R[2] * z + 0.0);
      a.y = (float)(R[3] * x + R[4] * y + R[5] * z + 0.0);
      a.z = (float)(R[6] * x + R[7] * y + R[8] * z + 0.0);
    }
    raw[0] = a; raw[1] = b;
  }
}

}  // namespace

int deskew_upload(ImuDevice& imu, DeskewPlan& plan, DeskewPlan* d_plan, cudaStream_t copy) {
  const long long total = imu.total();
  long long from = imu.dev_count > imu.base ? imu.dev_count : imu.base;   // samples dropped before they were uploaded are of no use any more
  while (from < total) {   // ring upload: at most two pieces per call in practice
    const int slot = (int)(from & (imu.dev_cap - 1));
    const long long piece = std::min<long long>(total - from, imu.dev_cap - slot);
    FLOAM_CUDA_OK(cudaMemcpyAsync(imu.d_samples + slot, imu.host.data() + (size_t)(from - imu.base), (size_t)piece * sizeof(ImuSample), cudaMemcpyHostToDevice, copy));
    from += piece;
  }
  imu.dev_count = total;
  plan.g_hi = total;
  plan.g_lo = imu.base;   // the host window never exceeds 5/8 of the ring, so [g_lo, g_hi) is resident and stays so for 3/8 of a ring more
  plan.ring_mask = imu.dev_cap - 1;
  plan.slerp = imu.slerp ? 1 : 0;
  FLOAM_CUDA_OK(cudaMemcpyAsync(d_plan, &plan, sizeof(DeskewPlan), cudaMemcpyHostToDevice, copy));  // pageable source: staged before the call returns
  return FLOAM_OK;
}

void deskew_launch(ImuDevice& imu, const DeskewPlan* d_plan, PointIRT* d_pts, const int* d_n, int n_max, cudaStream_t s) {
  int g = (n_max + kThreads - 1) / kThreads;
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  if (g < 1) g = 1;
  FLOAM_LAUNCH(K_DESKEW_ALIGN, deskew_align_kernel, g, kThreads, s, d_pts, d_n, d_plan, imu.d_samples);
}

int deskew_align_device(ImuDevice& imu, const DeskewPlan& plan_in, PointIRT* d_pts, const int* d_n, int n_max, cudaStream_t s) {
  DeskewPlan plan = plan_in;
  const int rc = deskew_upload(imu, plan, imu.d_plan, s);
  if (rc) return rc;
  deskew_launch(imu, imu.d_plan, d_pts, d_n, n_max, s);
  return FLOAM_OK;
}

}  // namespace floam
