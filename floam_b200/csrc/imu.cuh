// dmapping::ImuHandler mirror and the IMU-aided deskew (CenterTime + dmapping::Compensate + IMU alignment) on the device.
// Reference: src/dataHandler.cpp:24-122, src/laserProcessingNode.cpp:65-78,108-116. See imu.cu.
#pragma once
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace floam {

struct ImuSample {   // device record
  double stamp;
  double q[4];       // x, y, z, w (Eigen coefficient order)
};

struct DeskewPlan;
// The reference keeps every sample for ever (std::vector, :34).  Here the history is a sliding window: sample number g (counted from
// the first one ever pushed) lives at host[g - base] and at d_samples[g & (dev_cap - 1)] (a ring; dev_cap is a power of two).  The
// host drops the oldest samples once more than dev_cap / 2 are held, lookups search the newest dev_cap / 2 only, and an upload only
// overwrites ring slots more than dev_cap samples old — so nothing a frame in flight can still read is touched (40+ minutes of
// history at 400 Hz with the default 2^20 ring).  Validity rules that refer to "the first sample" (:57, :77) keep referring to the
// first sample ever pushed: they use the global number / the remembered first stamp.
struct ImuDevice {
  std::vector<ImuSample> host;   // time-sorted, same admission rule as ImuHandler::AddMsg (:24-40); window [base, base + host.size())
  long long base = 0;            // global number of host[0]
  double first_stamp = 0.0;      // stamp of sample 0 (data_.front() of the reference)
  ImuSample* d_samples = nullptr;
  long long dev_count = 0;       // samples already uploaded (global count)
  int dev_cap = 0;               // ring size, power of two
  bool slerp = false;            // FLOAM_FIX_IMU_SLERP: Get() interpolates instead of holding the sample before the stamp
  struct DeskewPlan* d_plan = nullptr;   // plan of the stand-alone entry point
  long long total() const { return base + (long long)host.size(); }
};

// ImuHandler::AddMsg: keeps the sample iff it is the first or more than 10 us after the previous one
void imu_push(ImuDevice& imu, double stamp, const double q_xyzw[4]);
// ImuHandler::Get(t, data) (:51-69): zero-order hold with the reference's validity rule; false -> q untouched
bool imu_get(const ImuDevice& imu, double stamp, double q_xyzw[4]);
bool imu_time_contained(const ImuDevice& imu, double t);   // :76-81

struct DeskewPlan {       // everything the per-point kernel needs, computed on the host from the scan header
  double t_scan_old;      // stamp before CenterTime
  double t_center;        // mid-scan time
  double t_scan_new;      // stamp after CenterTime (microsecond-truncated like pcl_conversions)
  double q_init_inv[4];   // (Imu(t_scan_new) * extrinsics)^-1
  double extr[4];
  double R_align[9];      // rotation matrix of Imu(t_scan_new) * extrinsics (row-major)
  int can_compensate;     // dmapping::Compensate's return value
  int do_center, do_compensate, do_align;   // which of CenterTime / Compensate / alignment run (floam_deskew_flags)
  uint64_t stamp_us_new;
  long long g_lo, g_hi;   // global sample numbers [g_lo, g_hi) the kernel may search (resident in the device ring when it runs)
  int ring_mask;          // dev_cap - 1
  int slerp;              // FLOAM_FIX_IMU_SLERP
};
// ros::Time / pcl stamp conversions + CenterTime + the host part of Compensate (TimeContained, qInit) and of the alignment
void deskew_plan(const ImuDevice& imu, uint64_t stamp_us, float time_front, float time_back, const double extr_xyzw[4], int flags, DeskewPlan* plan);
// uploads new samples (if any) and runs the fused per-point kernel in place: time re-centring always; rotation-only deskew and
// alignment only when plan.can_compensate
int deskew_align_device(ImuDevice& imu, const DeskewPlan& plan, PointIRT* d_pts, const int* d_n, int n_max, cudaStream_t s);
// graph-friendly split of the same thing: (1) uploads new IMU samples and the plan into d_plan on `copy`, (2) launches the kernel,
// which reads the plan from device memory, on `s` (fixed arguments -> capturable)
int deskew_upload(ImuDevice& imu, DeskewPlan& plan, DeskewPlan* d_plan, cudaStream_t copy);
void deskew_launch(ImuDevice& imu, const DeskewPlan* d_plan, PointIRT* d_pts, const int* d_n, int n_max, cudaStream_t s);

}  // namespace floam
