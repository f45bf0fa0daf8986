// ORACLE — TEST INFRASTRUCTURE ONLY.  C interface over the REFERENCE'S OWN class sources, compiled unmodified from
// /root/reference/src/{laserProcessingClass,dataHandler,lidar,lidarOptimization,odomEstimationClass,laserMappingClass}.cpp
// (+ laserProcessingNode.cpp for CenterTime) against the stand-in third-party headers of oracle/stubs/ into
// oracle/_ref/libfloam_ref.so (recipe: oracle/Makefile, target `ref`).  Same entry-point names and signatures as oracle/capi.cpp, so
// oracle/pyoracle.py drives either library; tests pin the restatement (and the CUDA path) against this one.
//
// What is the reference's and what is not: every line of the classes above is the reference's.  PCL containers are re-typed
// stand-ins; VoxelGrid / CropBox / KdTreeFLANN / Eigen solvers / ceres::Solve behind their real interfaces are the restatements
// of oracle/*.cpp — those libraries are not in this image.
//
// `private` is opened for this translation unit only so that stage tests can read / seed last_odom, optimization_count and
// parameters (same object layout; the class sources themselves are compiled untouched).
#include <algorithm>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <math.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/filters/filter.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/filters/passthrough.h>
#include <pcl/kdtree/kdtree_flann.h>
#include <pcl/filters/statistical_outlier_removal.h>
#include <pcl/filters/extract_indices.h>
#include <pcl/filters/crop_box.h>
#include <pcl_ros/impl/transforms.hpp>
#include <ceres/ceres.h>
#include <ceres/rotation.h>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <ros/ros.h>
#include "lidar.h"
#include "lidarOptimization.h"
#include "dataHandler.h"
#include "utils.h"
#define private public
#include "odomEstimationClass.h"
#include "laserMappingClass.h"
#include "laserProcessingClass.h"
#undef private
#include <chrono>
#include <cstring>

void CenterTime(const pcl::PointCloud<vel_point::PointXYZIRT>::Ptr cloud);  // src/laserProcessingNode.cpp:65-78 (that file is built with -Dmain=...)

// the node source calls these; nothing in this harness reaches them
template <typename T> void pcl::fromROSMsg(const sensor_msgs::PointCloud2&, pcl::PointCloud<T>&) { std::fprintf(stderr, "fromROSMsg: not part of the stand-in\n"); std::abort(); }
template <typename T> void pcl::toROSMsg(const pcl::PointCloud<T>&, sensor_msgs::PointCloud2&) { std::fprintf(stderr, "toROSMsg: not part of the stand-in\n"); std::abort(); }
template void pcl::fromROSMsg<vel_point::PointXYZIRT>(const sensor_msgs::PointCloud2&, pcl::PointCloud<vel_point::PointXYZIRT>&);
template void pcl::toROSMsg<vel_point::PointXYZIRT>(const pcl::PointCloud<vel_point::PointXYZIRT>&, sensor_msgs::PointCloud2&);

// dump-on-exit writers of the odometry node (src/odomEstimationNode.cpp:66-121; that file is built with -Dmain=... and its other symbols localised)
void SaveMerged(const std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr> clouds, const std::vector<Eigen::Affine3d> poses, const std::string& directory, double downsample_size);
void SavePosesHomogeneousBALM(const std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr> clouds, const std::vector<Eigen::Affine3d> poses, const std::string& directory, double downsample_size);

namespace {
typedef pcl::PointCloud<vel_point::PointXYZIRT> CloudIRT;
typedef pcl::PointCloud<pcl::PointXYZI> CloudI;
static_assert(sizeof(vel_point::PointXYZIRT) == 32, "PointXYZIRT is 32 bytes (include/lidar.h:14-22)");
static_assert(offsetof(vel_point::PointXYZIRT, intensity) == 16 && offsetof(vel_point::PointXYZIRT, ring) == 20 && offsetof(vel_point::PointXYZIRT, time) == 24,
              "PointXYZIRT field offsets");

CloudIRT::Ptr make_irt(const void* pts, int n) {
  CloudIRT::Ptr c(new CloudIRT());
  c->points.resize(n);
  if (n) std::memcpy(static_cast<void*>(c->points.data()), pts, 32 * (size_t)n);
  c->width = n; c->height = 1;
  return c;
}
CloudI::Ptr make_i(const void* pts, int n) {
  CloudI::Ptr c(new CloudI());
  c->points.resize(n);
  if (n) std::memcpy(static_cast<void*>(c->points.data()), pts, 32 * (size_t)n);
  c->width = n; c->height = 1;
  return c;
}
template <class C>
int copy_out(const C& cloud, void* out, int cap) {
  const int n = (int)cloud.points.size();
  if (out && cap > 0 && n > 0) std::memcpy(out, static_cast<const void*>(cloud.points.data()), 32 * (size_t)std::min(n, cap));
  return n;
}
lidar::Lidar make_lidar(int num_lines, double scan_period, double min_dis, double max_dis) {
  lidar::Lidar p;  // set the way the nodes do (src/laserProcessingNode.cpp:190-194)
  p.setScanPeriod(scan_period); p.setVerticalAngle(2.0); p.setLines(num_lines); p.setMaxDistance(max_dis); p.setMinDistance(min_dis);
  return p;
}
void put_iso(const Eigen::Isometry3d& T, double* o) {
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) o[i * 4 + j] = T(i, j);
}
void get_iso(const double* o, Eigen::Isometry3d& T) {
  T = Eigen::Isometry3d::Identity();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) T(i, j) = o[i * 4 + j];
}
void pack_lm(const fo::LmSummary& s, double* o) {
  o[0] = s.iterations; o[1] = s.accepted; o[2] = s.initial_cost; o[3] = s.final_cost; o[4] = s.termination;
  std::memcpy(o + 5, s.H0, sizeof(s.H0)); std::memcpy(o + 41, s.g0, sizeof(s.g0));
}

struct OdomHandle {
  OdomEstimationClass est;
  std::vector<fo::LmSummary> solves;  // of the last update call
  size_t kf_count_before = 0;
  bool keyframe = false;
  CloudI ds_edge, ds_surf;
};
struct ImuHandle {
  dmapping::ImuHandler h;
};

void begin_update(OdomHandle* h) {
  h->solves.clear();
  ceres::floam_stub::solve_log() = &h->solves;
}
void end_update(OdomHandle* h, const Eigen::Isometry3d& before_last_kf, bool had_kf) {
  ceres::floam_stub::solve_log() = NULL;
  // KeyFrameUpdate (src/odomEstimationClass.cpp:320-343) pushed a keyframe iff the newest stored pose changed
  const keyframes& k = h->est.keyframes_;
  h->keyframe = false;
  if (!k.empty()) {
    bool same = had_kf;
    if (had_kf) for (int i = 0; i < 4 && same; ++i) for (int j = 0; j < 4; ++j) if (k.back().pose(i, j) != before_last_kf(i, j)) { same = false; break; }
    h->keyframe = !same;
    if (h->keyframe) { h->ds_edge = *k.back().edge_cloud; h->ds_surf = *k.back().surf_cloud; }
  }
}
}  // namespace

extern "C" {

const char* fo_backend() { return "reference"; }

// ---------- LaserProcessingClass::featureExtraction (src/laserProcessingClass.cpp:72-118) ----------
// The selection never reads `intensity` (it is only copied through, :17,:146,:225): the harness passes the input index in that
// field to recover which scan point every feature is, then puts the real intensity back.  total_order is ignored: the sort is the
// reference's own std::sort (:123-126).
int fo_feature_extract(const void* pts, int n, int num_lines, double min_dis, double max_dis, int total_order, void* edge, int* edge_src, int edge_cap,
                       int* ne, void* surf, int* surf_src, int surf_cap, int* ns, long* ties) {
  (void)total_order;
  LaserProcessingClass lp;
  lp.init(make_lidar(num_lines, 0.1, min_dis, max_dis));
  CloudIRT::Ptr in = make_irt(pts, n);
  std::vector<float> real_intensity(n);
  for (int i = 0; i < n; ++i) { real_intensity[i] = in->points[i].intensity; in->points[i].intensity = (float)i; }
  CloudIRT::Ptr e(new CloudIRT()), s(new CloudIRT());
  lp.featureExtraction(in, e, s);
  auto restore = [&](CloudIRT& c, int* src, int cap) {
    for (size_t i = 0; i < c.points.size(); ++i) {
      const int id = (int)c.points[i].intensity;
      if (src && (int)i < cap) src[i] = id;
      c.points[i].intensity = real_intensity[id];
    }
  };
  restore(*e, edge_src, edge_cap); restore(*s, surf_src, surf_cap);
  *ne = copy_out(*e, edge, edge_cap);
  *ns = copy_out(*s, surf, surf_cap);
  if (ties) *ties = -1;  // not observable from outside the reference
  return 0;
}

// ---------- lidar.cpp / lidarOptimization.cpp ----------
void fo_euler2quat(double roll, double pitch, double yaw, double q_xyzw[4]) {  // src/lidar.cpp:8-16
  const Eigen::Quaterniond q = euler2Quaternion(roll, pitch, yaw);
  q_xyzw[0] = q.x(); q_xyzw[1] = q.y(); q_xyzw[2] = q.z(); q_xyzw[3] = q.w();
}
void fo_se3_plus(const double x[7], const double delta[6], double out[7]) {  // PoseSE3Parameterization::Plus :77-91
  PoseSE3Parameterization p;
  p.Plus(x, delta, out);
}
// residual record = 10 doubles: kind (0 edge / 1 surf), curr(3), a(3), b(3)   [surf: a = unit normal, b[0] = negative_OA_dot_norm]
static ceres::CostFunction* make_cost(const double* r) {
  const Eigen::Vector3d c(r[1], r[2], r[3]), a(r[4], r[5], r[6]), b(r[7], r[8], r[9]);
  if ((int)r[0] == 0) return new EdgeAnalyticCostFunction(c, a, b);
  return new SurfNormAnalyticCostFunction(c, a, r[7]);
}
int fo_evaluate_residual(const double rec[10], const double x[7], double* r, double jac7[7]) {  // ::Evaluate :12-43, :51-74
  ceres::CostFunction* cf = make_cost(rec);
  double const* params[1] = {x};
  double* jacs[1] = {jac7};
  const bool ok = cf->Evaluate(params, r, jac7 ? jacs : NULL);
  delete cf;
  bool fin = std::isfinite(*r);
  if (jac7) for (int j = 0; j < 7; ++j) fin = fin && std::isfinite(jac7[j]);
  return ok && fin ? 0 : 1;
}
// the problem exactly as src/odomEstimationClass.cpp:83-108 builds and solves it, over the given residual blocks
int fo_lm_solve(const double* recs, int n, int loss, double x[7], int max_iter, double* summary_out) {
  ceres::LossFunction* loss_function = NULL;
  if (loss == 1) loss_function = new ceres::HuberLoss(0.1);
  else if (loss == 2) loss_function = new ceres::CauchyLoss(0.2);
  ceres::Problem::Options problem_options;
  ceres::Problem problem(problem_options);
  problem.AddParameterBlock(x, 7, new PoseSE3Parameterization());
  for (int i = 0; i < n; ++i) problem.AddResidualBlock(make_cost(recs + 10 * i), loss_function, x);
  if (n == 0) delete loss_function;
  ceres::Solver::Options options;
  options.linear_solver_type = ceres::DENSE_QR;
  options.max_num_iterations = max_iter;
  options.minimizer_progress_to_stdout = false;
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);
  if (summary_out) pack_lm(summary.lm, summary_out);
  return 0;
}

// ---------- dataHandler.cpp + the node's folding of it ----------
void* fo_imu_create() { return new ImuHandle(); }
void fo_imu_destroy(void* h) { delete (ImuHandle*)h; }
void fo_imu_add(void* h, double stamp, const double q_xyzw[4]) {
  boost::shared_ptr<sensor_msgs::Imu> msg(new sensor_msgs::Imu());
  msg->header.stamp.fromSec(stamp);
  msg->orientation.x = q_xyzw[0]; msg->orientation.y = q_xyzw[1]; msg->orientation.z = q_xyzw[2]; msg->orientation.w = q_xyzw[3];
  ((ImuHandle*)h)->h.AddMsg(msg);
}
int fo_imu_size(void* h) { return (int)((ImuHandle*)h)->h.size(); }
int fo_imu_get(void* h, double t, double q_xyzw[4]) {
  sensor_msgs::Imu data;
  // Get() dereferences std::prev(begin()) when nothing precedes t; the condition at :57 then rejects it.  Guard the empty buffer only.
  const bool ok = ((ImuHandle*)h)->h.size() > 0 && ((ImuHandle*)h)->h.Get(t, data);
  q_xyzw[0] = data.orientation.x; q_xyzw[1] = data.orientation.y; q_xyzw[2] = data.orientation.z; q_xyzw[3] = data.orientation.w;
  return ok ? 1 : 0;
}
// the sequence of src/laserProcessingNode.cpp:99-116: CenterTime, Compensate, IMU alignment.  returns 0 ok, 1 = cannot compensate.
int fo_deskew_align(void* h, void* pts, int n, unsigned long long* stamp_us, const double extr_xyzw[4]) {
  dmapping::ImuHandler& imuHandler = ((ImuHandle*)h)->h;
  Eigen::Quaterniond exstrinsics(extr_xyzw[3], extr_xyzw[0], extr_xyzw[1], extr_xyzw[2]);
  CloudIRT::Ptr pointcloud_in = make_irt(pts, n);
  pointcloud_in->header.stamp = *stamp_us;
  CloudIRT::Ptr compensated(new CloudIRT()), imu_aligned(new CloudIRT());
  CenterTime(pointcloud_in);                                                                    // :100
  ros::Time pointcloud_time = pcl_conversions::fromPCL(pointcloud_in->header.stamp);            // :101
  *stamp_us = pointcloud_in->header.stamp;
  bool can_compensate = dmapping::Compensate(*pointcloud_in, *compensated, imuHandler, exstrinsics);  // :108
  if (!can_compensate) { copy_out(*pointcloud_in, pts, n); return 1; }
  Eigen::Quaterniond q(dmapping::Imu2Orientation(imuHandler.Get(pointcloud_time.toSec())) * exstrinsics);  // :113
  Eigen::Affine3d ImuNowT(q);                                                                   // :114
  pcl::transformPointCloud(*compensated, *imu_aligned, ImuNowT);                                // :116
  copy_out(*imu_aligned, pts, n);
  return 0;
}
void fo_compensate_velocity(void* pts, int n, const double v[3]) {  // dmapping::CompensateVelocity :82-91
  CloudIRT::Ptr c = make_irt(pts, n);
  dmapping::CompensateVelocity(c, Eigen::Vector3d(v[0], v[1], v[2]));
  copy_out(*c, pts, n);
}

// ---------- OdomEstimationClass ----------
// total_order / use_kdtree set the stand-in libraries' test knobs (process-wide): stable order inside a voxel and brute-force
// (distance, index) neighbour order = the deterministic contract of the CUDA path.  Defaults (0, 1) = what PCL / FLANN do.
// KeyFrameUpdate's `first` flag is function-static in the reference (Q10): ONE instance per loaded copy of this library.
void* fo_odom_create(int num_lines, double scan_period, double min_dis, double max_dis, double map_resolution, const char* loss, int total_order, int use_kdtree) {
  OdomHandle* h = new OdomHandle();
  pcl::floam_stub::voxel_total_order() = total_order != 0;
  pcl::floam_stub::knn_bruteforce() = use_kdtree == 0;
  h->est.init(make_lidar(num_lines, scan_period, min_dis, max_dis), map_resolution, loss);
  return h;
}
void fo_odom_destroy(void* h) { delete (OdomHandle*)h; }
void fo_odom_init_map(void* h, const void* edge, int ne, const void* surf, int ns) {
  ((OdomHandle*)h)->est.initMapWithPoints(make_i(edge, ne), make_i(surf, ns));
}
static void report_pose(OdomEstimationClass& est, double* pose) {
  if (pose) std::memcpy(pose, est.parameters, sizeof(double) * 7);
}
void fo_odom_update(void* hv, void* edge, int ne, void* surf, int ns, int deskew, double pose_q_xyzw_t[7]) {
  OdomHandle* h = (OdomHandle*)hv;
  CloudIRT::Ptr e = make_irt(edge, ne), s = make_irt(surf, ns);
  const bool had = !h->est.keyframes_.empty();
  const Eigen::Isometry3d last_kf = had ? h->est.keyframes_.back().pose : Eigen::Isometry3d::Identity();
  begin_update(h);
  h->est.UpdatePointsToMapSelector(e, s, deskew != 0);
  end_update(h, last_kf, had);
  if (deskew) { copy_out(*e, edge, ne); copy_out(*s, surf, ns); }  // the reference deskews the caller's clouds in place (:42-43)
  report_pose(h->est, pose_q_xyzw_t);
}
void fo_odom_update_xyzi(void* hv, const void* edge, int ne, const void* surf, int ns, int type, double pose_q_xyzw_t[7]) {
  OdomHandle* h = (OdomHandle*)hv;
  const bool had = !h->est.keyframes_.empty();
  const Eigen::Isometry3d last_kf = had ? h->est.keyframes_.back().pose : Eigen::Isometry3d::Identity();
  begin_update(h);
  h->est.updatePointsToMap(make_i(edge, ne), make_i(surf, ns), (OdomEstimationClass::UpdateType)type);
  end_update(h, last_kf, had);
  report_pose(h->est, pose_q_xyzw_t);
}
void fo_odom_get(void* hv, double odom16[16], double last_odom16[16], double velocity[3], int* optimization_count) {
  OdomEstimationClass& est = ((OdomHandle*)hv)->est;
  if (odom16) put_iso(est.odom, odom16);
  if (last_odom16) put_iso(est.last_odom, last_odom16);
  if (velocity) { const Eigen::Vector3d v = est.GetVelocity(); velocity[0] = v(0); velocity[1] = v(1); velocity[2] = v(2); }
  if (optimization_count) *optimization_count = est.optimization_count;
}
void fo_odom_set_state(void* hv, const double odom16[16], const double last_odom16[16], int optimization_count) {
  OdomEstimationClass& est = ((OdomHandle*)hv)->est;
  get_iso(odom16, est.odom); get_iso(last_odom16, est.last_odom);
  est.optimization_count = optimization_count;
}
void fo_odom_set_map(void* hv, const void* edge, int ne, const void* surf, int ns) {
  OdomEstimationClass& est = ((OdomHandle*)hv)->est;
  *est.laserCloudCornerMap = *make_i(edge, ne);
  *est.laserCloudSurfMap = *make_i(surf, ns);
}
int fo_odom_map_sizes(void* hv, int* n_edge, int* n_surf) {
  OdomEstimationClass& est = ((OdomHandle*)hv)->est;
  *n_edge = (int)est.laserCloudCornerMap->points.size(); *n_surf = (int)est.laserCloudSurfMap->points.size();
  return 0;
}
int fo_odom_get_map(void* hv, void* edge, int ecap, void* surf, int scap) {
  OdomEstimationClass& est = ((OdomHandle*)hv)->est;
  copy_out(*est.laserCloudCornerMap, edge, ecap); copy_out(*est.laserCloudSurfMap, surf, scap);
  return 0;
}
long fo_odom_knn_queries(void*) { return pcl::floam_stub::knn_query_count(); }
// taps of the last update call that are observable from outside the class: 0/1 = downsampled edge / surf cloud (only when the
// frame became a keyframe: the class stores them there), 9 = LM summary of the last solve, 10 = {solves run, keyframe}.
int fo_odom_debug(void* hv, int what, void* out, int cap) {
  OdomHandle* h = (OdomHandle*)hv;
  switch (what) {
    case 0: return h->keyframe ? copy_out(h->ds_edge, out, cap) : 0;
    case 1: return h->keyframe ? copy_out(h->ds_surf, out, cap) : 0;
    case 9:
      if (cap >= 47) { if (h->solves.empty()) std::memset(out, 0, 47 * sizeof(double)); else pack_lm(h->solves.back(), (double*)out); }
      return 47;
    case 10:
      if (cap >= 2) { ((int*)out)[0] = (int)h->solves.size(); ((int*)out)[1] = h->keyframe ? 1 : 0; }
      return 2;
    default: return 0;
  }
}

// ---------- LaserMappingClass ----------
void* fo_mapping_create(double map_resolution, int total_order) {
  pcl::floam_stub::voxel_total_order() = total_order != 0;
  LaserMappingClass* m = new LaserMappingClass();
  m->init(map_resolution);
  return m;
}
void fo_mapping_destroy(void* m) { delete (LaserMappingClass*)m; }
void fo_mapping_update(void* m, const void* pts, int n, const double pose16[16]) {
  Eigen::Isometry3d T;
  get_iso(pose16, T);
  ((LaserMappingClass*)m)->updateCurrentPointsToMap(make_i(pts, n), T);
}
int fo_mapping_get_map(void* m, void* out, int cap) { return copy_out(*((LaserMappingClass*)m)->getMap(), out, cap); }

}  // extern "C"

// ---------- on-disk outputs (src/utils.cpp:3-106, src/odomEstimationNode.cpp:66-121) ----------
namespace {
struct DumpArgs {
  std::vector<Eigen::Affine3d> poses;
  std::vector<double> stamps;
  std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr> clouds;
};
// poses16: row-major 4x4 per scan; stamps: seconds; clouds: concatenated 32-byte PointXYZI, offsets[n + 1]
DumpArgs make_dump(const double* poses16, const double* stamps, const void* clouds, const long long* offsets, int n) {
  DumpArgs d;
  for (int i = 0; i < n; ++i) {
    Eigen::Matrix4d M;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) M(r, c) = poses16[16 * i + 4 * r + c];
    d.poses.push_back(Eigen::Affine3d(M));
    d.stamps.push_back(stamps[i]);
    pcl::PointCloud<pcl::PointXYZI>::Ptr c = make_i((const char*)clouds + 32 * offsets[i], (int)(offsets[i + 1] - offsets[i]));
    pcl_conversions::toPCL(ros::Time(stamps[i]), c->header.stamp);   // the node stamps every stored cloud (src/odomEstimationNode.cpp:279)
    d.clouds.push_back(c);
  }
  return d;
}
}  // namespace
extern "C" {
void fo_save_posegraph(const char* dir, const double* poses16, const double* stamps, const void* clouds, const long long* offsets, int n) {
  DumpArgs d = make_dump(poses16, stamps, clouds, offsets, n);
  SavePosegraph(dir, d.poses, d.stamps, d.clouds);
}
void fo_save_odom(const char* dir, const double* poses16, const double* stamps, const void* clouds, const long long* offsets, int n) {
  DumpArgs d = make_dump(poses16, stamps, clouds, offsets, n);
  SaveOdom(dir, d.poses, d.stamps, d.clouds);
}
void fo_save_balm(const char* dir, const double* poses16, const double* stamps, const void* clouds, const long long* offsets, int n) {
  DumpArgs d = make_dump(poses16, stamps, clouds, offsets, n);
  SavePosesHomogeneousBALM(d.clouds, d.poses, dir, 0.0);
}
void fo_save_merged(const char* dir, const double* poses16, const double* stamps, const void* clouds, const long long* offsets, int n, double downsample_size,
                    int total_order) {
  pcl::floam_stub::voxel_total_order() = total_order != 0;
  DumpArgs d = make_dump(poses16, stamps, clouds, offsets, n);
  SaveMerged(d.clouds, d.poses, dir, downsample_size);
}
}  // extern "C"

// ---------- timed whole-sequence replay: featureExtraction + odometry per frame on one thread, the way the nodes call them ----------
static double replay_impl(const void* scans, const long long* offsets, int n_frames, int num_lines, double scan_period, double min_dis, double max_dis,
                          double map_resolution, const char* loss, int deskew, double* poses_out, double* per_frame_ms, long* knn_queries, int stage_skip,
                          double* stage_ms_out) {
  LaserProcessingClass laserProcessing;
  laserProcessing.init(make_lidar(num_lines, scan_period, min_dis, max_dis));
  OdomEstimationClass odomEstimation;
  odomEstimation.init(make_lidar(num_lines, scan_period, min_dis, max_dis), map_resolution, loss);
  bool is_odom_inited = false;
  double total = 0, feature_s = 0;
  const long q0 = pcl::floam_stub::knn_query_count();
  for (int f = 0; f < n_frames; ++f) {
    auto t0 = std::chrono::steady_clock::now();
    CloudIRT::Ptr pointcloud_in = make_irt((const char*)scans + 32 * offsets[f], (int)(offsets[f + 1] - offsets[f]));
    CloudIRT::Ptr pointcloud_edge(new CloudIRT()), pointcloud_surf(new CloudIRT());
    laserProcessing.featureExtraction(pointcloud_in, pointcloud_edge, pointcloud_surf);  // src/laserProcessingNode.cpp:129
    if (stage_ms_out && f >= stage_skip) feature_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (is_odom_inited == false) {  // src/odomEstimationNode.cpp:218-224
      odomEstimation.initMapWithPoints(VelToIntensityCopy(pointcloud_edge), VelToIntensityCopy(pointcloud_surf));
      is_odom_inited = true;
    } else {
      odomEstimation.UpdatePointsToMapSelector(pointcloud_edge, pointcloud_surf, deskew != 0);  // :228
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    total += ms;
    if (per_frame_ms) per_frame_ms[f] = ms;
    if (poses_out) {
      Eigen::Quaterniond q_current(odomEstimation.odom.rotation());  // :242-244
      Eigen::Vector3d t_current = odomEstimation.odom.translation();
      double* o = poses_out + 7 * f;
      o[0] = q_current.x(); o[1] = q_current.y(); o[2] = q_current.z(); o[3] = q_current.w(); o[4] = t_current.x(); o[5] = t_current.y(); o[6] = t_current.z();
    }
  }
  if (knn_queries) *knn_queries = pcl::floam_stub::knn_query_count() - q0;
  if (stage_ms_out) {  // only the feature stage is separable from outside the class
    const double nf = n_frames > stage_skip ? (double)(n_frames - stage_skip) : 1.0;
    for (int k = 0; k < 6; ++k) stage_ms_out[k] = 0.0;
    stage_ms_out[0] = feature_s * 1e3 / nf;
  }
  return total * 1e-3;
}

extern "C" {
double fo_replay_sequence(const void* scans, const long long* offsets, int n_frames, int num_lines, double scan_period, double min_dis, double max_dis,
                          double map_resolution, const char* loss, int deskew, double* poses_out, double* per_frame_ms, long* knn_queries) {
  return replay_impl(scans, offsets, n_frames, num_lines, scan_period, min_dis, max_dis, map_resolution, loss, deskew, poses_out, per_frame_ms, knn_queries, 0, nullptr);
}
double fo_replay_sequence_stages(const void* scans, const long long* offsets, int n_frames, int num_lines, double scan_period, double min_dis, double max_dis,
                                 double map_resolution, const char* loss, int deskew, double* poses_out, double* per_frame_ms, long* knn_queries, int stage_skip,
                                 double* stage_ms_out) {
  return replay_impl(scans, offsets, n_frames, num_lines, scan_period, min_dis, max_dis, map_resolution, loss, deskew, poses_out, per_frame_ms, knn_queries, stage_skip, stage_ms_out);
}
}  // extern "C"
