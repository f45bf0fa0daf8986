// ORACLE — TEST INFRASTRUCTURE ONLY (see types.h).  Pinned: tests/test_reference_pin.py runs this restatement against the reference's own
// sources compiled unmodified (oracle/_ref, `make ref`) on identical inputs — identical selections, bytes and poses.
// Flat C interface over the restatement so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs can drive it through ctypes.  Nothing under floam_b200/ or include/ may use this.
#include "floam_oracle.h"
#include <chrono>
#include <cstring>

using namespace fo;

namespace {
template <class T>
int copy_out(const std::vector<T>& v, T* out, int cap) {
  int n = (int)v.size();
  if (out && cap > 0) std::memcpy(out, v.data(), sizeof(T) * (size_t)std::min(n, cap));
  return n;
}
struct OdomHandle {
  OdomEstimation est;
  OdomDebug dbg;
};
struct ImuHandle {
  ImuHandler h;
};
}  // namespace

extern "C" {

// ---------- feature extraction ----------
int fo_feature_extract(const PointXYZIRT* pts, int n, int num_lines, double min_dis, double max_dis, int total_order,
                       PointXYZIRT* edge, int* edge_src, int edge_cap, int* ne, PointXYZIRT* surf, int* surf_src, int surf_cap, int* ns,
                       long* ties) {
  LaserProcessing lp;
  LidarParam p; p.num_lines = num_lines; p.min_distance = min_dis; p.max_distance = max_dis;
  lp.init(p);
  CloudIRT in(pts, pts + n), e, s;
  std::vector<int> es, ss;
  FeatureStats st;
  lp.featureExtraction(in, e, s, total_order != 0, &st, &es, &ss);
  *ne = copy_out(e, edge, edge_cap); copy_out(es, edge_src, edge_cap);
  *ns = copy_out(s, surf, surf_cap); copy_out(ss, surf_src, surf_cap);
  if (ties) *ties = st.curvature_ties;
  return 0;
}

// ---------- PCL filters ----------
int fo_voxel_grid(const PointXYZI* pts, int n, float leaf, int total_order, PointXYZI* out, int cap, int* passthrough) {
  CloudI in(pts, pts + n), o;
  bool pt = false;
  voxel_grid_filter(in, leaf, o, total_order != 0, &pt);
  if (passthrough) *passthrough = pt;
  return copy_out(o, out, cap);
}
int fo_crop_box(const PointXYZI* pts, int n, const float mn[3], const float mx[3], PointXYZI* out, int cap) {
  CloudI in(pts, pts + n), o;
  crop_box_filter(in, mn, mx, o);
  return copy_out(o, out, cap);
}

// ---------- kNN ----------
int fo_knn(const PointXYZI* map, int m, const PointXYZI* queries, int nq, int k, int use_kdtree, int* ids, float* d2) {
  CloudI cloud(map, map + m);
  KdTreeFlann tree;
  GridKnn grid;
  if (use_kdtree == 2) {
    grid.setInputCloud(cloud);
    for (int i = 0; i < nq; ++i) grid.nearestKSearch(queries[i], k, ids + (size_t)i * k, d2 + (size_t)i * k);
    return 0;
  }
  if (use_kdtree) tree.setInputCloud(cloud);
  for (int i = 0; i < nq; ++i) {
    int idb[8]; float db[8];
    for (int j = 0; j < 8; ++j) { idb[j] = -1; db[j] = 0.f; }
    if (use_kdtree) tree.nearestKSearch(queries[i], k, idb, db);
    else knn_bruteforce(cloud, queries[i], k, idb, db);
    for (int j = 0; j < k; ++j) { ids[i * k + j] = idb[j]; d2[i * k + j] = db[j]; }
  }
  return 0;
}

// ---------- small linear algebra (for self-validation against numpy) ----------
void fo_eigen3(const double A[9], double values[3], double vectors[9]) {
  Mat3 m;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.m[i][j] = A[i * 3 + j];
  Eigen3 e = self_adjoint_eigen3(m);
  for (int i = 0; i < 3; ++i) { values[i] = e.values[i]; for (int j = 0; j < 3; ++j) vectors[i * 3 + j] = e.vectors[i][j]; }
}
void fo_colpiv_qr_solve(const double* A_rowmajor, const double* b, int rows, double x[3]) { colpiv_qr_solve_nx3(A_rowmajor, b, rows, x); }
void fo_se3_plus(const double x[7], const double delta[6], double out[7]) { se3_plus(x, delta, out); }
void fo_quat_from_matrix(const double R[9], double q_xyzw[4]) {
  Mat3 m; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.m[i][j] = R[i * 3 + j];
  Quat q = quat_from_matrix(m); q_xyzw[0] = q.x; q_xyzw[1] = q.y; q_xyzw[2] = q.z; q_xyzw[3] = q.w;
}
void fo_euler2quat(double roll, double pitch, double yaw, double q_xyzw[4]) {
  Quat q = euler2Quaternion(roll, pitch, yaw); q_xyzw[0] = q.x; q_xyzw[1] = q.y; q_xyzw[2] = q.z; q_xyzw[3] = q.w;
}

// ---------- residuals / LM ----------
// residual record = 10 doubles: kind, curr(3), a(3), b(3)
static Residual unpack(const double* r) {
  return Residual{(int)r[0], Vec3{r[1], r[2], r[3]}, Vec3{r[4], r[5], r[6]}, Vec3{r[7], r[8], r[9]}};
}
int fo_evaluate_residual(const double rec[10], const double x[7], double* r, double jac7[7]) {
  return evaluate_residual(unpack(rec), x, r, jac7) ? 0 : 1;
}
// summary_out: iterations, accepted, initial_cost, final_cost, termination, H0[36], g0[6]  (47 doubles)
int fo_lm_solve(const double* recs, int n, int loss, double x[7], int max_iter, double* summary_out) {
  std::vector<Residual> blocks(n);
  for (int i = 0; i < n; ++i) blocks[i] = unpack(recs + 10 * i);
  LmSummary s;
  ceres_solve_pose(blocks, (LossKind)loss, x, &s, max_iter);
  if (summary_out) {
    summary_out[0] = s.iterations; summary_out[1] = s.accepted; summary_out[2] = s.initial_cost; summary_out[3] = s.final_cost;
    summary_out[4] = s.termination;
    std::memcpy(summary_out + 5, s.H0, sizeof(s.H0));
    std::memcpy(summary_out + 41, s.g0, sizeof(s.g0));
  }
  return 0;
}

// ---------- IMU / deskew ----------
void* fo_imu_create() { return new ImuHandle(); }
void fo_imu_destroy(void* h) { delete (ImuHandle*)h; }
void fo_imu_add(void* h, double stamp, const double q_xyzw[4]) { ((ImuHandle*)h)->h.AddMsg(stamp, Quat{q_xyzw[0], q_xyzw[1], q_xyzw[2], q_xyzw[3]}); }
int fo_imu_size(void* h) { return (int)((ImuHandle*)h)->h.size(); }
int fo_imu_get(void* h, double t, double q_xyzw[4]) {
  Quat q{0, 0, 0, 0};
  bool ok = ((ImuHandle*)h)->h.Get(t, q);
  q_xyzw[0] = q.x; q_xyzw[1] = q.y; q_xyzw[2] = q.z; q_xyzw[3] = q.w;
  return ok ? 1 : 0;
}
// CenterTime + Compensate + ImuAlign (src/laserProcessingNode.cpp:100-116). returns 0 ok, 1 = cannot compensate.
int fo_deskew_align(void* h, PointXYZIRT* pts, int n, unsigned long long* stamp_us, const double extr_xyzw[4]) {
  CloudIRT in(pts, pts + n), comp, aligned;
  std::uint64_t st = *stamp_us;
  CenterTime(in, st);
  *stamp_us = st;
  Quat ex{extr_xyzw[0], extr_xyzw[1], extr_xyzw[2], extr_xyzw[3]};
  if (!Compensate(in, st, comp, ((ImuHandle*)h)->h, ex)) { std::memcpy(pts, in.data(), sizeof(PointXYZIRT) * (size_t)n); return 1; }
  ImuAlign(comp, st, ((ImuHandle*)h)->h, ex, aligned);
  std::memcpy(pts, aligned.data(), sizeof(PointXYZIRT) * (size_t)n);
  return 0;
}
void fo_compensate_velocity(PointXYZIRT* pts, int n, const double v[3]) {
  CloudIRT c(pts, pts + n);
  CompensateVelocity(c, Vec3{v[0], v[1], v[2]});
  std::memcpy(pts, c.data(), sizeof(PointXYZIRT) * (size_t)n);
}

// ---------- OdomEstimationClass ----------
void* fo_odom_create(int num_lines, double scan_period, double min_dis, double max_dis, double map_resolution, const char* loss,
                     int total_order, int use_kdtree) {
  OdomHandle* h = new OdomHandle();
  LidarParam p; p.num_lines = num_lines; p.scan_period = scan_period; p.min_distance = min_dis; p.max_distance = max_dis;
  h->est.init(p, map_resolution, loss);
  h->est.total_order = total_order != 0;
  h->est.use_kdtree = use_kdtree == 1;   // 0 = brute force, 1 = FLANN-style kd-tree, 2 = 27-cell grid ((distance, index) order like 0, at kd-tree speed)
  h->est.use_grid = use_kdtree == 2;
  h->est.debug = &h->dbg;
  return h;
}
void fo_odom_destroy(void* h) { delete (OdomHandle*)h; }
// opt-in fixes (floam_fix bits 1 | 2) — restatement only: the reference build has no such modes
void fo_odom_set_fixes(void* h, int fixes) { ((OdomHandle*)h)->est.fixes = fixes; }
void fo_imu_set_slerp(void* h, int on) { ((ImuHandle*)h)->h.slerp = on != 0; }
void fo_odom_init_map(void* h, const PointXYZI* edge, int ne, const PointXYZI* surf, int ns) {
  ((OdomHandle*)h)->est.initMapWithPoints(CloudI(edge, edge + ne), CloudI(surf, surf + ns));
}
// mutates edge/surf in deskew mode like the reference (src/odomEstimationClass.cpp:42-43)
void fo_odom_update(void* h, PointXYZIRT* edge, int ne, PointXYZIRT* surf, int ns, int deskew, double pose_q_xyzw_t[7]) {
  OdomEstimation& est = ((OdomHandle*)h)->est;
  CloudIRT e(edge, edge + ne), s(surf, surf + ns);
  est.UpdatePointsToMapSelector(e, s, deskew != 0);
  if (deskew) { std::memcpy(edge, e.data(), sizeof(PointXYZIRT) * (size_t)ne); std::memcpy(surf, s.data(), sizeof(PointXYZIRT) * (size_t)ns); }
  if (pose_q_xyzw_t) std::memcpy(pose_q_xyzw_t, est.parameters, sizeof(double) * 7);
}
// single updatePointsToMap(PointXYZI) call, type: 0 VANILLA, 1 INITIAL_ITERATION, 2 REFINEMENT_AND_UPDATE
void fo_odom_update_xyzi(void* h, const PointXYZI* edge, int ne, const PointXYZI* surf, int ns, int type, double pose_q_xyzw_t[7]) {
  OdomEstimation& est = ((OdomHandle*)h)->est;
  est.updatePointsToMap(CloudI(edge, edge + ne), CloudI(surf, surf + ns), (OdomEstimation::UpdateType)type);
  if (pose_q_xyzw_t) std::memcpy(pose_q_xyzw_t, est.parameters, sizeof(double) * 7);
}
void fo_odom_get(void* h, double odom_rowmajor16[16], double last_odom16[16], double velocity[3], int* optimization_count) {
  OdomEstimation& est = ((OdomHandle*)h)->est;
  auto put = [](const Iso3& T, double* o) {
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) o[i * 4 + j] = T.R.m[i][j]; }
    o[3] = T.t.x; o[7] = T.t.y; o[11] = T.t.z; o[12] = o[13] = o[14] = 0; o[15] = 1;
  };
  if (odom_rowmajor16) put(est.odom, odom_rowmajor16);
  if (last_odom16) put(est.last_odom, last_odom16);
  if (velocity) { Vec3 v = est.GetVelocity(); velocity[0] = v.x; velocity[1] = v.y; velocity[2] = v.z; }
  if (optimization_count) *optimization_count = est.optimization_count;
}
// set pose state directly (stage-parity tests start both implementations from the same state)
void fo_odom_set_state(void* h, const double odom16[16], const double last_odom16[16], int optimization_count) {
  OdomEstimation& est = ((OdomHandle*)h)->est;
  auto get = [](const double* o, Iso3& T) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T.R.m[i][j] = o[i * 4 + j];
    T.t = Vec3{o[3], o[7], o[11]};
  };
  get(odom16, est.odom); get(last_odom16, est.last_odom);
  est.optimization_count = optimization_count;
}
void fo_odom_set_map(void* h, const PointXYZI* edge, int ne, const PointXYZI* surf, int ns) {
  OdomEstimation& est = ((OdomHandle*)h)->est;
  est.laserCloudCornerMap.assign(edge, edge + ne);
  est.laserCloudSurfMap.assign(surf, surf + ns);
}
int fo_odom_map_sizes(void* h, int* n_edge, int* n_surf) {
  OdomEstimation& est = ((OdomHandle*)h)->est;
  *n_edge = (int)est.laserCloudCornerMap.size(); *n_surf = (int)est.laserCloudSurfMap.size();
  return 0;
}
int fo_odom_get_map(void* h, PointXYZI* edge, int ecap, PointXYZI* surf, int scap) {
  OdomEstimation& est = ((OdomHandle*)h)->est;
  copy_out(est.laserCloudCornerMap, edge, ecap); copy_out(est.laserCloudSurfMap, surf, scap);
  return 0;
}
long fo_odom_knn_queries(void* h) { return ((OdomHandle*)h)->est.stat_knn_queries; }
// debug taps of the last updatePointsToMap call: what = 0 ds_edge, 1 ds_surf (PointXYZI); 2 edge_knn, 3 surf_knn (int x5);
// 4 edge_d2, 5 surf_d2 (float x5); 6 edge_ok, 7 surf_ok (u8); 8 residual records (10 doubles each); 9 LM summary (47 doubles);
// 10 scalars {outer_iterations, keyframe} (2 ints). Returns the element count.
int fo_odom_debug(void* h, int what, void* out, int cap) {
  OdomDebug& d = ((OdomHandle*)h)->dbg;
  switch (what) {
    case 0: return copy_out(d.ds_edge, (PointXYZI*)out, cap);
    case 1: return copy_out(d.ds_surf, (PointXYZI*)out, cap);
    case 2: return copy_out(d.edge_knn, (int*)out, cap);
    case 3: return copy_out(d.surf_knn, (int*)out, cap);
    case 4: return copy_out(d.edge_d2, (float*)out, cap);
    case 5: return copy_out(d.surf_d2, (float*)out, cap);
    case 6: return copy_out(d.edge_ok, (unsigned char*)out, cap);
    case 7: return copy_out(d.surf_ok, (unsigned char*)out, cap);
    case 8: {
      int n = (int)d.residuals.size();
      double* o = (double*)out;
      for (int i = 0; i < std::min(n, cap); ++i) {
        const Residual& r = d.residuals[i];
        double rec[10] = {(double)r.kind, r.curr_point.x, r.curr_point.y, r.curr_point.z, r.a.x, r.a.y, r.a.z, r.b.x, r.b.y, r.b.z};
        std::memcpy(o + 10 * i, rec, sizeof(rec));
      }
      return n;
    }
    case 9: {
      if (cap >= 47) {
        double* o = (double*)out;
        o[0] = d.lm.iterations; o[1] = d.lm.accepted; o[2] = d.lm.initial_cost; o[3] = d.lm.final_cost; o[4] = d.lm.termination;
        std::memcpy(o + 5, d.lm.H0, sizeof(d.lm.H0)); std::memcpy(o + 41, d.lm.g0, sizeof(d.lm.g0));
      }
      return 47;
    }
    case 10: {
      if (cap >= 2) { ((int*)out)[0] = d.outer_iterations; ((int*)out)[1] = d.keyframe ? 1 : 0; }
      return 2;
    }
  }
  return -1;
}

// ---------- LaserMappingClass ----------
void* fo_mapping_create(double map_resolution, int total_order) {
  LaserMapping* m = new LaserMapping();
  m->init(map_resolution);
  m->total_order = total_order != 0;
  return m;
}
void fo_mapping_destroy(void* m) { delete (LaserMapping*)m; }
void fo_mapping_update(void* m, const PointXYZI* pts, int n, const double pose16[16]) {
  Iso3 T;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T.R.m[i][j] = pose16[i * 4 + j];
  T.t = Vec3{pose16[3], pose16[7], pose16[11]};
  ((LaserMapping*)m)->updateCurrentPointsToMap(CloudI(pts, pts + n), T);
}
int fo_mapping_get_map(void* m, PointXYZI* out, int cap) { return copy_out(((LaserMapping*)m)->getMap(), out, cap); }

// ---------- timed whole-sequence replay (cpu_baseline / --impl reference): featureExtraction + odometry per frame ----------
// scans: concatenated frames, offsets[f]..offsets[f+1]. Returns seconds of wall-clock over the replayed frames
// (steady_clock, single thread like the reference's one worker thread per class). poses_out: 7 doubles per frame.
}  // extern "C"

static double replay_sequence_impl(const PointXYZIRT* scans, const long long* offsets, int n_frames, int num_lines, double scan_period, double min_dis,
                          double max_dis, double map_resolution, const char* loss, int deskew, double* poses_out, double* per_frame_ms,
                          long* knn_queries, int stage_skip, double* stage_ms_out) {
  LaserProcessing lp;
  LidarParam p; p.num_lines = num_lines; p.scan_period = scan_period; p.min_distance = min_dis; p.max_distance = max_dis;
  lp.init(p);
  OdomEstimation est;
  est.init(p, map_resolution, loss);
  bool inited = false;
  double total = 0, feature_s = 0;
  for (int f = 0; f < n_frames; ++f) {
    est.stage_timing = stage_ms_out && f >= stage_skip;
    auto t0 = std::chrono::steady_clock::now();
    CloudIRT in(scans + offsets[f], scans + offsets[f + 1]), e, s;
    lp.featureExtraction(in, e, s);
    if (est.stage_timing) feature_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!inited) {  // src/odomEstimationNode.cpp:219-224
      est.initMapWithPoints(VelToIntensityCopy(e), VelToIntensityCopy(s));
      inited = true;
    } else {
      est.UpdatePointsToMapSelector(e, s, deskew != 0);
    }
    auto t1 = std::chrono::steady_clock::now();
    double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    total += ms;
    if (per_frame_ms) per_frame_ms[f] = ms;
    if (poses_out) {
      Quat q = quat_from_matrix(est.odom.R);
      double* o = poses_out + 7 * f;
      o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w; o[4] = est.odom.t.x; o[5] = est.odom.t.y; o[6] = est.odom.t.z;
    }
  }
  if (knn_queries) *knn_queries = est.stat_knn_queries;
  if (stage_ms_out) {   // milliseconds per frame over frames [stage_skip, n_frames): features, downsample, kd build, association, solve, map update
    const double nf = n_frames > stage_skip ? (double)(n_frames - stage_skip) : 1.0;
    stage_ms_out[0] = feature_s * 1e3 / nf;
    for (int k = 0; k < 5; ++k) stage_ms_out[1 + k] = est.stage_s[k] * 1e3 / nf;
  }
  return total * 1e-3;
}

extern "C" {

double fo_replay_sequence(const PointXYZIRT* scans, const long long* offsets, int n_frames, int num_lines, double scan_period, double min_dis,
                          double max_dis, double map_resolution, const char* loss, int deskew, double* poses_out, double* per_frame_ms,
                          long* knn_queries) {
  return replay_sequence_impl(scans, offsets, n_frames, num_lines, scan_period, min_dis, max_dis, map_resolution, loss, deskew, poses_out, per_frame_ms, knn_queries, 0, nullptr);
}

double fo_replay_sequence_stages(const PointXYZIRT* scans, const long long* offsets, int n_frames, int num_lines, double scan_period, double min_dis,
                          double max_dis, double map_resolution, const char* loss, int deskew, double* poses_out, double* per_frame_ms,
                          long* knn_queries, int stage_skip, double* stage_ms_out) {
  return replay_sequence_impl(scans, offsets, n_frames, num_lines, scan_period, min_dis, max_dis, map_resolution, loss, deskew, poses_out, per_frame_ms, knn_queries, stage_skip, stage_ms_out);
}

}  // extern "C"

extern "C" void fo_householder_qr_solve(const double* A_rowmajor, const double* b, int rows, int cols, double* x) {
  std::vector<double> A((size_t)rows * cols), rhs(b, b + rows);
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) A[(size_t)c * rows + r] = A_rowmajor[(size_t)r * cols + c];
  householder_qr_solve(A.data(), rhs.data(), rows, cols, x);
}
