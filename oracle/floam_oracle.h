// ORACLE — TEST INFRASTRUCTURE ONLY (see types.h).  Pinned: tests/test_reference_pin.py runs this restatement against the reference's own
// sources compiled unmodified (oracle/_ref, `make ref`) on identical inputs — identical selections, bytes and poses.
//
// Declarations of the CPU restatement.  Structure follows the reference classes 1:1 so it can be diffed by eye
// (SURVEY.md §8c): LaserProcessing <-> src/laserProcessingClass.cpp, OdomEstimation <-> src/odomEstimationClass.cpp,
// cost functions <-> src/lidarOptimization.cpp, ImuHandler/Compensate <-> src/dataHandler.cpp,
// LaserMapping <-> src/laserMappingClass.cpp.  Library internals (PCL/FLANN/Eigen/Ceres) follow Appendix A.
#pragma once
#include "linalg.h"
#include "types.h"
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace fo {

typedef std::vector<PointXYZIRT> CloudIRT;
typedef std::vector<PointXYZI> CloudI;

struct LidarParam {  // include/lidar.h:53-86 ; defaults = code defaults of src/laserProcessingNode.cpp:175-179
  double max_distance = 60.0;
  double min_distance = 2.0;
  int num_lines = 64;
  double scan_period = 0.1;
  double vertical_angle = 2.0;
};

// ---- src/laserProcessingClass.cpp ----
struct FeatureStats {
  long curvature_ties = 0;  // adjacent equal curvature values after sorting (Q8 exposure)
  int rings_used = 0;
  int sectors = 0;
};
class LaserProcessing {
 public:
  void init(const LidarParam& p) { lidar_param = p; }
  // src/laserProcessingClass.cpp:72-118 ; appends to the caller's clouds like the reference.
  // stable_sort=false -> std::sort with the reference comparator (reference-faithful, Q8);
  // stable_sort=true  -> total order (value, id): the contract the CUDA path implements.
  void featureExtraction(const CloudIRT& pc_in, CloudIRT& pc_out_edge, CloudIRT& pc_out_surf, bool total_order = false,
                         FeatureStats* stats = nullptr, std::vector<int>* edge_src = nullptr,
                         std::vector<int>* surf_src = nullptr) const;
  LidarParam lidar_param;
};

// ---- PCL filters (Appendix A.1 / A.2) ----
// pcl::VoxelGrid<PointXYZI>::filter, leaf as given to setLeafSize (cast to float there).
// total_order=false: std::sort on idx only (PCL 1.8.1, unstable inside a voxel);
// total_order=true : stable (ascending point index inside a voxel) — the CUDA path's contract.
void voxel_grid_filter(const PointXYZI* in, size_t n, float leaf, CloudI& out, bool total_order = false, bool* passthrough = nullptr);
inline void voxel_grid_filter(const CloudI& in, float leaf, CloudI& out, bool total_order = false, bool* passthrough = nullptr) {
  voxel_grid_filter(in.data(), in.size(), leaf, out, total_order, passthrough);
}
// pcl::CropBox<PointXYZI>::filter, min/max already cast to float, negative=false.
void crop_box_filter(const PointXYZI* in, size_t n, const float mn[3], const float mx[3], CloudI& out);
inline void crop_box_filter(const CloudI& in, const float mn[3], const float mx[3], CloudI& out) { crop_box_filter(in.data(), in.size(), mn, mx, out); }

// ---- pcl::KdTreeFLANN<PointXYZI> (Appendix A.3) ----
class KdTreeFlann {
 public:
  void setInputCloud(const PointXYZI* cloud, size_t n);
  void setInputCloud(const CloudI& cloud) { setInputCloud(cloud.data(), cloud.size()); }
  // returns number of neighbours found (min(k, N)); ids/sqdist ascending
  int nearestKSearch(const PointXYZI& q, int k, int* ids, float* sqdist) const;
  size_t size() const { return n_; }

 private:
  struct Node {
    int left, right;  // leaf: point range
    int divfeat;
    float divlow, divhigh;
    int child1, child2;  // -1 for leaf
  };
  struct Interval {
    float low, high;
  };
  int divideTree(int left, int right, Interval bbox[3]);
  void middleSplit(int* ind, int count, int& index, int& cutfeat, float& cutval, const Interval bbox[3]);
  void planeSplit(int* ind, int count, int cutfeat, float cutval, int& lim1, int& lim2);
  struct ResultSet;
  void searchLevel(ResultSet& rs, const float* vec, int node, float mindistsq, float dists[3]) const;
  std::vector<float> pts_;      // original order, xyz
  std::vector<float> data_;     // reordered by vind_
  std::vector<int> vind_;
  std::vector<Node> nodes_;
  Interval root_bbox_[3];
  int root_ = -1;
  size_t n_ = 0;
};
// brute-force kNN with FLANN's L2_Simple float accumulation; ties by (distance, index). Ground truth for the build.
int knn_bruteforce(const PointXYZI* cloud, size_t n, const PointXYZI& q, int k, int* ids, float* sqdist);
inline int knn_bruteforce(const CloudI& cloud, const PointXYZI& q, int k, int* ids, float* sqdist) { return knn_bruteforce(cloud.data(), cloud.size(), q, k, ids, sqdist); }

// The same answer as knn_bruteforce for every query the reference uses (5th squared distance < 1.0, src/odomEstimationClass.cpp:154,207),
// at grid speed: points bucketed by 1 m cell, the 27 cells around the query scanned exhaustively with L2_Simple, (distance, index)
// order.  A point outside those cells is at least one cell width away along some axis, so its computed squared distance is >= 1.0
// and it can neither be nor displace one of five neighbours that pass the gate.  Queries that fail the gate report ids -1 and
// distances FLT_MAX (the reference never looks at them).  Checked against knn_bruteforce in tests/test_oracle.py.
class GridKnn {
 public:
  void setInputCloud(const PointXYZI* cloud, size_t n);
  void setInputCloud(const CloudI& cloud) { setInputCloud(cloud.data(), cloud.size()); }
  int nearestKSearch(const PointXYZI& q, int k, int* ids, float* sqdist) const;
 private:
  std::vector<float> pts_;        // xyz in cell order
  std::vector<int> index_;        // original index of pts_[i]
  std::vector<int> cell_start_;   // [ncells + 1]
  int ix0_ = 0, iy0_ = 0, iz0_ = 0, nx_ = 0, ny_ = 0, nz_ = 0;
  size_t n_ = 0;
};

// ---- src/lidarOptimization.cpp + Ceres (Appendix A.5) ----
enum LossKind { LOSS_TRIVIAL = 0, LOSS_HUBER = 1, LOSS_CAUCHY_TRUE = 2 };
struct Residual {   // one ceres residual block
  int kind;         // 0 = EdgeAnalyticCostFunction, 1 = SurfNormAnalyticCostFunction
  Vec3 curr_point;  // sensor frame
  Vec3 a, b;        // edge: last_point_a/b ; surf: a = plane_unit_norm, b.x = negative_OA_dot_norm
};
void getTransformFromSe3(const double se3[6], Quat& q, Vec3& t);          // src/lidarOptimization.cpp:101-137
void se3_plus(const double x[7], const double delta[6], double out[7]);   // PoseSE3Parameterization::Plus :77-91
// cost-function Evaluate: residual and (optionally) the 1x7 row-major global Jacobian. false on failure.
bool evaluate_residual(const Residual& rb, const double x[7], double* r, double* jac7);
struct LmSummary {
  int iterations = 0;       // step attempts
  int accepted = 0;
  double initial_cost = 0, final_cost = 0;
  int termination = 0;      // 0 max-iter, 1 param tol, 2 function tol, 3 gradient tol, 4 failure, 5 no residuals, 6 radius
  double H0[36]; double g0[6];  // J^T J and J^T r (unscaled) at the starting point, for stage parity
};
// What the trust-region loop needs from a problem with one 7-dof parameter block (local size 6).
struct LmProgram {
  virtual ~LmProgram() {}
  virtual size_t num_residuals() const = 0;
  // ProgramEvaluator::Evaluate: cost, corrected residuals (C), corrected local Jacobian (C x 6 row-major), gradient J^T r.
  // residuals / jacobian / gradient may be null (cost-only evaluation of a candidate).  false = evaluation failure.
  virtual bool evaluate(const double x[7], double* cost, std::vector<double>* residuals, std::vector<double>* jacobian, double gradient[6]) = 0;
  virtual void plus(const double x[7], const double delta[6], double x_plus_delta[7]) = 0;  // LocalParameterization::Plus
};
void trust_region_lm(LmProgram& program, double x[7], LmSummary* summary = nullptr, int max_num_iterations = 4);
// ceres::Solve(DENSE_QR, max_num_iterations=4) on one 7-dof block with PoseSE3Parameterization.
void ceres_solve_pose(const std::vector<Residual>& blocks, LossKind loss, double x[7], LmSummary* summary = nullptr,
                      int max_num_iterations = 4);

// ---- src/dataHandler.cpp ----
class ImuHandler {
 public:
  void AddMsg(double stamp, const Quat& orientation);            // :24-46
  bool Get(double tStamp, Quat& data) const;                     // :51-69
  Quat Get(double tStamp) const;                                 // :71-75 (default Imu = zero quaternion on failure)
  bool TimeContained(double t) const;                            // :76-81
  bool slerp = false;   // NOT the reference: opt-in fix FLOAM_FIX_IMU_SLERP (what Interpolate(tSlerp, ...) was meant to do, :48-50)
  size_t size() const { return data_.size(); }
  const std::vector<std::pair<double, Quat>>& data() const { return data_; }
 private:
  std::vector<std::pair<double, Quat>> data_;
};
// src/laserProcessingNode.cpp:65-78 ; stamp in microseconds (pcl header stamp)
void CenterTime(CloudIRT& cloud, std::uint64_t& stamp_us);
// src/dataHandler.cpp:93-122
bool Compensate(const CloudIRT& input, std::uint64_t stamp_us, CloudIRT& compensated, const ImuHandler& handler, const Quat& extrinsics);
// src/laserProcessingNode.cpp:113-116 : q = Imu(stamp)*extr ; pcl::transformPointCloud(compensated, aligned, Affine3d(q))
void ImuAlign(const CloudIRT& compensated, std::uint64_t stamp_us, const ImuHandler& handler, const Quat& extrinsics, CloudIRT& aligned);
// src/dataHandler.cpp:82-91
void CompensateVelocity(CloudIRT& input, Vec3 velocity);
// NOT the reference: opt-in fix FLOAM_FIX_ROTATED_VELOCITY (world-frame velocity rotated into the sensor frame first, Q14)
void CompensateVelocityRotated(CloudIRT& input, Vec3 velocity, const Mat3& R_world_sensor);

// ---- src/odomEstimationClass.cpp ----
CloudI VelToIntensityCopy(const CloudIRT& c);  // :308-318

struct OdomDebug {  // per updatePointsToMap call, last outer iteration (stage taps for parity tests)
  std::vector<int> edge_knn, surf_knn;         // 5 ids per query (-1 when map < k)
  std::vector<float> edge_d2, surf_d2;         // 5 squared distances per query
  std::vector<unsigned char> edge_ok, surf_ok; // accepted as residual
  std::vector<Residual> residuals;             // of the last outer iteration
  LmSummary lm;                                // of the last outer iteration
  CloudI ds_edge, ds_surf;
  int outer_iterations = 0;
  bool keyframe = false;
};

class OdomEstimation {
 public:
  enum UpdateType { VANILLA, INITIAL_ITERATION, REFINEMENT_AND_UPDATE };
  void init(const LidarParam& lidar_param, double map_resolution, const std::string& loss_function);   // :7-26
  void initMapWithPoints(const CloudI& edge_in, const CloudI& surf_in);                                // :28-32
  void UpdatePointsToMapSelector(CloudIRT& edge_in, CloudIRT& surf_in, bool deskew);                    // :34-50
  void updatePointsToMap(const CloudIRT& edge_in, const CloudIRT& surf_in, UpdateType t = VANILLA);     // :52-56
  void updatePointsToMap(const CloudI& edge_in, const CloudI& surf_in, UpdateType t = VANILLA);         // :57-124
  void getMap(CloudI& out) const;                                                                       // :296-300
  Vec3 GetVelocity() const { return (odom.t - last_odom.t) / lidar_param_.scan_period; }               // include/odomEstimationClass.h:78
  bool KeyFrameUpdate(const Iso3& pose);                                                                // :320-343

  Iso3 odom;
  CloudI laserCloudCornerMap, laserCloudSurfMap;
  // knobs that are not in the reference (test infrastructure)
  bool total_order = false;   // stable voxel order (see voxel_grid_filter)
  bool use_kdtree = true;     // false -> neighbours in (distance, index) order instead of the kd-tree's traversal order
  int fixes = 0;              // NOT the reference: opt-in fixes, same bits as floam_fix in include/floam_b200.h (1 single prediction, 2 rotated velocity)
  bool use_grid = false;      // with use_kdtree == false: GridKnn instead of the O(M) brute force (same results wherever the reference looks)
  OdomDebug* debug = nullptr;
  double parameters[7] = {0, 0, 0, 1, 0, 0, 0};
  Iso3 last_odom;
  int optimization_count = 2;
  long stat_knn_queries = 0;
  // wall-clock seconds per stage, accumulated while stage_timing is set (bench.py's per-stage CPU baseline):
  // [0] downSamplingToMap, [1] kd-tree builds, [2] addEdge/SurfCostFactor (kNN + fits), [3] ceres solve, [4] addPointsToMap
  bool stage_timing = false;
  double stage_s[5] = {0, 0, 0, 0, 0};

 private:
  void pointAssociateToMap(const PointXYZI& pi, PointXYZI& po) const;                                   // :126-135
  void downSamplingToMap(const CloudI& e_in, CloudI& e_out, const CloudI& s_in, CloudI& s_out) const;   // :137-142
  void addEdgeCostFactor(const CloudI& pc_in, const CloudI& map_in, std::vector<Residual>& problem, bool tap);   // :144-196
  void addSurfCostFactor(const CloudI& pc_in, const CloudI& map_in, std::vector<Residual>& problem, bool tap);   // :198-251
  void addPointsToMap(const CloudI& ds_edge, const CloudI& ds_surf);                                    // :253-294
  KdTreeFlann kdtreeEdgeMap, kdtreeSurfMap;
  GridKnn gridEdgeMap, gridSurfMap;
  float leaf_edge_ = 0.4f, leaf_surf_ = 0.8f;
  std::string loss_function_;
  LidarParam lidar_param_;
  const double keyframe_min_transl_ = 0.07;
  const double keyframe_min_rot_ = 2 * M_PI / 180.0;
  bool kf_first_ = true;  // Q10: function-static in the reference; per-instance here
  Iso3 kf_last_pose_;
};

// ---- src/laserMappingClass.cpp ----
class LaserMapping {
 public:
  void init(double map_resolution);                                         // :7-32
  void updateCurrentPointsToMap(const CloudI& pc_in, const Iso3& pose);     // :148-186
  CloudI getMap() const;                                                    // :188-200
  bool total_order = false;
 private:
  void checkPoints(int& x, int& y, int& z);                                 // :106-145
  typedef std::shared_ptr<CloudI> Cell;
  std::vector<std::vector<std::vector<Cell>>> map;
  int origin_in_map_x, origin_in_map_y, origin_in_map_z, map_width, map_height, map_depth;
  float leaf_;
};

}  // namespace fo
