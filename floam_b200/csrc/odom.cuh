// Subsystems (3) and (4): local maps with a uniform-grid exact 5-NN, fused line/plane fit, on-device Levenberg-Marquardt,
// the pose/keyframe state machine and the keyframe map update. See odom.cu.
#pragma once
#include "common.cuh"
#include "voxel.cuh"

namespace floam {

// Everything the per-frame state machine needs lives in one device struct so a frame needs no host decision.
struct PoseState {
  double odom[12];       // Isometry3d odom: R row-major (9) then t (3)   (include/odomEstimationClass.h:82)
  double last_odom[12];  // :94
  double x[7];           // parameters[7] = qx qy qz qw tx ty tz            (:90-92)
  double kf_pose[12];    // keyframes_.back().pose
  double velocity[3];    // GetVelocity() of the last update
  float crop_bounds[6];  // CropBox min xyz / max xyz (cast to float like Eigen::Vector4f(x_min, ...))
  int kf_first;          // Q10: first KeyFrameUpdate call returns true
  int not_keyframe;      // device-side skip flag for the map-update kernels
  int skip_solve;        // map too small ("not enough points in map to associate")
  int keyframe;          // result of the last KeyFrameUpdate
  // ---- Levenberg-Marquardt (Ceres trust-region loop) ----
  int lm_done, lm_phase, iteration, accepted, termination, last_successful, reuse_diag;
  double cost, H[21], g[6], scale[6], diag[6], radius, decrease_factor, x_norm, gmax, model_cost_change;
  double x_cand[7];
  double initial_cost, H0[21], g0[6];
  int lm_iterations_last, lm_accepted_last, lm_termination_last;
  double lm_final_cost_last;
  unsigned int ticket;   // last-CTA-done counter for the reduction
  int outer_iterations;
  int error_flags;       // bit 0: map capacity exceeded, bit 1: grid capacity exceeded — of the CURRENT frame: cleared once mailed
  int error_sticky;      // OR of every mailed frame's flags (bits 0-1 as above, bits 4-5 the feature-extraction flags word)
  int n_corr;            // accepted correspondences of the last association
  int n_corr_acc;        // accumulator of the running association
  int frame_counter;     // frames completed (index of the next trajectory record)
  int map_points[2];     // edge / surf local map sizes at the start of the last update (the host's size hints come from here)
  // development aid (FLOAM_DBG_TIMELINE): globaltimer stamps of the BACK half — predict start, finish start, start of the two grid
  // scatters (its last kernels) — and their running sums: [0] BACK duration, [1] gap between consecutive BACKs, [2] solve part, [3] frames
  long long tl_predict, tl_finish, tl_end[2], tl_sum[4];
  long long dbg_clk[8];  // clock64 stamps of the last lm_cluster_kernel attempt (development aid, FLOAM_DBG_CLOCKS)
};

struct GridDims {
  int ix0, iy0, iz0;  // cell coordinate of the grid origin
  int nx, ny, nz;
  int ncells;
};

struct LocalMap {
  P4* pts;          // the map cloud (reference order: ascending voxel index after every keyframe)
  int* d_n;
  P4* tmp;          // scratch (append -> crop -> voxel)
  int *d_ntmp, *d_ncrop;
  float4* cell_pts; // points bucketed by 1 m cell: xyz + map index (as int bits in w)
  int* cell_start;  // [ncells_cap + 1]
  int* cell_count;  // [ncells_cap]
  int* tile_sums;   // points per kScanTile cells: filled by the count kernel, consumed by the scan, zeroed again by the scatter kernel
  GridDims* dims;
  unsigned int* bbox;
  int* d_ncells;
  int cap, ncells_cap;
};

constexpr int kLmTerms = 29;  // 21 H (upper triangle) + 6 g + cost + number of correspondences

struct OdomDevice;
// which filter the next keyframe update of the surf (k = 0) / edge (k = 1) map will be enqueued with: part of the frame graph's key
bool odom_map_update_merges(const OdomDevice& od, int k);

struct OdomDevice {
  PoseState* state;
  LocalMap edge_map, surf_map;
  P4 *ds_edge, *ds_surf;       // downsampled current features (sensor frame): aliases of the buffer set of the frame being enqueued
  int *d_nds_edge, *d_nds_surf;
  P4 *ds_edge_b[2], *ds_surf_b[2];     // two sets: frame k+1 is downsampled while frame k is still being solved / merged into the map
  int *d_nds_edge_b[2], *d_nds_surf_b[2];
  int qcap;                    // capacity of each downsampled cloud
  // correspondences, slot = query index (edge slots [0,qcap), surf slots [qcap, 2*qcap))
  double* corr;                // [6][2*qcap]: edge a(3) b(3); surf n(3) d
  unsigned char* corr_ok;      // [2*qcap]
  int* knn_ids;                // [2*qcap*5] debug taps / floam_knn5
  float* knn_d2;               // [2*qcap*5]
  float4* knn_q;               // [2*qcap] query position of the last search + lower bound of the distance to every other map point
  double* partials;            // [CTAs of the association kernel][kLmTerms]
  VoxelWorkspace* vws;         // main-branch workspace (surf side)
  VoxelWorkspace* vws_aux;     // second workspace: the edge side runs as a parallel branch of the frame graph
  cudaStream_t aux_stream;     // fork/join partner of the context stream
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int lm_cluster_ctas = 8;     // CTAs of the solve's thread-block cluster (8, or 16 for dense configurations)
  // Keyframe map update: full re-sort of map + new points, or classify + sort of the out-of-place points + merge (voxel.cu). Same
  // maps; the merge wins from ~250k map points on (tools/probes/map_merge_sweep.py), the re-sort below. 0 = re-sort, 1 = by the
  // host's size hint (default), 2 = merge always. FLOAM_MAP_MERGE / floam_set_map_merge.
  int map_merge_mode = 1;
  int map_merge_min_points = 250000;            // FLOAM_MAP_MERGE_MIN
  int edge_map_hint = -1, surf_map_hint = -1;   // host-side guesses of the map sizes the next keyframe update will filter (-1: unknown)
  bool knn_staged = true;      // association kNN brings sparse neighbourhoods into shared memory by bulk copies (cp.async.bulk + mbarrier)
  float leaf_edge, leaf_surf;
  double scan_period;
  int loss;
  int fixes;                   // floam_fix bits (opt-in deviations from the reference, default 0)
  int optimization_count;      // host mirror (deterministic schedule, Q4)
  double* traj;                // [traj_cap][7] device-side trajectory log (pose of every completed frame)
  int traj_cap;
};

int local_map_alloc(LocalMap& map, int cap, int ncells_cap, void* (*alloc)(void*, size_t), void* alloc_ctx, cudaStream_t s);
int odom_device_init(OdomDevice& od, const floam_params& prm, VoxelWorkspace* vws, VoxelWorkspace* vws_aux, cudaStream_t aux, void* (*alloc)(void*, size_t),
                     void* alloc_ctx, cudaStream_t s);
void odom_reset_state(OdomDevice& od, cudaStream_t s);
// appends the current pose to the device trajectory log (first frame: the update path does it in its finish kernel)
void odom_record_pose(OdomDevice& od, cudaStream_t s);
// End-of-frame mailbox: the state (and the feature-extraction flags word) written straight into pinned, device-mapped host memory by
// a kernel — posted stores instead of two copy-engine nodes at the tail of the frame graph.
void odom_mail_state(OdomDevice& od, const int* d_flags, PoseState* h_state, int* h_flags, cudaStream_t s);   // also clears the frame's error flags

// append (replace = 0) or overwrite (replace = 1) a map with a strided device cloud and rebuild its grid
void local_map_load(OdomDevice& od, LocalMap& map, const void* d_pts, const int* d_n, int stride, int n_max, int replace, cudaStream_t s);
// OdomEstimationClass::initMapWithPoints: append (stride-32 or stride-16 clouds), optimization_count = 12, rebuild grids
void odom_init_map_device(OdomDevice& od, const void* d_edge, const int* d_ne, const void* d_surf, const int* d_ns, int stride, int n_max, int replace,
                          cudaStream_t s);
// OdomEstimationClass::updatePointsToMap on device-resident clouds (stride 32: PointXYZI / PointXYZIRT, stride 16: float4).
// tap != 0: the last outer iteration also writes the kNN ids / distances for floam_debug_fetch.
void odom_update_device(OdomDevice& od, const void* d_edge, const int* d_ne, const void* d_surf, const int* d_ns, int stride, int n_max, int update_type,
                        int ds_ready, cudaStream_t s);
// downSamplingToMap (:137-142) on its own: edge cloud on `aux` with ws_edge, surf cloud on `s` with ws_surf (fork / join through the
// two events). Used by the frame pipeline to downsample frame k+1 while frame k is still in its solve / map update.
void odom_downsample_device(OdomDevice& od, const void* d_edge, const int* d_ne, const void* d_surf, const int* d_ns, int stride, int n_max,
                            VoxelWorkspace& ws_surf, VoxelWorkspace& ws_edge, cudaStream_t s, cudaStream_t aux, cudaEvent_t ev_fork, cudaEvent_t ev_join);
void odom_select_buffers(OdomDevice& od, int parity);
// dmapping::CompensateVelocity with GetVelocity() read from the device state
void compensate_velocity_device(OdomDevice& od, PointIRT* d_pts, const int* d_n, int n_max, cudaStream_t s);
void compensate_velocity_explicit_device(PointIRT* d_pts, const int* d_n, int n_max, const double v[3], cudaStream_t s);
void odom_rebuild_grids(OdomDevice& od, cudaStream_t s);
// stand-alone exact 5-NN of raw queries against a map whose grid is built (floam_knn5)
void knn5_device(OdomDevice& od, LocalMap& map, const P4* d_queries, const int* d_nq, int nq_max, int* d_ids, float* d_d2, cudaStream_t s);

}  // namespace floam
