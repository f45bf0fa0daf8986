// Subsystems (3) and (4) plus the pose / keyframe state machine of OdomEstimationClass, all device-resident.
//
//   reference                                            here
//   ---------------------------------------------------  ------------------------------------------------------------
//   updatePointsToMap :57-71 (prediction, Q2)            predict_kernel (1 thread)
//   downSamplingToMap :137-142                           voxel_grid_device x2 (voxel.cu)
//   kdtree setInputCloud :78-79 (rebuilt every call)     1 m uniform grid, rebuilt only when the map changed (Q12):
//                                                         grid_bbox -> grid_dims -> grid_count -> scan -> grid_scatter
//   addEdgeCostFactor/addSurfCostFactor :144-251         assoc_eval_kernel: pointAssociateToMap, exact 5-NN over the 27
//                                                         neighbouring cells with (distance, index) order, PCA line fit /
//                                                         5x3 QR plane fit, and the iteration-0 residual+Jacobian reduction
//   ceres::Solve :100-108 (LM, <= 4 step attempts)       lm_cluster_kernel: one 8-CTA cluster evaluates each candidate pose and
//                                                         CTA 0 runs the trust-region bookkeeping (accept / reject, radius law,
//                                                         tolerances, next step) between two cluster barriers, on the device
//   odom write-back, KeyFrameUpdate :114-118,320-343     finish_kernel (1 thread)
//   addPointsToMap :253-294                              append_kernel -> crop_box_device -> voxel_grid_device -> grid rebuild,
//                                                         all predicated on the device-side keyframe flag (no host sync)
//
// Compiled with -fmad=false: the float distance / pointAssociateToMap arithmetic must round exactly like the reference's
// non-contracted x86 code so that neighbour ids are bit-exact.
#include "odom.cuh"

#include <cfloat>
#include <cstddef>
#include <cstdlib>

#include <cooperative_groups.h>

#include "odom_math.cuh"

namespace floam {
namespace {

constexpr int kThreads = 256;
constexpr int kAssocBlocks = kNumSMs / 2;   // 74 CTAs x 128 threads, grid-stride over the query slots: one partial row per CTA
constexpr int kKnnBlocks = kNumSMs * 4;     // 592 CTAs x 8 warps, one warp per query, grid-stride
static_assert(kAssocBlocks <= 1024, "partials rows");

// the kernels stride; two CTAs per SM unless the input is known to be large (see voxel.cu, grid_for)
inline int grid_for(int n_max, int ctas_per_sm = 2) {
  int g = (n_max + kThreads - 1) / kThreads;
  const int cap = kNumSMs * ctas_per_sm;
  return g < 1 ? 1 : (g > cap ? cap : g);
}

__device__ __forceinline__ float4 load_xyzi(const char* base, int stride, int i) {
  const char* p = base + (size_t)i * stride;
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  if (stride == 16) return a;
  return make_float4(a.x, a.y, a.z, __ldg(reinterpret_cast<const float*>(p + 16)));
}

// ------------------------------------------------------------------------------------------------------------------
// state machine: prediction and write-back
// ------------------------------------------------------------------------------------------------------------------
__global__ void state_init_kernel(PoseState* S) {
  pdl_prologue();
  if (threadIdx.x != 0) return;
  PoseState z;
  memset(&z, 0, sizeof(z));
  z.odom[0] = z.odom[4] = z.odom[8] = 1.0;
  z.last_odom[0] = z.last_odom[4] = z.last_odom[8] = 1.0;
  z.kf_pose[0] = z.kf_pose[4] = z.kf_pose[8] = 1.0;
  z.x[3] = 1.0;
  z.kf_first = 1;
  z.not_keyframe = 1;
  *S = z;
}

// updatePointsToMap :62-71. The motion prediction is unconditional (Q2).
__device__ __forceinline__ long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}

__global__ void predict_kernel(PoseState* S, const int* n_edge_map, const int* n_surf_map, int keep_pose) {
  pdl_prologue();
  if (threadIdx.x != 0) return;
  S->map_points[0] = *n_edge_map; S->map_points[1] = *n_surf_map;
  {
    const long long now = global_ns();
    if (S->tl_predict != 0 && S->tl_end[0] != 0) {
      const long long end = S->tl_end[0] > S->tl_end[1] ? S->tl_end[0] : S->tl_end[1];
      S->tl_sum[0] += end - S->tl_predict; S->tl_sum[1] += now - end; S->tl_sum[2] += S->tl_finish - S->tl_predict; S->tl_sum[3] += 1;
    }
    S->tl_predict = now; S->tl_end[0] = S->tl_end[1] = 0;
  }
  double inv[12], rel[12], pred[12];
  if (keep_pose) {   // FLOAM_FIX_SINGLE_PREDICTION, second deskew pass: odom is the pass-1 registration, last_odom the previous frame's pose (:66)
    for (int i = 0; i < 12; ++i) pred[i] = S->odom[i];
  } else {
    m::iso_inverse(S->last_odom, inv);
    m::iso_mul(inv, S->odom, rel);
    m::iso_mul(S->odom, rel, pred);
    for (int i = 0; i < 12; ++i) { S->last_odom[i] = S->odom[i]; S->odom[i] = pred[i]; }
  }
  double q[4];
  m::quat_from_matrix(pred, q);
  S->x[0] = q[0]; S->x[1] = q[1]; S->x[2] = q[2]; S->x[3] = q[3];
  S->x[4] = pred[9]; S->x[5] = pred[10]; S->x[6] = pred[11];
  S->skip_solve = (*n_edge_map > 10 && *n_surf_map > 50) ? 0 : 1;  // :77 "not enough points in map to associate"
  S->outer_iterations = 0;
  S->not_keyframe = 1;
  S->keyframe = 0;
  S->lm_iterations_last = 0; S->lm_accepted_last = 0; S->lm_termination_last = 5; S->lm_final_cost_last = 0.0; S->initial_cost = 0.0;
  for (int i = 0; i < 21; ++i) S->H0[i] = 0.0;
  for (int i = 0; i < 6; ++i) S->g0[i] = 0.0;
  S->n_corr = 0;
}

// :114-121 + KeyFrameUpdate :320-343 + the CropBox bounds of addPointsToMap :270-279
// One thread. S may be the global state or CTA 0's shared-memory copy inside the cluster kernel (generic pointer).
struct FinishArgs {
  int enabled;       // the cluster kernel of the last outer iteration runs the write-back itself (no separate launch)
  int update_type;
  double scan_period;
  double* traj;
  int traj_cap;
};
__device__ __noinline__ void finish_body(PoseState* S, int update_type, double scan_period, double* traj, int traj_cap) {
  S->tl_finish = global_ns();
  double R[9];
  m::quat_to_matrix(S->x, R);
  for (int i = 0; i < 9; ++i) S->odom[i] = R[i];
  S->odom[9] = S->x[4]; S->odom[10] = S->x[5]; S->odom[11] = S->x[6];
  for (int a = 0; a < 3; ++a) S->velocity[a] = (S->odom[9 + a] - S->last_odom[9 + a]) / scan_period;  // GetVelocity()
  int kf = 0;
  if (update_type == FLOAM_VANILLA || update_type == FLOAM_REFINEMENT_AND_UPDATE) {
    if (S->kf_first) {
      S->kf_first = 0;
      kf = 1;
    } else {
      double inv[12], delta[12];
      m::iso_inverse(S->kf_pose, inv);
      m::iso_mul(inv, S->odom, delta);
      const double mov = sqrt(delta[9] * delta[9] + delta[10] * delta[10] + delta[11] * delta[11]);
      const double rot = m::rotation_angle(delta);
      if (mov > 0.07 || rot > 2.0 * M_PI / 180.0) kf = 1;
    }
    if (kf) {
      for (int i = 0; i < 12; ++i) S->kf_pose[i] = S->odom[i];
      for (int a = 0; a < 3; ++a) {
        S->crop_bounds[a] = (float)(S->odom[9 + a] - 100.0);
        S->crop_bounds[3 + a] = (float)(S->odom[9 + a] + 100.0);
      }
    }
  }
  S->keyframe = kf;
  S->not_keyframe = kf ? 0 : 1;
  if (update_type != FLOAM_INITIAL_ITERATION) {
    double* rec = traj + (size_t)(S->frame_counter % traj_cap) * 7;
    for (int k = 0; k < 7; ++k) rec[k] = S->x[k];
    S->frame_counter++;
  }
}
__global__ void finish_kernel(PoseState* S, int update_type, double scan_period, double* traj, int traj_cap) {
  pdl_prologue();
  if (threadIdx.x != 0) return;
  finish_body(S, update_type, scan_period, traj, traj_cap);
}

__global__ void record_pose_kernel(PoseState* S, double* traj, int traj_cap) {
  pdl_prologue();
  if (threadIdx.x != 0) return;
  double* rec = traj + (size_t)(S->frame_counter % traj_cap) * 7;
  for (int k = 0; k < 7; ++k) rec[k] = S->x[k];
  S->frame_counter++;
}

// ------------------------------------------------------------------------------------------------------------------
// local map: append, uniform 1 m grid
// ------------------------------------------------------------------------------------------------------------------
// initMapWithPoints :28-32 / set_map: raw append of a strided cloud (no transform)
__global__ void __launch_bounds__(kThreads) map_append_raw_kernel(const char* __restrict__ in, int stride, const int* __restrict__ d_nin, P4* __restrict__ map,
                                                                   int* d_nmap, int cap, int replace, int* d_err) {
  pdl_prologue();
  const int nin = *d_nin;
  const int base = replace ? 0 : *d_nmap;
  const bool fits = base + nin <= cap;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nin && fits; i += gridDim.x * kThreads) map[base + i] = load_xyzi(in, stride, i);
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0 && !fits) atomicOr(d_err, 1);
}
__global__ void map_bump_kernel(int* d_nmap, const int* d_nin, int cap, int replace, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  if (threadIdx.x != 0) return;
  const int base = replace ? 0 : *d_nmap;
  if (base + *d_nin <= cap) *d_nmap = base + *d_nin;
}

__device__ __forceinline__ void bbox_accumulate(float (&mn)[3], float (&mx)[3], bool any, unsigned int* __restrict__ bbox) {
  if (!__any_sync(0xffffffffu, any)) return;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  }
  if (lane_id() == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&bbox[a], float_flip(mn[a]));
      atomicMax(&bbox[3 + a], float_flip(mx[a]));
    }
  }
}
__global__ void __launch_bounds__(kThreads) grid_bbox_kernel(const P4* __restrict__ pts, const int* __restrict__ d_n, unsigned int* __restrict__ bbox,
                                                              const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  const int n = *d_n;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  bool any = false;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const float4 p = __ldg(pts + i);
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    any = true;
  }
  bbox_accumulate(mn, mx, any, bbox);
}

// grid extent from the bounding box (every CTA of the count kernel derives it by itself; CTA 0 also stores it for the kernels after)
__device__ __forceinline__ GridDims grid_dims_from_bbox(const unsigned int* bbox, int n, int ncells_cap, bool* overflow) {
  GridDims g;
  g.ix0 = g.iy0 = g.iz0 = 0; g.nx = g.ny = g.nz = 0; g.ncells = 0;
  *overflow = false;
  if (n > 0) {
    const float mnx = float_unflip(bbox[0]), mny = float_unflip(bbox[1]), mnz = float_unflip(bbox[2]);
    const float mxx = float_unflip(bbox[3]), mxy = float_unflip(bbox[4]), mxz = float_unflip(bbox[5]);
    g.ix0 = (int)floorf(mnx); g.iy0 = (int)floorf(mny); g.iz0 = (int)floorf(mnz);
    const long long nx = (long long)floorf(mxx) - g.ix0 + 1, ny = (long long)floorf(mxy) - g.iy0 + 1, nz = (long long)floorf(mxz) - g.iz0 + 1;
    if (nx * ny * nz <= (long long)ncells_cap) {
      g.nx = (int)nx; g.ny = (int)ny; g.nz = (int)nz; g.ncells = (int)(nx * ny * nz);
    } else {
      *overflow = true;  // grid capacity exceeded: every query of this map is rejected
    }
  }
  return g;
}

__device__ __forceinline__ int cell_of(const GridDims& g, float x, float y, float z) {
  const int cx = (int)floorf(x) - g.ix0, cy = (int)floorf(y) - g.iy0, cz = (int)floorf(z) - g.iz0;
  return cx + g.nx * (cy + g.ny * cz);
}

__global__ void __launch_bounds__(kThreads) grid_count_kernel(const P4* __restrict__ pts, const int* __restrict__ d_n, const unsigned int* __restrict__ bbox,
                                                               GridDims* __restrict__ dims, int ncells_cap, PoseState* S, int* __restrict__ cell_count,
                                                               int* __restrict__ tile_sums, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  static_assert(kScanTile == 4096, "tile_sums index is cell >> 12");
  const int n = *d_n;
  bool overflow;
  const GridDims g = grid_dims_from_bbox(bbox, n, ncells_cap, &overflow);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *dims = g;
    if (overflow) atomicOr(&S->error_flags, 2);
  }
  if (g.ncells == 0) return;
  // the scan's per-tile sums come for free here; they are gathered per CTA in shared memory first (a map spans a handful of 4096-cell
  // tiles, and every warp of the grid adding to the same few global words was most of this kernel on dense maps)
  constexpr int kTileBins = 2048;
  __shared__ int s_tiles[kTileBins];
  const int ntiles = (g.ncells + kScanTile - 1) / kScanTile;
  const bool binned = ntiles <= kTileBins;
  if (binned) {
    for (int t = threadIdx.x; t < ntiles; t += kThreads) s_tiles[t] = 0;
    __syncthreads();
  }
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const float4 p = __ldg(pts + i);
    const int c = cell_of(g, p.x, p.y, p.z);
    // one atomic per group of lanes that hit the same cell: the cloud is in voxel order, so neighbouring lanes mostly share a 1 m
    // cell (dense maps: a dozen voxels per cell edge) and per-point atomics would queue on one address
    const unsigned int active = __activemask();
    const unsigned int same = __match_any_sync(active, c);
    if (lane_id() == __ffs(same) - 1) atomicAdd(&cell_count[c], __popc(same));
    const int t = c >> 12;
    const unsigned int peers = __match_any_sync(active, t);
    if (lane_id() == __ffs(peers) - 1) atomicAdd(binned ? &s_tiles[t] : &tile_sums[t], __popc(peers));
  }
  if (binned) {
    __syncthreads();
    for (int t = threadIdx.x; t < ntiles; t += kThreads)
      if (s_tiles[t]) atomicAdd(&tile_sums[t], s_tiles[t]);
  }
}

// cell_count doubles as the fill cursor: it is counted down to zero here, which is also the state the next build expects.
// The order of points inside a cell is arbitrary; the search orders candidates by (distance, index), so results do not depend on it.
__global__ void __launch_bounds__(kThreads) grid_scatter_kernel(const P4* __restrict__ pts, const int* __restrict__ d_n, const GridDims* __restrict__ dims,
                                                                 const int* __restrict__ cell_start, int* __restrict__ cell_count,
                                                                 float4* __restrict__ cell_pts, unsigned int* bbox, int* __restrict__ tile_sums,
                                                                 P4* __restrict__ home, long long* stamp, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  if (stamp && blockIdx.x == 0 && threadIdx.x == 0) *stamp = global_ns();
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // every reader of the bounding box (grid_count_kernel) is done: re-arm it for the next build
    bbox[0] = bbox[1] = bbox[2] = 0xffffffffu;
    bbox[3] = bbox[4] = bbox[5] = 0u;
  }
  const GridDims g = *dims;
  const int n = *d_n;
  // the scan before this kernel was the only reader of the tile sums: back to zero for the next build
  for (int t = blockIdx.x * kThreads + threadIdx.x; t * kScanTile < g.ncells; t += gridDim.x * kThreads) tile_sums[t] = 0;
  // home (optional): pts is the filter's output buffer; every point passes through here, so this is also where the cloud goes back
  // into the map's home buffer (a device-side pointer swap would need every consumer to chase a pointer)
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const float4 p = __ldg(pts + i);
    if (home) home[i] = p;
    if (g.ncells == 0) continue;
    const int c = cell_of(g, p.x, p.y, p.z);
    // lanes of the same cell take their slots from one atomic (see grid_count_kernel)
    const unsigned int same = __match_any_sync(__activemask(), c);
    const int leader = __ffs(same) - 1;
    int top = 0;
    if (lane_id() == leader) top = atomicSub(&cell_count[c], __popc(same));
    top = __shfl_sync(same, top, leader);
    const int pos = cell_start[c] + top - 1 - __popc(same & ((1u << lane_id()) - 1u));
    cell_pts[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// exact 5-NN over the 27 cells around the query (pcl::KdTreeFLANN::nearestKSearch, k = 5, for d5^2 < 1)
// ------------------------------------------------------------------------------------------------------------------
struct Knn5 {
  float d[5];
  int id[5];
};

// rej: smallest distance this lane has seen and does NOT hold in its list (rejected candidates and evicted entries); with the heads
// left over after the merge it yields the 6th distance, which is what lets the next outer iteration reuse the neighbour set
__device__ __forceinline__ void knn5_insert(Knn5& k, float dist, int idx, float& rej) {
  if (dist > k.d[4] || (dist == k.d[4] && idx > k.id[4])) { rej = fminf(rej, dist); return; }
  rej = fminf(rej, k.d[4]);
  // replace the current worst, then bubble up by (distance, index); static indices keep the set in registers
  k.d[4] = dist; k.id[4] = idx;
#pragma unroll
  for (int j = 4; j >= 1; --j) {
    const bool lt = k.d[j] < k.d[j - 1] || (k.d[j] == k.d[j - 1] && k.id[j] < k.id[j - 1]);
    if (lt) {
      const float td = k.d[j]; k.d[j] = k.d[j - 1]; k.d[j - 1] = td;
      const int ti = k.id[j]; k.id[j] = k.id[j - 1]; k.id[j - 1] = ti;
    }
  }
}

// Upper bound of the warp-wide 5th smallest distance seen so far: the private lists are ascending, so five rounds of
// {redux.min over the list heads, lowest lane holding the minimum pops} walk the five smallest distances of the union.
__device__ __forceinline__ float knn5_warp_bound(const Knn5& k) {
  const int l = lane_id();
  unsigned int v0 = __float_as_uint(k.d[0]), v1 = __float_as_uint(k.d[1]), v2 = __float_as_uint(k.d[2]), v3 = __float_as_uint(k.d[3]),
               v4 = __float_as_uint(k.d[4]);   // distances are >= 0: the bit patterns order like the floats
  unsigned int m = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    m = __reduce_min_sync(0xffffffffu, v0);
    const unsigned int who = __ballot_sync(0xffffffffu, v0 == m);
    if (l == __ffs(who) - 1) { v0 = v1; v1 = v2; v2 = v3; v3 = v4; v4 = __float_as_uint(FLT_MAX); }
  }
  return __uint_as_float(m);
}

__device__ __forceinline__ void knn5_scan_range(int b, int e, float qx, float qy, float qz, float bound, const float4* __restrict__ cell_pts, Knn5& k,
                                                float& rej) {
  for (int i = b + lane_id(); i < e; i += 32) {
    const float4 p = __ldg(cell_pts + i);
    // flann::L2_Simple: ((0 + dx^2) + dy^2) + dz^2 in float, no contraction
    const float dx = fsub(qx, p.x), dy = fsub(qy, p.y), dz = fsub(qz, p.z);
    const float dist = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
    if (dist <= bound) knn5_insert(k, dist, __float_as_int(p.w), rej);
    else rej = fminf(rej, dist);
  }
}

// The same search by a whole warp for ONE query. Lanes 0..8 load the bounds of the nine x-contiguous cell runs, lanes 0..26 the
// bounds of the 27 cells; lanes stride over candidates keeping a private top-5, and five rounds of warp arg-min over the packed
// (distance, index) heads merge the 32 lists. Every lane returns the final set.
//  * few candidates (sparse maps, the usual case): the nine runs are scanned whole.
//  * many candidates (dense maps): the query's own cell is scanned first and gives an upper bound of the 5th distance; a
//    neighbouring cell is skipped when the distance from the query to the cell, computed with the SAME rounded operations as a
//    candidate distance (per axis q - floor(q) or floor(q) + 1 - q, or 0), exceeds the bound. Rounding is monotonic, so every
//    candidate of a skipped cell has a computed distance >= that figure > bound >= the final 5th distance: the result is the
//    same set, ties included, as the exhaustive scan.
constexpr int kKnnPruneAbove = 256;   // candidates in the 27 cells

// ---- TMA-staged cell tiles (north_star (3)) ----
// Sparse neighbourhoods (<= kKnnPruneAbove candidates, the usual case at the default map resolution): instead of nine dependent
// scan loops over the nine x-contiguous runs — each one an L2 round trip the warp sits out before it can issue the next — the nine
// runs are brought into a per-warp shared-memory slab by up to nine bulk copies (cp.async.bulk global -> shared, lanes 0..8 issue one
// each, completion counted in bytes on the warp's mbarrier) that are all in flight together; the lanes then read the candidates
// from shared memory. One round trip instead of nine, same candidate set, same (distance, index) result.
struct KnnStage {
  float4* slab;            // this warp's kKnnPruneAbove entries
  unsigned int bar;        // shared-space address of this warp's mbarrier
  unsigned int phase;      // parity of the next completion
};
__device__ __forceinline__ bool mbar_try_wait(unsigned int bar, unsigned int parity) {
  unsigned int ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void knn5_scan_staged(int rb, int re, int total, float qx, float qy, float qz, const float4* __restrict__ cell_pts, KnnStage& st,
                                                 Knn5& k, float& rej) {
  const int l = lane_id();
  const int len = re - rb;          // lanes >= 9 hold 0
  int incl = len;
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (l >= o) incl += t;
  }
  const int excl = incl - len;
  // the previous query's reads of the slab (generic proxy) are done before the async proxy writes it again
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (l == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(st.bar), "r"((unsigned int)(total * 16)) : "memory");
  __syncwarp();
  if (len > 0) {
    const unsigned int dst = (unsigned int)__cvta_generic_to_shared(st.slab + excl);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(cell_pts + rb), "r"((unsigned int)(len * 16)), "r"(st.bar) : "memory");
  }
  while (!mbar_try_wait(st.bar, st.phase)) {}
  st.phase ^= 1u;
  for (int i = l; i < total; i += 32) {
    const float4 p = st.slab[i];
    const float dx = fsub(qx, p.x), dy = fsub(qy, p.y), dz = fsub(qz, p.z);
    const float dist = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
    knn5_insert(k, dist, __float_as_int(p.w), rej);
  }
}

// others_min: lower bound of the computed distance of every map point that is inside the 27 cells and NOT in the result.
template <bool kStaged>
__device__ __forceinline__ void knn5_search_warp(const GridDims& g, const int* __restrict__ cell_start, const float4* __restrict__ cell_pts, float qx,
                                                 float qy, float qz, Knn5& out, float& others_min, KnnStage* st = nullptr) {
  const int l = lane_id();
  float rej = FLT_MAX, skipped = FLT_MAX;
  Knn5 k;
#pragma unroll
  for (int j = 0; j < 5; ++j) { k.d[j] = FLT_MAX; k.id[j] = 0x7fffffff; }
  const float flx = floorf(qx), fly = floorf(qy), flz = floorf(qz);
  const int cx = (int)flx - g.ix0, cy = (int)fly - g.iy0, cz = (int)flz - g.iz0;   // float -> int saturates: far-away queries find no cell
  const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
  int rb = 0, re = 0;  // lane r < 9 owns run r = (dz+1)*3 + (dy+1)
  if (l < 9 && g.ncells != 0 && x0 <= x1) {
    const int y = cy + (l % 3) - 1, z = cz + (l / 3) - 1;
    if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
      const int row = g.nx * (y + g.ny * z);
      rb = __ldg(cell_start + row + x0);
      re = __ldg(cell_start + row + x1 + 1);
    }
  }
  int cb = 0, ce = 0;  // lane c < 27 owns cell c = (dz+1)*9 + (dy+1)*3 + (dx+1)
  float cdm = FLT_MAX; // ... and its distance from the query
  if (l < 27 && g.ncells != 0) {
    const int ox = l % 3 - 1, oy = (l / 3) % 3 - 1, oz = l / 9 - 1;
    const int X = cx + ox, Y = cy + oy, Z = cz + oz;
    if (X >= 0 && X < g.nx && Y >= 0 && Y < g.ny && Z >= 0 && Z < g.nz) {
      const int cell = X + g.nx * (Y + g.ny * Z);
      cb = __ldg(cell_start + cell);
      ce = __ldg(cell_start + cell + 1);
    }
    const float gx = ox < 0 ? fsub(qx, flx) : ox > 0 ? fsub(fadd(flx, 1.0f), qx) : 0.0f;
    const float gy = oy < 0 ? fsub(qy, fly) : oy > 0 ? fsub(fadd(fly, 1.0f), qy) : 0.0f;
    const float gz = oz < 0 ? fsub(qz, flz) : oz > 0 ? fsub(fadd(flz, 1.0f), qz) : 0.0f;
    cdm = fadd(fadd(fmul(gx, gx), fmul(gy, gy)), fmul(gz, gz));
  }
  const int total = __reduce_add_sync(0xffffffffu, re - rb);
  if (kStaged && total <= kKnnPruneAbove) {
    if (total > 0) knn5_scan_staged(rb, re, total, qx, qy, qz, cell_pts, *st, k, rej);
  } else if (total <= kKnnPruneAbove) {
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int b = __shfl_sync(0xffffffffu, rb, r), e = __shfl_sync(0xffffffffu, re, r);
      knn5_scan_range(b, e, qx, qy, qz, FLT_MAX, cell_pts, k, rej);
    }
  } else {
    knn5_scan_range(__shfl_sync(0xffffffffu, cb, 13), __shfl_sync(0xffffffffu, ce, 13), qx, qy, qz, FLT_MAX, cell_pts, k, rej);
    float bound = knn5_warp_bound(k);
#pragma unroll 1
    for (int c = 0; c < 27; ++c) {
      const int b = __shfl_sync(0xffffffffu, cb, c), e = __shfl_sync(0xffffffffu, ce, c);
      const float dm = __shfl_sync(0xffffffffu, cdm, c);
      if (c == 13 || e <= b) continue;                 // warp-uniform
      if (dm > bound) { skipped = fminf(skipped, dm); continue; }
      knn5_scan_range(b, e, qx, qy, qz, bound, cell_pts, k, rej);
      if (e - b >= 32 || bound == FLT_MAX) bound = knn5_warp_bound(k);
    }
  }
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    // distances are >= 0, so their bit patterns order like the floats; ties fall to the smaller index
    const unsigned long long mine = ((unsigned long long)__float_as_uint(k.d[0]) << 32) | (unsigned int)k.id[0];
    unsigned long long best = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    out.d[j] = __uint_as_float((unsigned int)(best >> 32));
    out.id[j] = (int)(unsigned int)(best & 0xffffffffu);
    if (mine == best) {  // the winner pops its head
      k.d[0] = k.d[1]; k.id[0] = k.id[1]; k.d[1] = k.d[2]; k.id[1] = k.id[2]; k.d[2] = k.d[3]; k.id[2] = k.id[3];
      k.d[3] = k.d[4]; k.id[3] = k.id[4]; k.d[4] = FLT_MAX; k.id[4] = 0x7fffffff;
    }
  }
  // what is left: every lane's unpopped head and whatever it turned away, plus the cells that were pruned whole
  const float mine_other = fminf(rej, k.d[0]);
  others_min = fminf(__uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(mine_other))), skipped);
}

// ------------------------------------------------------------------------------------------------------------------
// residuals, Jacobians and the 28-term reduction (21 H upper triangle, 6 g, cost)
// ------------------------------------------------------------------------------------------------------------------
struct Accum {
  double v[kLmTerms];
};

__device__ __forceinline__ void loss_correct(int loss, double& r, double* J, double& cost_term) {
  const double s = r * r;
  if (loss == FLOAM_LOSS_TRIVIAL) { cost_term = 0.5 * s; return; }
  double rho0, rho1;
  if (loss == FLOAM_LOSS_HUBER) {  // ceres::HuberLoss(0.1)
    const double a = 0.1, b = a * a;
    if (s > b) {
      const double rt = sqrt(s);
      rho0 = 2.0 * a * rt - b;
      rho1 = fmax(DBL_MIN, a / rt);
    } else {
      rho0 = s; rho1 = 1.0;
    }
  } else {  // ceres::CauchyLoss(0.2), opt-in only (Q1)
    const double a = 0.2, b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    rho0 = b * log(sum);
    rho1 = fmax(DBL_MIN, 1.0 / sum);
  }
  cost_term = 0.5 * rho0;
  // Corrector with rho'' <= 0: residual and Jacobian scaled by sqrt(rho')
  const double sr = sqrt(rho1);
  r *= sr;
#pragma unroll
  for (int j = 0; j < 6; ++j) J[j] *= sr;
}

// The residual / Jacobian evaluation is not part of the bit-exact arithmetic (only the 1e-4 pose bar applies), and it is what the
// FP64 pipe spends its time on: explicit fused multiply-adds (the library is built with -fmad=false) and one reciprocal per norm.
__device__ __forceinline__ m::V3 fcross(m::V3 a, m::V3 b) {
  return {fma(a.y, b.z, -(a.z * b.y)), fma(a.z, b.x, -(a.x * b.z)), fma(a.x, b.y, -(a.y * b.x))};
}
__device__ __forceinline__ double fdot(m::V3 a, m::V3 b) { return fma(a.z, b.z, fma(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ m::V3 frotate_translate(const double* x, m::V3 v) {  // q * v + t, Eigen's _transformVector form
  const m::V3 qv{x[0], x[1], x[2]};
  m::V3 uv = fcross(qv, v);
  uv = {uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  const m::V3 c = fcross(qv, uv);
  return {fma(x[3], uv.x, v.x) + c.x + x[4], fma(x[3], uv.y, v.y) + c.y + x[5], fma(x[3], uv.z, v.z) + c.z + x[6]};
}
// EdgeAnalyticCostFunction::Evaluate (src/lidarOptimization.cpp:12-43)
__device__ __forceinline__ void eval_edge(const double* x, m::V3 p, m::V3 a, m::V3 b, double& r, double* J) {
  const m::V3 lp = frotate_translate(x, p);
  const m::V3 nu = fcross(m::sub(lp, a), m::sub(lp, b));
  const m::V3 de = m::sub(a, b);
  const double inv_de = rsqrt(fdot(de, de)), nun = sqrt(fdot(nu, nu));
  r = nun * inv_de;
  const double k = -inv_de / nun;          // w / |de|, w = -nu / |nu|
  const m::V3 ws = fcross(m::V3{nu.x * k, nu.y * k, nu.z * k}, de);   // (w^T skew(de)) / |de|
  const m::V3 jr = fcross(lp, ws);                                    // ... * (-skew(lp))
  J[0] = jr.x; J[1] = jr.y; J[2] = jr.z; J[3] = ws.x; J[4] = ws.y; J[5] = ws.z;
}
// SurfNormAnalyticCostFunction::Evaluate (:51-74)
__device__ __forceinline__ void eval_surf(const double* x, m::V3 p, m::V3 n, double d, double& r, double* J) {
  const m::V3 pw = frotate_translate(x, p);
  r = fdot(n, pw) + d;
  const m::V3 jr = fcross(pw, n);      // n^T * (-skew(pw))
  J[0] = jr.x; J[1] = jr.y; J[2] = jr.z; J[3] = n.x; J[4] = n.y; J[5] = n.z;
}

__device__ __forceinline__ void accumulate(Accum& A, double r, const double* J, double cost_term) {
  int k = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = a; b < 6; ++b) { A.v[k] = fma(J[a], J[b], A.v[k]); ++k; }
#pragma unroll
  for (int a = 0; a < 6; ++a) A.v[21 + a] = fma(J[a], r, A.v[21 + a]);
  A.v[27] += cost_term;
  A.v[28] += 1.0;   // correspondence count (exact in double)
}

// Sum of each of the 28 accumulators over the 32 lanes of a warp with a transposing butterfly: at every level a lane hands half of
// its (padded to 32) values to its partner and keeps the other half, so 16+8+4+2+1 = 31 shuffles replace 28 x 5; lane t ends up
// holding the warp total of term t. The summation order is fixed.
__device__ __forceinline__ double warp_reduce_terms(const Accum& A) {
  const int l = lane_id();
  double v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = k < kLmTerms ? A.v[k] : 0.0;
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool upper = (l & h) != 0;   // this lane keeps terms [h, 2h) of its current 2h, its partner keeps [0, h)
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const double send = upper ? v[i] : v[i + h];
      const double keep = upper ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

// PoseSE3Parameterization gradient projection used for gradient_max_norm: |x - Plus(x, -g)|_inf
__device__ __noinline__ double gradient_max_norm(const double* x, const double* g) {
  // Only ever compared with gradient_tolerance = 1e-10. With max|g| >= 1e-3 the projected step moves x by far more than that
  // (quaternion part by |g_w|/4 at least; if g_w is below 4e-10 the translation part moves by |g_v| - |g_w x t| > 1e-3 - 4e-5),
  // so the exact value (a full se3 exponential) is only worth computing for tiny gradients.
  double gm = 0.0;
  for (int j = 0; j < 6; ++j) gm = fmax(gm, fabs(g[j]));
  if (gm >= 1e-3 && isfinite(gm)) return 1.0;
  double ng[6], proj[7];
  for (int j = 0; j < 6; ++j) ng[j] = -g[j];
  m::se3_plus(x, ng, proj);
  double mx = 0.0;
  for (int j = 0; j < 7; ++j) mx = fmax(mx, fabs(x[j] - proj[j]));
  return mx;
}
__device__ __forceinline__ double norm7(const double* a) {
  double s = 0;
  for (int i = 0; i < 7; ++i) s += a[i] * a[i];
  return sqrt(s);
}
__device__ __forceinline__ int tri(int a, int b) {  // index of (a,b), a <= b, in the packed upper triangle
  return a * 6 - a * (a - 1) / 2 + (b - a);
}

__device__ void lm_finish(PoseState& S, int termination) {
  S.lm_done = 1;
  S.termination = termination;
  S.lm_iterations_last = S.iteration;
  S.lm_accepted_last = S.accepted;
  S.lm_termination_last = termination;
  S.lm_final_cost_last = S.cost;
  S.outer_iterations++;
}

// Ceres TrustRegionMinimizer loop head up to the point where a candidate has to be evaluated.
// Everything is a function of H = J^T J, g = J^T r and the cost at the current point.
// __noinline__: called from lm_start and lm_after_candidate; one copy keeps the cluster kernel's instruction footprint down
__device__ __noinline__ void lm_next_candidate(PoseState& S) {
  for (;;) {
    if (S.iteration >= 4) { lm_finish(S, 0); return; }                                 // max_num_iterations
    if (S.last_successful && S.gmax <= 1e-10) { lm_finish(S, 3); return; }             // gradient_tolerance
    if (S.radius <= 1e-32) { lm_finish(S, 6); return; }                                // min_trust_region_radius
    S.iteration++;
    S.last_successful = 0;
    if (!S.reuse_diag) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const double d = S.scale[j] * S.scale[j] * S.H[tri(j, j)];
        S.diag[j] = fmin(fmax(d, 1e-6), 1e32);
      }
    }
    // lower triangle of H_s = S H S (Jacobi-scaled normal matrix); the damped system adds D^2 on the diagonal
    double A[36], gs[6], y[6], hs_diag[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      gs[a] = S.scale[a] * S.g[a];
#pragma unroll
      for (int b = 0; b <= a; ++b) A[a * 6 + b] = S.scale[a] * S.scale[b] * S.H[tri(b, a)];
    }
    // D^2 = diag / radius (Ceres forms sqrt(diag / radius) and the QR squares it again; one reciprocal serves the six entries)
    const double inv_radius = 1.0 / S.radius;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      hs_diag[j] = A[j * 6 + j];
      A[j * 6 + j] += S.diag[j] * inv_radius;
    }
    const bool ok = m::cholesky6_solve(A, gs, y);
    S.reuse_diag = 1;
    double step[6], mcc = 0.0;
    if (ok) {
#pragma unroll
      for (int j = 0; j < 6; ++j) step[j] = -y[j];
      double lin = 0.0, quad = 0.0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        lin += step[a] * gs[a];
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < 6; ++b) t += (a == b ? hs_diag[a] : (b < a ? A[a * 6 + b] : A[b * 6 + a])) * step[b];
        quad += step[a] * t;
      }
      mcc = -(lin + 0.5 * quad);
    }
    if (!ok || !(mcc > 0.0)) {  // HandleInvalidStep
      S.radius *= 0.5;
      continue;
    }
    S.model_cost_change = mcc;
    double delta[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) delta[j] = step[j] * S.scale[j];
    m::se3_plus(S.x, delta, S.x_cand);
    return;  // candidate pending
  }
}

// iteration 0: H, g, cost at the starting point are in place
__device__ void lm_start(PoseState& S, const double* sums, int n_corr) {
  S.lm_done = 0; S.iteration = 0; S.accepted = 0; S.termination = 0;
  S.radius = 1e4; S.decrease_factor = 2.0; S.reuse_diag = 0; S.last_successful = 1;
  for (int i = 0; i < 21; ++i) { S.H[i] = sums[i]; S.H0[i] = sums[i]; }
  for (int i = 0; i < 6; ++i) { S.g[i] = sums[21 + i]; S.g0[i] = sums[21 + i]; }
  S.cost = sums[27];
  S.initial_cost = S.cost;
  S.n_corr = n_corr;
  if (n_corr == 0) { lm_finish(S, 5); return; }           // no residual blocks: Solve returns immediately
  bool finite = isfinite(S.cost);
  for (int i = 0; i < 27; ++i) finite = finite && isfinite(sums[i]);
  if (!finite) { lm_finish(S, 4); return; }               // iteration-0 evaluation failure leaves x unchanged
  for (int j = 0; j < 6; ++j) S.scale[j] = 1.0 / (1.0 + sqrt(S.H[tri(j, j)]));  // Jacobi scaling, computed once per solve
  S.x_norm = norm7(S.x);
  S.gmax = gradient_max_norm(S.x, S.g);
  lm_next_candidate(S);
}

// the candidate's cost, H and g have been reduced
__device__ void lm_after_candidate(PoseState& S, const double* sums) {
  double cand_cost = sums[27];
  bool finite = isfinite(cand_cost);
  for (int i = 0; i < 27; ++i) finite = finite && isfinite(sums[i]);
  if (!finite) cand_cost = DBL_MAX;
  double diff[7];
  for (int j = 0; j < 7; ++j) diff[j] = S.x[j] - S.x_cand[j];
  if (norm7(diff) <= 1e-8 * (S.x_norm + 1e-8)) { lm_finish(S, 1); return; }        // parameter tolerance: candidate discarded
  const double cost_change = S.cost - cand_cost;
  if (fabs(cost_change) <= 1e-6 * S.cost) { lm_finish(S, 2); return; }              // function tolerance: candidate discarded
  const double rel = cost_change / S.model_cost_change;
  if (rel > 1e-3) {  // HandleSuccessfulStep
    for (int j = 0; j < 7; ++j) S.x[j] = S.x_cand[j];
    S.x_norm = norm7(S.x);
    for (int i = 0; i < 21; ++i) S.H[i] = sums[i];
    for (int i = 0; i < 6; ++i) S.g[i] = sums[21 + i];
    S.cost = cand_cost;
    S.gmax = gradient_max_norm(S.x, S.g);
    S.last_successful = 1;
    S.accepted++;
    const double t = 2.0 * rel - 1.0;
    S.radius = fmin(1e16, S.radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
    S.decrease_factor = 2.0;
    S.reuse_diag = 0;
  } else {           // HandleUnsuccessfulStep
    S.radius = S.radius / S.decrease_factor;
    S.decrease_factor *= 2.0;
    S.reuse_diag = 1;
  }
  lm_next_candidate(S);
}

constexpr int kEvalThreads = 128;

// The trust-region bookkeeping runs on one thread; it works on a shared-memory copy of the state (tens of dependent reads and
// writes per step: ~30-cycle shared accesses instead of L2 round trips) that the whole CTA copies in and out.
static_assert(sizeof(PoseState) % 4 == 0, "PoseState is copied as 32-bit words");
__device__ __forceinline__ void state_load(PoseState* shared_dst, const PoseState* global_src) {
  const unsigned int* s = reinterpret_cast<const unsigned int*>(global_src);
  unsigned int* d = reinterpret_cast<unsigned int*>(shared_dst);
  for (int i = threadIdx.x; i < (int)(sizeof(PoseState) / 4); i += blockDim.x) d[i] = __ldcg(s + i);  // L2: where the atomics of this kernel landed
}
__device__ __forceinline__ void state_store(PoseState* global_dst, const PoseState* shared_src) {
  const unsigned int* s = reinterpret_cast<const unsigned int*>(shared_src);
  unsigned int* d = reinterpret_cast<unsigned int*>(global_dst);
  for (int i = threadIdx.x; i < (int)(sizeof(PoseState) / 4); i += blockDim.x) d[i] = s[i];
}

// Block reduction of the accumulators into partials[blockIdx.x][*] (fixed order -> deterministic). The rows are added up by CTA 0 of
// the cluster kernel that follows.
__device__ void write_partials(const Accum& A, double* __restrict__ partials) {
  __shared__ double s_part[kEvalThreads / 32][kLmTerms];
  const int w = warp_id(), l = lane_id();
  {
    const double v = warp_reduce_terms(A);
    if (l < kLmTerms) s_part[w][l] = v;
  }
  __syncthreads();
  if (threadIdx.x < kLmTerms) {
    double v = 0.0;
#pragma unroll
    for (int ww = 0; ww < kEvalThreads / 32; ++ww) v += s_part[ww][threadIdx.x];
    partials[(size_t)blockIdx.x * kLmTerms + threadIdx.x] = v;
  }
}

// One outer iteration's association (:144-251), in two kernels.
// (1) assoc_knn_kernel: pointAssociateToMap + nearestKSearch(5), one WARP per query (lanes share the candidate scan).
//     Slots [0, nde) are edge queries against the edge map, [nde, nde+nds) surf queries against the surf map.
constexpr int kKnnThreads = 256;
// Later outer iterations of the same update (reuse != 0; the maps do not change in between) first try to KEEP the previous neighbour
// set: the previous search left, per query, the query position q_s and a lower bound B of the true distance from q_s to every map
// point outside the stored five (from the 6th computed distance; points outside the 27 searched cells are more than 1 m away). The
// pose moved the query by delta, so every other point is at least B - delta from the new position q'. If that, squared and shaved by
// the float rounding of a computed distance, still exceeds the largest computed distance D5' from q' to the stored five — and the
// five still lie in the 27 cells around q' — no other point can enter: the exhaustive search would return exactly these five, and
// they are re-sorted by (distance, index). Otherwise the full search runs. Either way the outputs are those of the full search.
template <bool kStaged>
__global__ void __launch_bounds__(kKnnThreads) assoc_knn_kernel(const PoseState* __restrict__ S, const P4* __restrict__ ds_edge, const int* __restrict__ d_nde,
                                                                 const P4* __restrict__ ds_surf, const int* __restrict__ d_nds, LocalMap emap, LocalMap smap,
                                                                 int qcap, int* __restrict__ knn_ids, float* __restrict__ knn_d2, float4* __restrict__ knn_q,
                                                                 int reuse) {
  pdl_prologue();
  // every scalar this kernel needs is requested before the first one is looked at: one L2 round trip, not three
  const int skip = S->skip_solve;
  const int nde = *d_nde, nds = *d_nds;
  double x[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) x[k] = S->x[k];
  const GridDims ge = *emap.dims, gs = *smap.dims;
  if (skip) return;
  const int l = lane_id();
  __shared__ __align__(128) float4 s_slab[kStaged ? kKnnThreads / 32 : 1][kStaged ? kKnnPruneAbove : 1];
  __shared__ __align__(8) unsigned long long s_bar[kKnnThreads / 32];
  KnnStage stage{s_slab[kStaged ? warp_id() : 0], (unsigned int)__cvta_generic_to_shared(&s_bar[warp_id()]), 0u};
  if (kStaged) {
    if (l == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(stage.bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  const int warps_total = gridDim.x * (kKnnThreads / 32);
  for (int slot = blockIdx.x * (kKnnThreads / 32) + warp_id(); slot < nde + nds; slot += warps_total) {
    const bool is_edge = slot < nde;
    const int qi = is_edge ? slot : slot - nde;
    const int out = is_edge ? qi : qcap + qi;
    const float4 p = __ldg((is_edge ? ds_edge : ds_surf) + qi);
    // pointAssociateToMap :126-135: double transform, float store
    const m::V3 pw = m::add(m::quat_rotate(x, m::V3{(double)p.x, (double)p.y, (double)p.z}), m::V3{x[4], x[5], x[6]});
    const float qx = (float)pw.x, qy = (float)pw.y, qz = (float)pw.z;
    const LocalMap& map = is_edge ? emap : smap;
    if (reuse) {
      const float4 qs = knn_q[out];
      const int pid = l < 5 ? knn_ids[(size_t)out * 5 + l] : 0;
      if (qs.w > 0.f && __shfl_sync(0xffffffffu, pid, 0) >= 0) {   // warp-uniform
        float dist = 0.f;
        bool inside = true;
        if (l < 5) {
          const float4 mp = __ldg(map.pts + pid);
          const float dx = fsub(qx, mp.x), dy = fsub(qy, mp.y), dz = fsub(qz, mp.z);
          dist = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
          inside = fabsf(floorf(mp.x) - floorf(qx)) <= 1.f && fabsf(floorf(mp.y) - floorf(qy)) <= 1.f && fabsf(floorf(mp.z) - floorf(qz)) <= 1.f;
        }
        const float d5 = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(dist)));   // distances >= 0: bit patterns order like floats
        const bool all_inside = __all_sync(0xffffffffu, inside);
        const double ddx = (double)qx - (double)qs.x, ddy = (double)qy - (double)qs.y, ddz = (double)qz - (double)qs.z;
        const double delta = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        const double slack = (double)qs.w - delta * (1.0 + 1e-6) - 1e-9;
        if (all_inside && slack > 0.0 && slack * slack * (1.0 - 1e-6) > (double)d5) {
          int rank = 0;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const float dk = __shfl_sync(0xffffffffu, dist, k);
            const int ik = __shfl_sync(0xffffffffu, pid, k);
            rank += (dk < dist || (dk == dist && ik < pid)) ? 1 : 0;
          }
          const bool near = d5 < 1.0f;
          if (l < 5) {
            knn_ids[(size_t)out * 5 + rank] = near ? pid : -1;
            knn_d2[(size_t)out * 5 + rank] = near ? dist : 0.f;
          }
          if (l == 0) knn_q[out] = make_float4(qx, qy, qz, near ? __double2float_rd(slack) : -1.f);
          continue;
        }
      }
    }
    Knn5 nn;
    float others;
    knn5_search_warp<kStaged>(is_edge ? ge : gs, map.cell_start, map.cell_pts, qx, qy, qz, nn, others, &stage);
    const bool near = nn.d[4] < 1.0f;  // pointSearchSqDis[4] < 1.0 : the only queries the reference uses
    if (l < 5) {
      const int j = l;
      const int id = j == 0 ? nn.id[0] : j == 1 ? nn.id[1] : j == 2 ? nn.id[2] : j == 3 ? nn.id[3] : nn.id[4];
      const float d = j == 0 ? nn.d[0] : j == 1 ? nn.d[1] : j == 2 ? nn.d[2] : j == 3 ? nn.d[3] : nn.d[4];
      knn_ids[(size_t)out * 5 + j] = near ? id : -1;
      knn_d2[(size_t)out * 5 + j] = near ? d : 0.f;
    }
    if (l == 0 && knn_q) {
      // lower bound of the TRUE distance to every point outside the result: the 6th computed distance (less its rounding), and one
      // cell width for everything outside the 27 cells
      double lb = 1.0 - 1e-6;
      if (others < FLT_MAX) lb = fmin(lb, sqrt((double)others) * (1.0 - 1e-6));
      knn_q[out] = make_float4(qx, qy, qz, near ? __double2float_rd(lb) : -1.f);
    }
  }
}

// (2) assoc_eval_kernel: one THREAD per query: PCA line fit / 5x3 QR plane fit of the five neighbours, acceptance gates, and the
//     iteration-0 residual + Jacobian of ceres::Solve reduced into the normal equations; the last CTA starts the LM.
__global__ void __launch_bounds__(kEvalThreads) assoc_eval_kernel(PoseState* __restrict__ S, const P4* __restrict__ ds_edge, const int* __restrict__ d_nde,
                                                                   const P4* __restrict__ ds_surf, const int* __restrict__ d_nds, LocalMap emap, LocalMap smap,
                                                                   int qcap, double* __restrict__ corr, unsigned char* __restrict__ corr_ok,
                                                                   const int* __restrict__ knn_ids, int loss, double* __restrict__ partials) {
  pdl_prologue();
  const int skip = S->skip_solve;   // scalars requested together (see assoc_knn_kernel)
  const int nde = *d_nde, nds = *d_nds;
  double x[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) x[k] = S->x[k];
  if (skip) return;
  Accum A;
#pragma unroll
  for (int k = 0; k < kLmTerms; ++k) A.v[k] = 0.0;
  const size_t cs = (size_t)2 * qcap;  // stride between the planes of corr
  for (int slot = blockIdx.x * kEvalThreads + threadIdx.x; slot < nde + nds; slot += gridDim.x * kEvalThreads) {
    const bool is_edge = slot < nde;
    const int qi = is_edge ? slot : slot - nde;
    const int out = is_edge ? qi : qcap + qi;
    int id[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) id[j] = __ldg(knn_ids + (size_t)out * 5 + j);
    const bool near = id[0] >= 0;
    bool ok = false;
    double r = 0.0, J[6], cost_term = 0.0;
    if (near) {
      const float4 p = __ldg((is_edge ? ds_edge : ds_surf) + qi);
      const m::V3 pc{(double)p.x, (double)p.y, (double)p.z};
      const P4* mpts = is_edge ? emap.pts : smap.pts;
      m::V3 q[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const float4 mp = __ldg(mpts + id[j]);
        q[j] = m::V3{(double)mp.x, (double)mp.y, (double)mp.z};
      }
      if (is_edge) {
        m::V3 c{0, 0, 0};
#pragma unroll
        for (int j = 0; j < 5; ++j) c = m::add(c, q[j]);
        c = m::V3{c.x / 5.0, c.y / 5.0, c.z / 5.0};
        double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const m::V3 d = m::sub(q[j], c);
          c00 += d.x * d.x; c01 += d.x * d.y; c02 += d.x * d.z; c11 += d.y * d.y; c12 += d.y * d.z; c22 += d.z * d.z;
        }
        double vals[3], u[3];
        m::eigen3_sym(c00, c01, c11, c02, c12, c22, vals, u);
        if (vals[2] > 3 * vals[1]) {
          const m::V3 a{0.1 * u[0] + c.x, 0.1 * u[1] + c.y, 0.1 * u[2] + c.z};
          const m::V3 b{-0.1 * u[0] + c.x, -0.1 * u[1] + c.y, -0.1 * u[2] + c.z};
          corr[0 * cs + out] = a.x; corr[1 * cs + out] = a.y; corr[2 * cs + out] = a.z;
          corr[3 * cs + out] = b.x; corr[4 * cs + out] = b.y; corr[5 * cs + out] = b.z;
          eval_edge(x, pc, a, b, r, J);
          ok = true;
        }
      } else {
        double matA[15];  // column-major 5x3
#pragma unroll
        for (int j = 0; j < 5; ++j) { matA[0 * 5 + j] = q[j].x; matA[1 * 5 + j] = q[j].y; matA[2 * 5 + j] = q[j].z; }
        const double matB[5] = {-1, -1, -1, -1, -1};
        double nv[3];
        m::colpiv_qr_solve_5x3(matA, matB, nv);
        const double nn_ = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
        const double d = 1 / nn_;
        const m::V3 n{nv[0] / nn_, nv[1] / nn_, nv[2] / nn_};
        bool valid = true;
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (fabs(n.x * q[j].x + n.y * q[j].y + n.z * q[j].z + d) > 0.2) valid = false;
        if (valid) {
          corr[0 * cs + out] = n.x; corr[1 * cs + out] = n.y; corr[2 * cs + out] = n.z; corr[3 * cs + out] = d;
          eval_surf(x, pc, n, d, r, J);
          ok = true;
        }
      }
    }
    corr_ok[out] = ok ? 1 : 0;
    if (ok) {
      loss_correct(loss, r, J, cost_term);
      accumulate(A, r, J, cost_term);
    }
  }
  write_partials(A, partials);
}

// The whole ceres::Solve step loop (<= 4 attempts) of one outer iteration in ONE thread-block cluster (8 CTAs on 8 SMs), for the
// usual problem sizes: every CTA evaluates its share of the correspondences at the pending candidate and reduces the 28 terms in
// its own shared memory; after a cluster barrier CTA 0 adds the eight partial rows through distributed shared memory (fixed order)
// and its thread 0 advances the trust-region state; a second barrier publishes the next candidate to the other CTAs, again through
// DSMEM. No partials in global memory, no ticket, no kernel boundary between attempts. (2048 threads stride over the slots, so a
// 100k-query problem costs ~50 evaluations per thread and attempt: still tens of microseconds.)
constexpr int kClusterCtas = 8;    // portable cluster size; 16 CTAs (non-portable) was measured: evaluation 3.8k instead of 5.7k cycles per
                                   // attempt, but the cluster barrier doubled (1.8k) and the frame got slower
constexpr int kClusterThreads = 256;
// Accepted correspondences of a CTA's share are staged in shared memory once per solve (stable compaction of its contiguous slot
// range: edge slots first, then surf): 6 double planes (a, b | n, d) + 3 float planes (the query point). Every step attempt then
// evaluates dense items out of shared memory instead of striding over the ~55 % rejected slots with dependent L2 loads.
constexpr int kStageCap = 3072;
constexpr int kCoordinatorOnlyBelow = 16384;   // query slots
constexpr size_t kLmStageBytes = (size_t)kStageCap * (6 * sizeof(double) + 3 * sizeof(float));   // 180 KB
struct ClusterShared {
  double part[kClusterThreads / 32][kLmTerms];
  double row[kLmTerms];   // this CTA's 28 totals
  double x[7];            // candidate pose (valid in CTA 0, read remotely)
  int done;
};
__global__ void __launch_bounds__(kClusterThreads)   // launched as ONE cluster of 8 (portable) or 16 CTAs (od.lm_cluster_ctas)
    lm_cluster_kernel(PoseState* S, const P4* __restrict__ ds_edge, const int* __restrict__ d_nde, const P4* __restrict__ ds_surf,
                      const int* __restrict__ d_nds, int qcap, const double* __restrict__ corr, const unsigned char* __restrict__ corr_ok, int loss,
                      const double* __restrict__ partials, int n_rows, FinishArgs fin) {
  pdl_prologue();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  // uniform across the cluster: decided from a value the previous kernels wrote
  if (S->skip_solve) {
    if (fin.enabled && cluster.block_rank() == 0 && threadIdx.x == 0) finish_body(S, fin.update_type, fin.scan_period, fin.traj, fin.traj_cap);
    return;
  }
  const int nde = *d_nde, nds = *d_nds;
  __shared__ ClusterShared sh;
  __shared__ PoseState st;   // CTA 0's working copy of the state
  __shared__ double s_sums[kLmTerms];
  const unsigned int rank = cluster.block_rank();
  ClusterShared* sh0 = cluster.map_shared_rank(&sh, 0);
  const size_t cs = (size_t)2 * qcap;
  const int w = warp_id(), l = lane_id();
  // iteration 0 of ceres::Solve: CTA 0 adds up the per-CTA rows the association kernel left (thread (g, k): term k of rows g, g+8, ...;
  // fixed order), starts the trust-region state and publishes the first candidate
  if (rank == 0) {
    state_load(&st, S);
    double v = 0.0;
    if (l < kLmTerms) {
#pragma unroll 4
      for (int r = w; r < n_rows; r += kClusterThreads / 32) v += __ldcg(partials + (size_t)r * kLmTerms + l);
      sh.part[w][l] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLmTerms) {
      double t = 0.0;
#pragma unroll
      for (int ww = 0; ww < kClusterThreads / 32; ++ww) t += sh.part[ww][threadIdx.x];
      s_sums[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      lm_start(st, s_sums, (int)s_sums[28]);
      sh.done = st.lm_done;
#pragma unroll
      for (int k = 0; k < 7; ++k) sh.x[k] = st.x_cand[k];
    }
  }
  // ---- stage this CTA's share: slots [lo, hi) of the cluster-wide slot range, warp ww owning the ww-th eighth of it ----
  // Small problems (the usual case) leave CTA 0 without a share: its iteration-0 bookkeeping above is the critical path of the
  // prologue, and the other seven CTAs stage their correspondences meanwhile.
  extern __shared__ __align__(16) unsigned char lm_dyn[];
  double* sc = reinterpret_cast<double*>(lm_dyn);                                        // [6][kStageCap]
  float* sp = reinterpret_cast<float*>(lm_dyn + (size_t)6 * kStageCap * sizeof(double));  // [3][kStageCap]
  __shared__ int s_wtot[kClusterThreads / 32], s_wedge[kClusterThreads / 32], s_over_lo;
  const int total = nde + nds;
  const int n_ctas = (int)cluster.num_blocks();
  const int n_eval = total <= kCoordinatorOnlyBelow ? n_ctas - 1 : n_ctas;
  const int erank = (int)rank - (n_ctas - n_eval);   // -1: no share
  const int chunk = (total + n_eval - 1) / n_eval;
  const int lo = erank < 0 ? total : min(erank * chunk, total), hi = min(lo + chunk, total);
  const int sub = (hi - lo + kClusterThreads / 32 - 1) / (kClusterThreads / 32);
  const int wlo = min(lo + w * sub, hi), whi = min(wlo + sub, hi);
  {
    int cnt = 0, cnt_edge = 0;
    for (int s0 = wlo; s0 < whi; s0 += 32) {
      const int slot = s0 + l;
      const bool ok = slot < whi && corr_ok[slot < nde ? slot : qcap + slot - nde] != 0;
      cnt += __popc(__ballot_sync(0xffffffffu, ok));
      cnt_edge += __popc(__ballot_sync(0xffffffffu, ok && slot < nde));
    }
    if (l == 0) { s_wtot[w] = cnt; s_wedge[w] = cnt_edge; }
    if (threadIdx.x == 0) s_over_lo = hi;
  }
  __syncthreads();
  int n_staged = 0, n_staged_edge = 0;
  {
    int base = 0;
#pragma unroll
    for (int ww = 0; ww < kClusterThreads / 32; ++ww) {
      if (ww < w) base += s_wtot[ww];
      n_staged += s_wtot[ww];
      n_staged_edge += s_wedge[ww];
    }
    n_staged = min(n_staged, kStageCap);
    n_staged_edge = min(n_staged_edge, kStageCap);
    for (int s0 = wlo; s0 < whi; s0 += 32) {
      const int slot = s0 + l;
      const bool is_edge = slot < nde;
      const int qi = is_edge ? slot : slot - nde;
      const int out = is_edge ? qi : qcap + qi;
      const bool ok = slot < whi && corr_ok[out] != 0;
      const unsigned int b = __ballot_sync(0xffffffffu, ok);
      const int pos = base + __popc(b & ((1u << l) - 1u));
      if (ok) {
        if (pos < kStageCap) {
          const float4 p = __ldg((is_edge ? ds_edge : ds_surf) + qi);
          sp[pos] = p.x; sp[kStageCap + pos] = p.y; sp[2 * kStageCap + pos] = p.z;
          const int planes = is_edge ? 6 : 4;
#pragma unroll
          for (int k = 0; k < 6; ++k)
            if (k < planes) sc[(size_t)k * kStageCap + pos] = corr[k * cs + out];
        } else if (pos == kStageCap) {
          s_over_lo = slot;   // first accepted slot that did not fit: [over_lo, hi) is evaluated from global memory
        }
      }
      base += __popc(b);
    }
  }
  __syncthreads();
  const int over_lo = s_over_lo;
  cluster.sync();
  double x[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) x[k] = sh0->x[k];
  const bool finished_at_start = sh0->done != 0;
  long long clk[6] = {0, 0, 0, 0, 0, 0};
  for (int attempt = 0; attempt < 4 && !finished_at_start; ++attempt) {
    clk[0] = clock64();
    Accum A;
#pragma unroll
    for (int k = 0; k < kLmTerms; ++k) A.v[k] = 0.0;
    for (int k = threadIdx.x; k < n_staged; k += kClusterThreads) {   // staged items: [0, n_staged_edge) edge, then surf
      const m::V3 pc{(double)sp[k], (double)sp[kStageCap + k], (double)sp[2 * kStageCap + k]};
      double r, J[6], cost_term;
      if (k < n_staged_edge) {
        const m::V3 a{sc[k], sc[kStageCap + k], sc[2 * kStageCap + k]};
        const m::V3 b{sc[3 * kStageCap + k], sc[4 * kStageCap + k], sc[5 * kStageCap + k]};
        eval_edge(x, pc, a, b, r, J);
      } else {
        const m::V3 n{sc[k], sc[kStageCap + k], sc[2 * kStageCap + k]};
        eval_surf(x, pc, n, sc[3 * kStageCap + k], r, J);
      }
      loss_correct(loss, r, J, cost_term);
      accumulate(A, r, J, cost_term);
    }
    for (int slot = over_lo + threadIdx.x; slot < hi; slot += kClusterThreads) {   // beyond the staging capacity (dense maps): from global
      const bool is_edge = slot < nde;
      const int qi = is_edge ? slot : slot - nde;
      const int out = is_edge ? qi : qcap + qi;
      if (!corr_ok[out]) continue;
      const float4 p = __ldg((is_edge ? ds_edge : ds_surf) + qi);
      const m::V3 pc{(double)p.x, (double)p.y, (double)p.z};
      double r, J[6], cost_term;
      if (is_edge) {
        const m::V3 a{corr[0 * cs + out], corr[1 * cs + out], corr[2 * cs + out]};
        const m::V3 b{corr[3 * cs + out], corr[4 * cs + out], corr[5 * cs + out]};
        eval_edge(x, pc, a, b, r, J);
      } else {
        const m::V3 n{corr[0 * cs + out], corr[1 * cs + out], corr[2 * cs + out]};
        eval_surf(x, pc, n, corr[3 * cs + out], r, J);
      }
      loss_correct(loss, r, J, cost_term);
      accumulate(A, r, J, cost_term);
    }
    clk[1] = clock64();
    {
      const double v = warp_reduce_terms(A);
      if (l < kLmTerms) sh.part[w][l] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLmTerms) {
      double v = 0.0;
#pragma unroll
      for (int ww = 0; ww < kClusterThreads / 32; ++ww) v += sh.part[ww][threadIdx.x];
      sh.row[threadIdx.x] = v;
    }
    clk[2] = clock64();
    cluster.sync();   // every CTA's row is in place
    clk[3] = clock64();
    if (rank == 0) {
      if (threadIdx.x < kLmTerms) {
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < n_ctas; ++c) v += cluster.map_shared_rank(&sh, c)->row[threadIdx.x];
        s_sums[threadIdx.x] = v;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        lm_after_candidate(st, s_sums);
        sh.done = st.lm_done;
#pragma unroll
        for (int k = 0; k < 7; ++k) sh.x[k] = st.x_cand[k];
        clk[4] = clock64();
      }
    }
    cluster.sync();   // CTA 0 has published the verdict and the next candidate
    clk[5] = clock64();
    if (rank == 0 && threadIdx.x == 0)
      for (int k = 0; k < 6; ++k) st.dbg_clk[k] = clk[k];
    const int done = sh0->done;
#pragma unroll
    for (int k = 0; k < 7; ++k) x[k] = sh0->x[k];
    if (done) break;
    // no third barrier: row[] is rewritten only after this barrier pair, and CTA 0 rewrites x/done only after the next round's first
    // barrier, which every CTA reaches after it has read this round's values
  }
  if (rank == 0) {
    // odom write-back, KeyFrameUpdate and the crop bounds (:114-121, :320-343) on the shared copy, after the last outer iteration
    if (fin.enabled && threadIdx.x == 0) finish_body(&st, fin.update_type, fin.scan_period, fin.traj, fin.traj_cap);
    __syncthreads();
    state_store(S, &st);   // ordered after thread 0's last update by the barrier pair above (and this barrier)
  }
  cluster.sync();     // CTA 0's shared memory must outlive the last remote read
}

// stand-alone 5-NN (floam_knn5): queries are used as given (no pose transform)
__global__ void __launch_bounds__(kKnnThreads) knn5_kernel(const P4* __restrict__ queries, const int* __restrict__ d_nq, LocalMap map, int* __restrict__ ids,
                                                            float* __restrict__ d2) {
  pdl_prologue();
  const int nq = *d_nq;
  const GridDims g = *map.dims;
  const int warps_total = gridDim.x * (kKnnThreads / 32);
  for (int i = blockIdx.x * (kKnnThreads / 32) + warp_id(); i < nq; i += warps_total) {   // one warp per query, like the association
    const float4 q = __ldg(queries + i);
    Knn5 nn;
    float others;
    knn5_search_warp<false>(g, map.cell_start, map.cell_pts, q.x, q.y, q.z, nn, others);
    const bool near = nn.d[4] < 1.0f;
    if (lane_id() < 5) {
      const int j = lane_id();
      const int id = j == 0 ? nn.id[0] : j == 1 ? nn.id[1] : j == 2 ? nn.id[2] : j == 3 ? nn.id[3] : nn.id[4];
      const float d = j == 0 ? nn.d[0] : j == 1 ? nn.d[1] : j == 2 ? nn.d[2] : j == 3 ? nn.d[3] : nn.d[4];
      ids[(size_t)i * 5 + j] = near ? id : -1;
      d2[(size_t)i * 5 + j] = near ? d : 0.f;
    }
  }
}

// dmapping::CompensateVelocity (src/dataHandler.cpp:82-91) with GetVelocity() taken from the device state (Q14: no rotation)
__global__ void __launch_bounds__(kThreads) compensate_velocity_kernel(PointIRT* __restrict__ pts, const int* __restrict__ d_n, const PoseState* __restrict__ S,
                                                                        int rotate) {
  pdl_prologue();
  const int n = *d_n;
  double vx = S->velocity[0], vy = S->velocity[1], vz = S->velocity[2];
  if (rotate) {   // FLOAM_FIX_ROTATED_VELOCITY: the velocity is a world-frame vector, the points are in the sensor frame: v_sensor = R(odom)^T v
    const double* R = S->odom;
    const double wx = vx, wy = vy, wz = vz;
    vx = R[0] * wx + R[3] * wy + R[6] * wz;
    vy = R[1] * wx + R[4] * wy + R[7] * wz;
    vz = R[2] * wx + R[5] * wy + R[8] * wz;
  }
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    PointIRT* p = pts + i;
    const double t = (double)p->time;
    p->x = (float)((double)p->x + vx * t);
    p->y = (float)((double)p->y + vy * t);
    p->z = (float)((double)p->z + vz * t);
  }
}

__global__ void __launch_bounds__(kThreads) compensate_velocity_explicit_kernel(PointIRT* __restrict__ pts, const int* __restrict__ d_n, double vx, double vy,
                                                                                 double vz) {
  pdl_prologue();
  const int n = *d_n;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    PointIRT* p = pts + i;
    const double t = (double)p->time;
    p->x = (float)((double)p->x + vx * t);
    p->y = (float)((double)p->y + vy * t);
    p->z = (float)((double)p->z + vz * t);
  }
}

int* dims_ncells_ptr(GridDims* dims) { return reinterpret_cast<int*>(reinterpret_cast<char*>(dims) + offsetof(GridDims, ncells)); }

// from_tmp: the cloud sits in map.tmp (output of the keyframe filter, which has also accumulated its bounding box); the scatter kernel
// copies it home into map.pts on the way
void rebuild_grid(OdomDevice& od, LocalMap& map, const int* d_skip, cudaStream_t s, VoxelWorkspace* ws = nullptr, bool from_tmp = false,
                  long long* stamp = nullptr, bool large = false) {
  if (!ws) ws = od.vws;
  const int g = grid_for(map.cap, large ? 4 : 2);
  const P4* src = from_tmp ? map.tmp : map.pts;
  if (!from_tmp) FLOAM_LAUNCH(K_GRID_BBOX, grid_bbox_kernel, g, kThreads, s, map.pts, map.d_n, map.bbox, d_skip);
  FLOAM_LAUNCH(K_GRID_COUNT, grid_count_kernel, g, kThreads, s, src, map.d_n, map.bbox, map.dims, map.ncells_cap, od.state, map.cell_count, map.tile_sums,
               d_skip);
  exclusive_scan_with_tile_sums(map.cell_count, map.cell_start, dims_ncells_ptr(map.dims), map.ncells_cap, map.tile_sums, d_skip, s);
  FLOAM_LAUNCH(K_GRID_SCATTER, grid_scatter_kernel, g, kThreads, s, src, map.d_n, map.dims, map.cell_start, map.cell_count, map.cell_pts, map.bbox,
               map.tile_sums, from_tmp ? map.pts : (P4*)nullptr, stamp, d_skip);
}

}  // namespace

int local_map_alloc(LocalMap& map, int cap, int ncells_cap, void* (*alloc)(void*, size_t), void* actx, cudaStream_t s) {
  map.cap = cap;
  map.ncells_cap = ncells_cap;
  map.pts = (P4*)alloc(actx, (size_t)cap * sizeof(P4));
  map.tmp = (P4*)alloc(actx, (size_t)cap * sizeof(P4));
  map.cell_pts = (float4*)alloc(actx, (size_t)cap * sizeof(float4));
  map.cell_start = (int*)alloc(actx, ((size_t)ncells_cap + 1) * 4);
  map.cell_count = (int*)alloc(actx, (size_t)ncells_cap * 4);
  const size_t tile_bytes = ((size_t)ncells_cap / kScanTile + 2) * 4;
  map.tile_sums = (int*)alloc(actx, tile_bytes);
  map.dims = (GridDims*)alloc(actx, sizeof(GridDims));
  map.bbox = (unsigned int*)alloc(actx, 32);
  int* ints = (int*)alloc(actx, 4 * 4);
  if (!map.pts || !map.tmp || !map.cell_pts || !map.cell_start || !map.cell_count || !map.dims || !map.bbox || !ints) return FLOAM_ERR_CUDA;
  map.d_n = ints; map.d_ntmp = ints + 1; map.d_ncrop = ints + 2; map.d_ncells = ints + 3;
  if (!map.tile_sums) return FLOAM_ERR_CUDA;
  FLOAM_CUDA_OK(cudaMemsetAsync(map.cell_count, 0, (size_t)ncells_cap * 4, s));
  FLOAM_CUDA_OK(cudaMemsetAsync(map.tile_sums, 0, tile_bytes, s));
  FLOAM_CUDA_OK(cudaMemsetAsync(map.dims, 0, sizeof(GridDims), s));
  FLOAM_CUDA_OK(cudaMemsetAsync(ints, 0, 16, s));
  const unsigned int bb[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  FLOAM_CUDA_OK(cudaMemcpyAsync(map.bbox, bb, sizeof(bb), cudaMemcpyHostToDevice, s));
  FLOAM_CUDA_OK(cudaStreamSynchronize(s));
  return FLOAM_OK;
}

int odom_device_init(OdomDevice& od, const floam_params& prm, VoxelWorkspace* vws, VoxelWorkspace* vws_aux, cudaStream_t aux, void* (*alloc)(void*, size_t),
                     void* actx, cudaStream_t s) {
  od.vws = vws;
  od.vws_aux = vws_aux;
  od.aux_stream = aux;
  FLOAM_CUDA_OK(cudaFuncSetAttribute(lm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLmStageBytes));
  FLOAM_CUDA_OK(cudaFuncSetAttribute(lm_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  // Cluster size of the solve: 8 CTAs (portable) for the usual few thousand correspondences; dense configurations (128-line sensors,
  // map resolution <= 0.2 m: tens of thousands of correspondences, FP64-issue bound on 8 SMs) get 16. A property of the context,
  // so that every entry path adds the normal equations up in the same order. FLOAM_LM_CLUSTER overrides (8 or 16).
  od.lm_cluster_ctas = (prm.num_lines >= 128 || prm.map_resolution <= 0.2) ? 16 : kClusterCtas;
  if (const char* e = std::getenv("FLOAM_LM_CLUSTER")) od.lm_cluster_ctas = std::atoi(e) == 16 ? 16 : kClusterCtas;
  if (od.lm_cluster_ctas == 16) {   // a part / partition whose GPCs cannot co-schedule 16 such CTAs must not fail every solve: fall back to 8
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16); cfg.blockDim = dim3(kClusterThreads); cfg.dynamicSmemBytes = kLmStageBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, lm_cluster_kernel, &cfg) != cudaSuccess || n_clusters < 1) {
      cudaGetLastError();
      od.lm_cluster_ctas = kClusterCtas;
    }
  }
  od.knn_staged = true;   // TMA-staged cell tiles for sparse neighbourhoods (A/B: profiles/r2_knn_tma_ab.md); FLOAM_KNN_TMA=0 -> direct loads
  if (const char* e = std::getenv("FLOAM_KNN_TMA")) od.knn_staged = std::atoi(e) != 0;
  od.map_merge_mode = 1;
  if (const char* e = std::getenv("FLOAM_MAP_MERGE")) od.map_merge_mode = std::max(0, std::min(2, std::atoi(e)));
  if (const char* e = std::getenv("FLOAM_MAP_MERGE_MIN")) od.map_merge_min_points = std::atoi(e);
  FLOAM_CUDA_OK(cudaEventCreateWithFlags(&od.ev_fork, cudaEventDisableTiming));
  FLOAM_CUDA_OK(cudaEventCreateWithFlags(&od.ev_join, cudaEventDisableTiming));
  od.leaf_edge = (float)prm.map_resolution;        // setLeafSize(float...) :13-14
  od.leaf_surf = (float)(prm.map_resolution * 2);
  od.scan_period = prm.scan_period;
  od.loss = prm.loss;
  od.fixes = prm.fixes;
  od.optimization_count = 2;                       // :22
  od.qcap = prm.max_scan_points;
  const int ncells_cap = prm.max_grid_cells;
  od.state = (PoseState*)alloc(actx, sizeof(PoseState));
  if (!od.state) return FLOAM_ERR_CUDA;
  int rc = local_map_alloc(od.edge_map, prm.max_map_points, ncells_cap, alloc, actx, s);
  if (rc) return rc;
  rc = local_map_alloc(od.surf_map, prm.max_map_points, ncells_cap, alloc, actx, s);
  if (rc) return rc;
  for (int k = 0; k < 2; ++k) {
    od.ds_edge_b[k] = (P4*)alloc(actx, (size_t)od.qcap * sizeof(P4));
    od.ds_surf_b[k] = (P4*)alloc(actx, (size_t)od.qcap * sizeof(P4));
    if (!od.ds_edge_b[k] || !od.ds_surf_b[k]) return FLOAM_ERR_CUDA;
  }
  int* ints = (int*)alloc(actx, 16);
  od.corr = (double*)alloc(actx, (size_t)6 * 2 * od.qcap * sizeof(double));
  od.corr_ok = (unsigned char*)alloc(actx, (size_t)2 * od.qcap);
  od.knn_ids = (int*)alloc(actx, (size_t)2 * od.qcap * 5 * 4);
  od.knn_q = (float4*)alloc(actx, (size_t)2 * od.qcap * sizeof(float4));
  od.knn_d2 = (float*)alloc(actx, (size_t)2 * od.qcap * 5 * 4);
  od.partials = (double*)alloc(actx, (size_t)kAssocBlocks * kLmTerms * sizeof(double));
  od.traj_cap = 1 << 16;
  od.traj = (double*)alloc(actx, (size_t)od.traj_cap * 7 * sizeof(double));
  if (!ints || !od.corr || !od.corr_ok || !od.knn_ids || !od.knn_d2 || !od.knn_q || !od.partials || !od.traj) return FLOAM_ERR_CUDA;
  od.d_nds_edge_b[0] = ints; od.d_nds_surf_b[0] = ints + 1; od.d_nds_edge_b[1] = ints + 2; od.d_nds_surf_b[1] = ints + 3;
  odom_select_buffers(od, 0);
  FLOAM_CUDA_OK(cudaMemsetAsync(ints, 0, 16, s));
  FLOAM_CUDA_OK(cudaMemsetAsync(od.corr_ok, 0, (size_t)2 * od.qcap, s));
  FLOAM_LAUNCH(K_STATE_INIT, state_init_kernel, 1, 32, s, od.state);
  FLOAM_CUDA_OK(cudaStreamSynchronize(s));
  return FLOAM_OK;
}

void odom_reset_state(OdomDevice& od, cudaStream_t s) {
  FLOAM_LAUNCH(K_STATE_INIT, state_init_kernel, 1, 32, s, od.state);
  od.optimization_count = 2;
}

__global__ void mail_state_kernel(PoseState* __restrict__ S, const int* __restrict__ d_flags, PoseState* h_state, int* h_flags) {
  pdl_prologue();
  const unsigned int* src = reinterpret_cast<const unsigned int*>(S);
  unsigned int* dst = reinterpret_cast<unsigned int*>(h_state);
  for (int i = threadIdx.x; i < (int)(sizeof(PoseState) / 4); i += blockDim.x) dst[i] = __ldcg(src + i);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ff = *d_flags;
    *h_flags = ff;
    // the flags belong to the frame just mailed: the next frame starts clean (the OR over frames stays in error_sticky)
    S->error_sticky |= S->error_flags | (ff << 4);
    S->error_flags = 0;
  }
  // no system fence: the host reads the mailbox only after an event recorded behind this kernel, and kernel completion flushes the stores
}
void odom_mail_state(OdomDevice& od, const int* d_flags, PoseState* h_state, int* h_flags, cudaStream_t s) {
  FLOAM_LAUNCH(K_MAIL_STATE, mail_state_kernel, 1, 128, s, od.state, d_flags, h_state, h_flags);
}

void odom_record_pose(OdomDevice& od, cudaStream_t s) { FLOAM_LAUNCH(K_RECORD_POSE, record_pose_kernel, 1, 32, s, od.state, od.traj, od.traj_cap); }

void odom_rebuild_grids(OdomDevice& od, cudaStream_t s) {
  rebuild_grid(od, od.edge_map, nullptr, s);
  rebuild_grid(od, od.surf_map, nullptr, s);
}

void local_map_load(OdomDevice& od, LocalMap& map, const void* d_pts, const int* d_n, int stride, int n_max, int replace, cudaStream_t s) {
  FLOAM_LAUNCH(K_MAP_APPEND_RAW, map_append_raw_kernel, grid_for(n_max), kThreads, s, (const char*)d_pts, stride, d_n, map.pts, map.d_n, map.cap, replace, &od.state->error_flags);
  FLOAM_LAUNCH(K_MAP_BUMP, map_bump_kernel, 1, 32, s, map.d_n, d_n, map.cap, replace, nullptr);
  rebuild_grid(od, map, nullptr, s);
}

void odom_init_map_device(OdomDevice& od, const void* d_edge, const int* d_ne, const void* d_surf, const int* d_ns, int stride, int n_max, int replace,
                          cudaStream_t s) {
  local_map_load(od, od.edge_map, d_edge, d_ne, stride, n_max, replace, s);
  local_map_load(od, od.surf_map, d_surf, d_ns, stride, n_max, replace, s);
}

bool odom_map_update_merges(const OdomDevice& od, int k) {
  const int hint = k == 0 ? od.surf_map_hint : od.edge_map_hint;
  return od.map_merge_mode == 2 || (od.map_merge_mode == 1 && hint >= od.map_merge_min_points);
}

void odom_update_device(OdomDevice& od, const void* d_edge, const int* d_ne, const void* d_surf, const int* d_ns, int stride, int n_max, int update_type,
                        int ds_ready, cudaStream_t s) {
  // the caller has already applied `if (optimization_count > 2) optimization_count--` (:59-60, Q4)
  PoseState* S = od.state;
  const int keep_pose = (update_type == FLOAM_REFINEMENT_AND_UPDATE && (od.fixes & FLOAM_FIX_SINGLE_PREDICTION)) ? 1 : 0;
  FLOAM_LAUNCH(K_PREDICT, predict_kernel, 1, 32, s, S, od.edge_map.d_n, od.surf_map.d_n, keep_pose);
  cudaStream_t a = od.aux_stream;
  // downSamplingToMap :137-142, unless the frame pipeline already did it ahead of time
  if (!ds_ready) odom_downsample_device(od, d_edge, d_ne, d_surf, d_ns, stride, n_max, *od.vws, *od.vws_aux, s, a, od.ev_fork, od.ev_join);
  for (int it = 0; it < od.optimization_count; ++it) {
    PdlSolveScope pdl;   // kNN, fit and LM may be scheduled while their predecessor drains (griddepcontrol.wait orders the data)
    if (od.knn_staged)
      FLOAM_LAUNCH(K_ASSOC_KNN, assoc_knn_kernel<true>, kKnnBlocks, kKnnThreads, s, S, od.ds_edge, od.d_nds_edge, od.ds_surf, od.d_nds_surf, od.edge_map,
                   od.surf_map, od.qcap, od.knn_ids, od.knn_d2, od.knn_q, it > 0 ? 1 : 0);
    else
      FLOAM_LAUNCH(K_ASSOC_KNN, assoc_knn_kernel<false>, kKnnBlocks, kKnnThreads, s, S, od.ds_edge, od.d_nds_edge, od.ds_surf, od.d_nds_surf, od.edge_map,
                   od.surf_map, od.qcap, od.knn_ids, od.knn_d2, od.knn_q, it > 0 ? 1 : 0);
    FLOAM_LAUNCH(K_ASSOC_EVAL, assoc_eval_kernel, kAssocBlocks, kEvalThreads, s, S, od.ds_edge, od.d_nds_edge, od.ds_surf, od.d_nds_surf, od.edge_map, od.surf_map,
                 od.qcap, od.corr, od.corr_ok, od.knn_ids, od.loss, od.partials);
    FLOAM_LAUNCH_CLUSTER(K_LM_CLUSTER, lm_cluster_kernel, od.lm_cluster_ctas, kClusterThreads, kLmStageBytes, s, S, od.ds_edge, od.d_nds_edge, od.ds_surf, od.d_nds_surf, od.qcap, od.corr, od.corr_ok,
                 od.loss, od.partials, kAssocBlocks, FinishArgs{it + 1 == od.optimization_count ? 1 : 0, update_type, od.scan_period, od.traj, od.traj_cap});
  }
  if (od.optimization_count <= 0) FLOAM_LAUNCH(K_FINISH, finish_kernel, 1, 32, s, S, update_type, od.scan_period, od.traj, od.traj_cap);
  if (update_type == FLOAM_INITIAL_ITERATION) return;
  // addPointsToMap :253-294, predicated on the device-side keyframe decision
  const int* skip = &S->not_keyframe;
  LocalMap* maps[2] = {&od.surf_map, &od.edge_map};
  P4* dss[2] = {od.ds_surf, od.ds_edge};
  int* nds[2] = {od.d_nds_surf, od.d_nds_edge};
  const float leaf[2] = {od.leaf_surf, od.leaf_edge};
  cudaEventRecord(od.ev_fork, s);
  cudaStreamWaitEvent(a, od.ev_fork, 0);
  for (int k = 0; k < 2; ++k) {   // k = 0: surf map on the main branch; k = 1: edge map on the aux branch
    LocalMap& mp = *maps[k];
    cudaStream_t st = k == 0 ? s : a;
    VoxelWorkspace& ws = k == 0 ? *od.vws : *od.vws_aux;
    // One filter does it all: its first kernel appends the transformed features (:256-268) while it takes the bounding box, CropBox
    // (:270-287) is folded into the VoxelGrid (:289-292), and its last kernel leaves the bounding box of the new map for the search
    // grid. The filter gathers from mp.pts through the sorted index into mp.tmp; the grid's scatter kernel copies the cloud home.
    const VoxelAppend app{dss[k], S->x, &S->error_flags};
    // The map is the previous filter's output, i.e. already in voxel order but for a few re-voxelised centroids: once it is large only
    // the new points and the out-of-place ones are sorted and merged in (voxel_grid_merge_device); identical result either way.
    const bool large = odom_map_update_merges(od, k);   // the size hint that picks the merge path also widens the grids
    ws.large_input = large;
    if (large) voxel_grid_merge_device(mp.pts, 16, mp.d_n, mp.cap, leaf[k], mp.tmp, mp.d_n, ws, skip, st, S->crop_bounds, nds[k], mp.cap, &app, mp.bbox);
    else voxel_grid_device(mp.pts, 16, mp.d_n, mp.cap, leaf[k], mp.tmp, mp.d_n, ws, skip, st, S->crop_bounds, nds[k], mp.cap, &app, mp.bbox);
    rebuild_grid(od, mp, skip, st, &ws, true, &S->tl_end[k], large);
    ws.large_input = false;
  }
  cudaEventRecord(od.ev_join, a);
  cudaStreamWaitEvent(s, od.ev_join, 0);
}

void odom_downsample_device(OdomDevice& od, const void* d_edge, const int* d_ne, const void* d_surf, const int* d_ns, int stride, int n_max,
                            VoxelWorkspace& ws_surf, VoxelWorkspace& ws_edge, cudaStream_t s, cudaStream_t aux, cudaEvent_t ev_fork, cudaEvent_t ev_join) {
  // the edge and surf clouds are independent until the association: the edge side runs on the aux stream (a parallel branch of the
  // graph) with its own workspace
  cudaEventRecord(ev_fork, s);
  cudaStreamWaitEvent(aux, ev_fork, 0);
  voxel_grid_device(d_edge, stride, d_ne, n_max, od.leaf_edge, od.ds_edge, od.d_nds_edge, ws_edge, nullptr, aux);
  voxel_grid_device(d_surf, stride, d_ns, n_max, od.leaf_surf, od.ds_surf, od.d_nds_surf, ws_surf, nullptr, s);
  cudaEventRecord(ev_join, aux);
  cudaStreamWaitEvent(s, ev_join, 0);
}

void odom_select_buffers(OdomDevice& od, int parity) {
  od.ds_edge = od.ds_edge_b[parity]; od.ds_surf = od.ds_surf_b[parity];
  od.d_nds_edge = od.d_nds_edge_b[parity]; od.d_nds_surf = od.d_nds_surf_b[parity];
}

void compensate_velocity_device(OdomDevice& od, PointIRT* d_pts, const int* d_n, int n_max, cudaStream_t s) {
  FLOAM_LAUNCH(K_COMPENSATE_VELOCITY, compensate_velocity_kernel, grid_for(n_max), kThreads, s, d_pts, d_n, od.state,
               (od.fixes & FLOAM_FIX_ROTATED_VELOCITY) ? 1 : 0);
}

void compensate_velocity_explicit_device(PointIRT* d_pts, const int* d_n, int n_max, const double v[3], cudaStream_t s) {
  FLOAM_LAUNCH(K_COMPENSATE_VELOCITY, compensate_velocity_explicit_kernel, grid_for(n_max), kThreads, s, d_pts, d_n, v[0], v[1], v[2]);
}

void knn5_device(OdomDevice& od, LocalMap& map, const P4* d_queries, const int* d_nq, int nq_max, int* d_ids, float* d_d2, cudaStream_t s) {
  int g = (nq_max + kKnnThreads / 32 - 1) / (kKnnThreads / 32);
  if (g > kKnnBlocks) g = kKnnBlocks;
  if (g < 1) g = 1;
  FLOAM_LAUNCH(K_KNN5, knn5_kernel, g, kKnnThreads, s, d_queries, d_nq, map, d_ids, d_d2);
}

}  // namespace floam
