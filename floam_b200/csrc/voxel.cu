// Subsystem (2): pcl::VoxelGrid and pcl::CropBox on the device.
// VoxelGrid (reference call sites src/odomEstimationClass.cpp:137-142,289-292, src/laserMappingClass.cpp:175-184; PCL 1.8.1
// applyFilter restated in SURVEY.md Appendix A.1): bbox reduction -> voxel key per point (same float floor/int arithmetic)
// -> stable radix sort of (key, index) -> segment heads -> exclusive scan -> one thread per voxel accumulates xyz and
// intensity in float, in ascending point index, and divides by the count.  CropBox (src/odomEstimationClass.cpp:270-287,
// Appendix A.2): predicate -> scan -> order-preserving scatter.
#include <cstdlib>
#include <algorithm>
#include "voxel.cuh"

#include <algorithm>

#include "odom_math.cuh"

namespace floam {
namespace {

// largest grids of the two decoupled look-back kernels: CTAs per SM (occupancy calculator, voxel_workspace_arm) x SMs
int g_rank_grid_limit = kNumSMs, g_classify_grid_limit = kNumSMs;

constexpr int kThreads = 256;

__device__ __forceinline__ float4 load_xyzi(const char* base, int stride, int i) {
  const char* p = base + (size_t)i * stride;
  if (stride == 16) return __ldg(reinterpret_cast<const float4*>(p));
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float in = __ldg(reinterpret_cast<const float*>(p + 16));
  return make_float4(a.x, a.y, a.z, in);
}

// pcl::CropBox ahead of the VoxelGrid (addPointsToMap, src/odomEstimationClass.cpp:270-292) is folded into the filter itself: a point
// outside the box contributes nothing to the bounding box, gets a key past the last voxel (so the sort parks it at the end) and the
// run detection / centroid kernels only look at the first n_kept sorted entries. No separate crop pass, no scratch copy.
__device__ __forceinline__ bool crop_out(const float4 p, const float* __restrict__ b) {
  return (p.x < b[0] || p.y < b[1] || p.z < b[2]) || (p.x > b[3] || p.y > b[4] || p.z > b[5]);
}
// n = *d_n, plus *d_extra freshly appended points when they fitted into `cap` (addPointsToMap's push_backs)
__device__ __forceinline__ int total_count(const int* __restrict__ d_n, const int* __restrict__ d_extra, int cap) {
  int n = *d_n;
  if (d_extra && n + *d_extra <= cap) n += *d_extra;
  return n;
}

// counts[0] = points in the input (incl. d_extra), counts[1] = points kept by the crop (== counts[0] without a crop box)
__global__ void __launch_bounds__(kThreads) voxel_bbox_kernel(const char* in, int stride, const int* __restrict__ d_n,
                                                               const int* __restrict__ d_extra, int cap, const float* __restrict__ crop,
                                                               unsigned int* __restrict__ bbox, int* __restrict__ counts, VoxelAppend app,
                                                               unsigned long long* __restrict__ merge_state, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  const int n = total_count(d_n, d_extra, cap);
  const int n_old = *d_n;
  if (merge_state)   // look-back states of voxel_classify_kernel (two words per kScanTile points), armed here
    for (int t = blockIdx.x * kThreads + threadIdx.x; t * kScanTile < n; t += gridDim.x * kThreads) { merge_state[2 * t] = 0ull; merge_state[2 * t + 1] = 0ull; }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts[0] = n;
    if (app.src && n_old + *d_extra > cap) atomicOr(app.err_flags, 1);   // the new points do not fit: they are dropped, the flag is sticky
  }
  double x[7];
  if (app.src) {
#pragma unroll
    for (int k = 0; k < 7; ++k) x[k] = app.pose7[k];
  }
  float mn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, mx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
  int kept = 0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    float4 p;
    if (app.src && i >= n_old) {   // addPointsToMap :256-268: pointAssociateToMap (double transform, float store) and push_back
      const float4 q = __ldg(app.src + (i - n_old));
      const m::V3 w = m::add(m::quat_rotate(x, m::V3{(double)q.x, (double)q.y, (double)q.z}), m::V3{x[4], x[5], x[6]});
      p = make_float4((float)w.x, (float)w.y, (float)w.z, q.w);
      *reinterpret_cast<float4*>(const_cast<char*>(in) + (size_t)i * stride) = p;   // stride is 16 for the map clouds
    } else {
      p = load_xyzi(in, stride, i);
    }
    if (crop && crop_out(p, crop)) continue;
    ++kept;
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
  // one set of atomics per CTA, not per warp: the seven accumulators are single addresses every voting warp would queue on
  __shared__ float s_mn[kThreads / 32][3], s_mx[kThreads / 32][3];
  __shared__ int s_kept[kThreads / 32];
  const int w = warp_id(), l = lane_id();
  if (l == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { s_mn[w][a] = mn[a]; s_mx[w][a] = mx[a]; }
    s_kept[w] = kept;
  }
  __syncthreads();
  if (w == 0) {
    int k = l < kThreads / 32 ? s_kept[l] : 0;
    float vmn[3], vmx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      vmn[a] = l < kThreads / 32 ? s_mn[l][a] : 3.402823466e38f;
      vmx[a] = l < kThreads / 32 ? s_mx[l][a] : -3.402823466e38f;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      k += __shfl_xor_sync(0xffffffffu, k, o);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        vmn[a] = fminf(vmn[a], __shfl_xor_sync(0xffffffffu, vmn[a], o));
        vmx[a] = fmaxf(vmx[a], __shfl_xor_sync(0xffffffffu, vmx[a], o));
      }
    }
    if (l == 0 && k > 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        atomicMin(&bbox[a], float_flip(vmn[a]));
        atomicMax(&bbox[3 + a], float_flip(vmx[a]));
      }
      atomicAdd(&counts[1], k);
    }
  }
}

__device__ __forceinline__ int bits_for(long long cells) {  // number of key bits for indices in [0, cells)
  int b = 0;
  while (b < 32 && (1ll << b) < cells) ++b;
  return b;
}

__global__ void __launch_bounds__(kThreads) voxel_keys_kernel(const char* __restrict__ in, int stride, const int* __restrict__ counts, float leaf,
                                                               const float* __restrict__ crop, const unsigned int* __restrict__ bbox,
                                                               unsigned int* __restrict__ keys, int* __restrict__ vals, int* d_nbits, int* d_passthrough,
                                                               int* __restrict__ tile_state, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  const int n = counts[0];
  // look-back states of voxel_rank_kernel (one per kScanTile sorted keys), armed here
  for (int t = blockIdx.x * kThreads + threadIdx.x; t * kScanTile < n; t += gridDim.x * kThreads) tile_state[t] = 0;
  const bool none = counts[1] == 0;   // nothing survives the crop (or the input is empty): the bounding box is undefined
  // every thread derives the grid from the bbox exactly like VoxelGrid::applyFilter (float arithmetic, no contraction)
  const float inv = __fdiv_rn(1.0f, leaf);
  float mnp[3], mxp[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) { mnp[a] = none ? 0.f : float_unflip(bbox[a]); mxp[a] = none ? 0.f : float_unflip(bbox[3 + a]); }
  const long long dx = (long long)fmul(fsub(mxp[0], mnp[0]), inv) + 1;
  const long long dy = (long long)fmul(fsub(mxp[1], mnp[1]), inv) + 1;
  const long long dz = (long long)fmul(fsub(mxp[2], mnp[2]), inv) + 1;
  const bool pass = (dx * dy * dz) > 2147483647ll;  // "Leaf size is too small for the input dataset" -> output = input (Q13)
  int min_b[3], div_b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int)floorf(fmul(mnp[a], inv));
    const int max_b = (int)floorf(fmul(mxp[a], inv));
    div_b[a] = max_b - min_b[a] + 1;
  }
  const int mul1 = div_b[0], mul2 = div_b[0] * div_b[1];
  // cropped-out points sort behind every voxel: key = number of voxels (or n in pass-through mode)
  const long long ncells = (long long)div_b[0] * div_b[1] * div_b[2];
  const unsigned int parked = pass ? (unsigned int)n : (unsigned int)ncells;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *d_passthrough = pass ? 1 : 0;
    *d_nbits = pass ? bits_for((long long)n + 1) : bits_for(ncells + 1);
  }
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    unsigned int key;
    const float4 p = load_xyzi(in, stride, i);
    if (crop && crop_out(p, crop)) {
      key = parked;
    } else if (pass) {
      key = (unsigned int)i;  // every point its own voxel, order kept: the centroid of one point is the point
    } else {
      const int ijk0 = (int)fsub(floorf(fmul(p.x, inv)), (float)min_b[0]);
      const int ijk1 = (int)fsub(floorf(fmul(p.y, inv)), (float)min_b[1]);
      const int ijk2 = (int)fsub(floorf(fmul(p.z, inv)), (float)min_b[2]);
      key = (unsigned int)(ijk0 + ijk1 * mul1 + ijk2 * mul2);
    }
    keys[i] = key;
    vals[i] = i;
  }
}

// ---- after the sort: voxel runs -> output ranks -> centroids, three kernels -------------------------------------------------
// (1) per 4096-key tile: number of run heads; (2) every tile adds up the tiles before it, ranks its heads and records where
// each run starts; (3) one thread per voxel walks its run [head_pos[v], head_pos[v+1]) — bounds known up front, so the loads of a
// run are independent of the loop control and pipeline.
__device__ __forceinline__ int is_head(const unsigned int* __restrict__ keys, int i) { return (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0; }

// Decoupled look-back by one WARP: the states of up to 32 predecessor tiles are fetched at once (a single thread walking back tile by
// tile pays one dependent L2 round trip per tile: with 15 tiles that walk WAS the kernel). A state word is payload << 2 | flag,
// flag 1 = the tile's own aggregate, 2 = inclusive prefix up to and including the tile, 0 = not published yet. Returns the combination
// of the payloads of all tiles before `tile`. All 32 lanes of the calling warp take part.
template <typename W, typename Op>
__device__ __forceinline__ W warp_lookback(const W* states, int stride, int tile, W identity, Op combine) {
  const int l = lane_id();
  W acc = identity;
  for (int t0 = tile - 1; t0 >= 0; t0 -= 32) {
    const int t = t0 - l;
    W v = (identity << 2) | (W)2;                 // before tile 0: an inclusive prefix of nothing
    if (t >= 0) {
      const volatile W* p = states + (size_t)stride * t;
      while (((v = *p) & (W)3) == (W)0) __nanosleep(20);
    }
    const unsigned int incl = __ballot_sync(0xffffffffu, (v & (W)3) == (W)2);
    const int first = __ffs(incl) - 1;            // lane of the nearest tile that already knows its inclusive prefix (there is one at the latest before tile 0)
    W part = (first < 0 || l <= first) ? (v >> 2) : identity;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part = combine(part, __shfl_xor_sync(0xffffffffu, part, o));
    acc = combine(acc, part);
    if (first >= 0) break;
  }
  return acc;
}

// Single pass: every tile counts and ranks its run heads, then finds the number of heads before it by decoupled look-back over the
// tiles' published states (state = value << 2 | 1: this tile's own count; value << 2 | 2: inclusive prefix up to this tile). The
// states are zeroed by voxel_keys_kernel, two launches earlier. Lower-numbered tiles are dispatched first, so the wait is short.
__global__ void __launch_bounds__(kScanThreads) voxel_rank_kernel(const unsigned int* __restrict__ keys, const int* __restrict__ d_n,
                                                                  int* tile_state, int* __restrict__ head_pos, int* d_nout, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  const int n = *d_n;
  if (n == 0) { if (blockIdx.x == 0 && threadIdx.x == 0) { *d_nout = 0; head_pos[0] = 0; } return; }
  __shared__ int smem[33];
  __shared__ int s_offset;
  const int ntiles = (n + kScanTile - 1) / kScanTile;
  // One wave of CTAs, each taking tiles b, b + gridDim.x, ...: a tile only ever waits for lower-numbered tiles, and the grid is never
  // larger than what the device keeps resident at once (g_rank_grid_limit, from the occupancy calculator), so every CTA a waiting
  // tile depends on is running or done.
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int base = tile * kScanTile + threadIdx.x * kScanItems;
    int h[kScanItems];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { h[k] = (base + k < n) ? is_head(keys, base + k) : 0; sum += h[k]; }
    int total;
    int rank = block_excl_scan(sum, smem, &total);
    if (warp_id() == 0) {
      volatile int* st = tile_state;
      if (tile > 0 && threadIdx.x == 0) st[tile] = (total << 2) | 1;
      const int before = warp_lookback<int>(tile_state, 1, tile, 0, [](int a, int b) { return a + b; });
      if (threadIdx.x == 0) {
        st[tile] = ((before + total) << 2) | 2;
        s_offset = before;
      }
    }
    __syncthreads();
    const int offset = s_offset;
    rank += offset;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (h[k]) head_pos[rank++] = base + k;
    if (tile == ntiles - 1 && threadIdx.x == 0) { *d_nout = offset + total; head_pos[offset + total] = n; }
    __syncthreads();   // s_offset / smem are reused by the next tile
  }
}

// pcl::CentroidPoint<PointXYZI>: float sums in run order (ascending input index), then / n. The ORDER of the additions is fixed, so
// a run is a serial chain of float adds — but its loads are not. Runs are mostly ~10 points and reach a thousand (ground rings next
// to the sensor; one such run took 160 us when a single thread also did its loads):
//   phase 1: one thread per voxel for runs <= 32 points, loads issued eight at a time (index, then point);
//   phase 2: one WARP per long run: the lanes gather 256 points per step into shared memory (coalesced index reads, parallel
//            point loads), lane 0 adds them up in order from shared memory.
constexpr int kShortRun = 32;
constexpr int kStage = 256;
__global__ void __launch_bounds__(kThreads) voxel_reduce_kernel(const char* __restrict__ in, int stride, const int* __restrict__ vals,
                                                                 const int* __restrict__ head_pos, const int* __restrict__ d_nout, P4* __restrict__ out,
                                                                 unsigned int* bbox, int* counts, unsigned int* out_bbox, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  __shared__ float4 s_stage[kThreads / 32][kStage];
  const int nv = *d_nout;
  float omn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, omx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
  bool wrote = false;
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // last kernel of the filter: re-arm the accumulators for the next one
    bbox[0] = bbox[1] = bbox[2] = 0xffffffffu;
    bbox[3] = bbox[4] = bbox[5] = 0u;
    counts[1] = 0;
  }
  for (int v = blockIdx.x * kThreads + threadIdx.x; v < nv; v += gridDim.x * kThreads) {
    const int b = __ldg(head_pos + v), e = __ldg(head_pos + v + 1);
    if (e - b > kShortRun) continue;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (int j = b; j < e; j += 8) {
      int idx[8];
      float4 p[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) idx[u] = (j + u < e) ? __ldg(vals + j + u) : -1;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (idx[u] >= 0) p[u] = load_xyzi(in, stride, idx[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (idx[u] >= 0) { sx = fadd(sx, p[u].x); sy = fadd(sy, p[u].y); sz = fadd(sz, p[u].z); si = fadd(si, p[u].w); }
    }
    const float cnt = (float)(e - b);
    const float4 c = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), __fdiv_rn(si, cnt));
    out[v] = c;
    omn[0] = fminf(omn[0], c.x); omn[1] = fminf(omn[1], c.y); omn[2] = fminf(omn[2], c.z);
    omx[0] = fmaxf(omx[0], c.x); omx[1] = fmaxf(omx[1], c.y); omx[2] = fmaxf(omx[2], c.z);
    wrote = true;
  }
  const int w = warp_id(), l = lane_id();
  const int warps_total = gridDim.x * (kThreads / 32);
  for (int v = blockIdx.x * (kThreads / 32) + w; v < nv; v += warps_total) {
    const int b = __ldg(head_pos + v), e = __ldg(head_pos + v + 1);
    if (e - b <= kShortRun) continue;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (int c = b; c < e; c += kStage) {
      int idx[kStage / 32];
#pragma unroll
      for (int u = 0; u < kStage / 32; ++u) { const int j = c + u * 32 + l; idx[u] = j < e ? __ldg(vals + j) : -1; }
#pragma unroll
      for (int u = 0; u < kStage / 32; ++u)
        if (idx[u] >= 0) s_stage[w][u * 32 + l] = load_xyzi(in, stride, idx[u]);
      __syncwarp();
      if (l == 0) {
        const int m = min(kStage, e - c);
#pragma unroll 8
        for (int i = 0; i < m; ++i) {
          const float4 p = s_stage[w][i];
          sx = fadd(sx, p.x); sy = fadd(sy, p.y); sz = fadd(sz, p.z); si = fadd(si, p.w);
        }
      }
      __syncwarp();
    }
    if (l == 0) {
      const float cnt = (float)(e - b);
      const float4 c = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), __fdiv_rn(si, cnt));
      out[v] = c;
      omn[0] = fminf(omn[0], c.x); omn[1] = fminf(omn[1], c.y); omn[2] = fminf(omn[2], c.z);
      omx[0] = fmaxf(omx[0], c.x); omx[1] = fmaxf(omx[1], c.y); omx[2] = fmaxf(omx[2], c.z);
      wrote = true;
    }
  }
  // bounding box of the output cloud for the caller's search grid: warp, then CTA reduction, one set of atomics per CTA
  if (out_bbox) {   // uniform: kernel argument
    float* s_box = reinterpret_cast<float*>(&s_stage[0][0]);   // the staging area is idle now: [warp][6]
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        omn[a] = fminf(omn[a], __shfl_xor_sync(0xffffffffu, omn[a], o));
        omx[a] = fmaxf(omx[a], __shfl_xor_sync(0xffffffffu, omx[a], o));
      }
    }
    const bool any = __any_sync(0xffffffffu, wrote);
    if (l == 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { s_box[w * 6 + a] = any ? omn[a] : 3.402823466e38f; s_box[w * 6 + 3 + a] = any ? omx[a] : -3.402823466e38f; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      const int a = threadIdx.x;
      float v = s_box[a];
#pragma unroll
      for (int ww = 1; ww < kThreads / 32; ++ww) v = a < 3 ? fminf(v, s_box[ww * 6 + a]) : fmaxf(v, s_box[ww * 6 + a]);
      if (a < 3) { if (v != 3.402823466e38f) atomicMin(&out_bbox[a], float_flip(v)); }
      else if (v != -3.402823466e38f) atomicMax(&out_bbox[a], float_flip(v));
    }
  }
}

// ---- keyframe update of an (almost) sorted map: classify, sort only what is out of place, merge -----------------------------------
// addPointsToMap (src/odomEstimationClass.cpp:253-294) filters map + new frame, but the map IS the previous filter's output: its
// points are already in ascending voxel order (the order is the lexicographic (kz, ky, kx) of the absolute voxel coordinates, it does
// not depend on the bounding box) except where a stored centroid re-voxelises one cell over (a centroid within an ulp of a voxel
// face), and only the Q new points are anywhere. Sorting all M + Q keys again is what the reference does; here
//   voxel_classify_kernel  computes every key like voxel_keys_kernel and splits the surviving points, order kept, into a MAIN list
//                          that is non-decreasing by construction and a SIDE list (everything else: new points, old points that are
//                          out of place). An old point joins MAIN iff it is not the start of a descent (key > next old key) and its key
//                          is >= the largest key of all such points before it (prefix maximum, single pass with decoupled look-back):
//                          for MAIN points a < b, a is one of the points b was compared with, so key_a <= key_b — whatever the input.
//   radix sort             of the SIDE list only (stable, so old out-of-place points stay ahead of new points of the same voxel);
//   voxel_merge_kernel     merges the two lists by (key, input index) - each is sorted by that pair - with one search per point:
//                          exactly the stable sort of the whole input by key.
// The run ranking and the centroid kernel then see the same (key, index) arrays the full sort would have produced. An unsorted input
// (first frame: raw features) just ends up mostly in SIDE: correct, and no slower than before.
constexpr int kMergeThreads = 256;
constexpr int kMainSampleShift = 8;     // every 256th MAIN key and
constexpr int kSideSampleShift = 5;     // every 32nd SIDE key are staged in shared memory: a search is a few shared-memory steps + one short global run
constexpr int kMaxSamples = 8192;       // per table (2M MAIN keys, 262k SIDE keys); beyond that the search runs in global memory

__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }

__global__ void __launch_bounds__(kScanThreads) voxel_classify_kernel(const char* __restrict__ in, int stride, const int* __restrict__ d_n_old,
                                                                      const int* __restrict__ counts, float leaf, const float* __restrict__ crop,
                                                                      const unsigned int* __restrict__ bbox, unsigned int* __restrict__ main_keys,
                                                                      int* __restrict__ main_vals, unsigned int* __restrict__ side_keys,
                                                                      int* __restrict__ side_vals, int* __restrict__ d_nms, int* d_nbits,
                                                                      int* d_passthrough, unsigned long long* mstate, int* __restrict__ rank_state,
                                                                      const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  const int n = counts[0];
  const int n_old = min(*d_n_old, n);
  for (int t = blockIdx.x * kScanThreads + threadIdx.x; t * kScanTile < n; t += gridDim.x * kScanThreads) rank_state[t] = 0;   // voxel_rank_kernel's look-back
  const bool none = counts[1] == 0;
  const float inv = __fdiv_rn(1.0f, leaf);
  float mnp[3], mxp[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) { mnp[a] = none ? 0.f : float_unflip(bbox[a]); mxp[a] = none ? 0.f : float_unflip(bbox[3 + a]); }
  const long long dx = (long long)fmul(fsub(mxp[0], mnp[0]), inv) + 1;
  const long long dy = (long long)fmul(fsub(mxp[1], mnp[1]), inv) + 1;
  const long long dz = (long long)fmul(fsub(mxp[2], mnp[2]), inv) + 1;
  const bool pass = (dx * dy * dz) > 2147483647ll;  // Q13
  int min_b[3], div_b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int)floorf(fmul(mnp[a], inv));
    const int max_b = (int)floorf(fmul(mxp[a], inv));
    div_b[a] = max_b - min_b[a] + 1;
  }
  const int mul1 = div_b[0], mul2 = div_b[0] * div_b[1];
  const long long ncells = (long long)div_b[0] * div_b[1] * div_b[2];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *d_passthrough = pass ? 1 : 0;
    *d_nbits = pass ? bits_for((long long)n + 1) : bits_for(ncells + 1);
    if (n == 0) { d_nms[0] = 0; d_nms[1] = 0; }
  }
  __shared__ int s_scan[33];
  __shared__ unsigned int s_wmax[kScanThreads / 32];
  __shared__ unsigned int s_tile_pmax;
  __shared__ int s_before[2];
  const int w = warp_id(), l = lane_id();
  const int ntiles = (n + kScanTile - 1) / kScanTile;
  // A tile waits for lower-numbered tiles only, and the grid is never larger than what the device keeps resident at once
  // (voxel_workspace_arm asks the occupancy calculator): every CTA a waiting tile depends on is running or done.
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int base = tile * kScanTile + threadIdx.x * kScanItems;
    // keys of this thread's four consecutive points and of the one behind them (the descent test looks one point ahead)
    unsigned int key[kScanItems + 1];
    bool dropped[kScanItems + 1];
#pragma unroll
    for (int k = 0; k <= kScanItems; ++k) {
      const int i = base + k;
      key[k] = 0u; dropped[k] = true;
      if (i < n) {
        const float4 p = load_xyzi(in, stride, i);
        if (!(crop && crop_out(p, crop))) {
          dropped[k] = false;
          if (pass) {
            key[k] = (unsigned int)i;
          } else {
            const int ijk0 = (int)fsub(floorf(fmul(p.x, inv)), (float)min_b[0]);
            const int ijk1 = (int)fsub(floorf(fmul(p.y, inv)), (float)min_b[1]);
            const int ijk2 = (int)fsub(floorf(fmul(p.z, inv)), (float)min_b[2]);
            key[k] = (unsigned int)(ijk0 + ijk1 * mul1 + ijk2 * mul2);
          }
        }
      }
    }
    // eligible: an old point that survives the crop and does not start a descent
    bool elig[kScanItems];
    unsigned int tmax = 0u;      // largest eligible key of this thread's points
    bool tany = false;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const int i = base + k;
      const bool next_old = i + 1 < n_old && !dropped[k + 1];
      elig[k] = i < n_old && !dropped[k] && !(next_old && key[k] > key[k + 1]);
      if (elig[k]) { tmax = tany ? max(tmax, key[k]) : key[k]; tany = true; }
    }
    // exclusive prefix maximum over the threads of the tile (keys are >= 0, so "nothing before" is 0)
    unsigned int incl = tany ? tmax : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (l >= o) incl = max(incl, t);
    }
    unsigned int excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (l == 0) excl = 0u;
    __syncthreads();   // the shared scalars of the previous tile have been read
    if (l == 31) s_wmax[w] = incl;
    __syncthreads();
    if (w == 0) {
      unsigned int v = s_wmax[l], vi = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, vi, o);
        if (l >= o) vi = max(vi, t);
      }
      unsigned int ve = __shfl_up_sync(0xffffffffu, vi, 1);
      if (l == 0) ve = 0u;
      s_wmax[l] = ve;                       // exclusive prefix maximum of the warps before warp l
      // ---- look-back 1: the largest eligible key of every tile before this one ----
      const unsigned int mine = __shfl_sync(0xffffffffu, vi, 31);   // this tile's own maximum
      if (tile > 0 && l == 0) st_state(&mstate[2 * tile], ((unsigned long long)mine << 2) | 1ull);
      const unsigned long long before = warp_lookback<unsigned long long>(mstate, 2, tile, 0ull, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; });
      if (l == 0) {
        st_state(&mstate[2 * tile], ((unsigned long long)max((unsigned int)before, mine) << 2) | 2ull);
        s_tile_pmax = (unsigned int)before;
      }
    }
    __syncthreads();
    unsigned int pm = max(max(s_tile_pmax, s_wmax[w]), excl);   // largest eligible key before this thread's first point
    int is_main[kScanItems], is_side[kScanItems], packed = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const bool m = elig[k] && key[k] >= pm;
      if (elig[k]) pm = max(pm, key[k]);
      is_main[k] = m ? 1 : 0;
      is_side[k] = (!dropped[k] && !m) ? 1 : 0;     // dropped[k] is also set for i >= n
      packed += is_main[k] + (is_side[k] << 16);
    }
    int total;
    const int ex = block_excl_scan(packed, s_scan, &total);
    if (w == 0) {
      // ---- look-back 2: MAIN and SIDE points of every tile before this one (two 31-bit counts in one payload) ----
      const unsigned long long mine = ((unsigned long long)(total & 0xffff) << 31) | (unsigned long long)(total >> 16);
      if (tile > 0 && l == 0) st_state(&mstate[2 * tile + 1], (mine << 2) | 1ull);
      const unsigned long long before = warp_lookback<unsigned long long>(mstate + 1, 2, tile, 0ull, [](unsigned long long a, unsigned long long b) { return a + b; });
      if (l == 0) {
        st_state(&mstate[2 * tile + 1], ((before + mine) << 2) | 2ull);
        s_before[0] = (int)(before >> 31);
        s_before[1] = (int)(before & 0x7fffffffull);
        if (tile == ntiles - 1) { d_nms[0] = s_before[0] + (total & 0xffff); d_nms[1] = s_before[1] + (total >> 16); }
      }
    }
    __syncthreads();
    int mpos = s_before[0] + (ex & 0xffff), spos = s_before[1] + (ex >> 16);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      if (is_main[k]) { main_keys[mpos] = key[k]; main_vals[mpos] = base + k; ++mpos; }
      if (is_side[k]) { side_keys[spos] = key[k]; side_vals[spos] = base + k; ++spos; }
    }
  }
}

// number of entries of a list sorted by (key, index) that come before (key, idx); the key part is narrowed by a shared-memory table of
// every (1 << shift)-th key first (n_samples == 0: no table), the index part only matters inside the run of equal keys
__device__ __forceinline__ int rank_in(const unsigned int* __restrict__ keys, const int* __restrict__ vals, int n, const unsigned int* s_samples,
                                       int n_samples, int shift, unsigned int key, int idx) {
  int lo = 0, hi = n;
  if (n_samples > 0) {
    int a = 0, b = n_samples;     // samples [0, a) are below the key
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (s_samples[mid] < key) a = mid + 1; else b = mid;
    }
    // sample a - 1 (entry (a - 1) << shift) is below the key, sample a is not: the first entry not below the key lies in between
    lo = a == 0 ? 0 : ((a - 1) << shift) + 1;
    hi = a == n_samples ? n : (a << shift);
  }
  while (lo < hi) {               // first entry whose key is not below `key`
    const int mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
  }
  if (lo >= n || __ldg(keys + lo) != key) return lo;
  // a run of equal keys starts here: entries of it with a smaller index come first as well (indices ascend inside a run)
  int end = lo + 1, step = 1;
  while (end < n && __ldg(keys + end) == key) { end = min(n, end + step); step <<= 1; }   // some entry at or past the end of the run
  hi = end;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) == key && __ldg(vals + mid) < idx) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Merge by (key, index): both lists are sorted by it (MAIN is a subsequence of the input with non-decreasing keys, SIDE comes out of a
// stable sort), so the result is the stable sort of the whole input by key, whatever the classification put where.
__global__ void __launch_bounds__(kMergeThreads) voxel_merge_kernel(const unsigned int* __restrict__ main_keys, const int* __restrict__ main_vals,
                                                                    const unsigned int* __restrict__ side_keys, const int* __restrict__ side_vals,
                                                                    const int* __restrict__ d_nms, unsigned int* __restrict__ out_keys,
                                                                    int* __restrict__ out_vals, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  const int n_main = d_nms[0], n_side = d_nms[1];
  extern __shared__ unsigned int s_tab[];      // [kMaxSamples] SIDE samples, then [kMaxSamples] MAIN samples
  unsigned int* s_side = s_tab;
  unsigned int* s_main = s_tab + kMaxSamples;
  int ns_side = (n_side + (1 << kSideSampleShift) - 1) >> kSideSampleShift;
  int ns_main = (n_main + (1 << kMainSampleShift) - 1) >> kMainSampleShift;
  if (ns_side > kMaxSamples) ns_side = 0;      // table too small: plain search
  if (ns_main > kMaxSamples) ns_main = 0;
  const bool has_main_work = (int)(blockIdx.x * kMergeThreads) < n_main, has_side_work = (int)(blockIdx.x * kMergeThreads) < n_side;
  if (has_main_work) for (int t = threadIdx.x; t < ns_side; t += kMergeThreads) s_side[t] = __ldg(side_keys + ((size_t)t << kSideSampleShift));
  if (has_side_work) for (int t = threadIdx.x; t < ns_main; t += kMergeThreads) s_main[t] = __ldg(main_keys + ((size_t)t << kMainSampleShift));
  __syncthreads();
  for (int i = blockIdx.x * kMergeThreads + threadIdx.x; i < n_main; i += gridDim.x * kMergeThreads) {
    const unsigned int key = __ldg(main_keys + i);
    const int idx = __ldg(main_vals + i);
    const int pos = i + (n_side ? rank_in(side_keys, side_vals, n_side, s_side, ns_side, kSideSampleShift, key, idx) : 0);
    out_keys[pos] = key; out_vals[pos] = idx;
  }
  for (int j = blockIdx.x * kMergeThreads + threadIdx.x; j < n_side; j += gridDim.x * kMergeThreads) {
    const unsigned int key = __ldg(side_keys + j);
    const int idx = __ldg(side_vals + j);
    const int pos = j + (n_main ? rank_in(main_keys, main_vals, n_main, s_main, ns_main, kMainSampleShift, key, idx) : 0);
    out_keys[pos] = key; out_vals[pos] = idx;
  }
}

// ---- CropBox: count per tile, then rank + scatter with the predicate recomputed (two kernels) -----------------------------------
__device__ __forceinline__ int crop_keep(const float4 p, const float* __restrict__ b) {
  const bool outside = (p.x < b[0] || p.y < b[1] || p.z < b[2]) || (p.x > b[3] || p.y > b[4] || p.z > b[5]);
  return outside ? 0 : 1;
}

// n = *d_n, plus *d_extra freshly appended points when they fitted into `cap` (addPointsToMap's push_backs)
__device__ __forceinline__ int crop_count(const int* __restrict__ d_n, const int* __restrict__ d_extra, int cap) {
  int n = *d_n;
  if (d_extra && n + *d_extra <= cap) n += *d_extra;
  return n;
}

__global__ void __launch_bounds__(kScanThreads) crop_flags_kernel(const P4* __restrict__ in, const int* __restrict__ d_n, const int* __restrict__ d_extra, int cap,
                                                                  const float* __restrict__ bounds, int* __restrict__ tile_sums, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  const int n = crop_count(d_n, d_extra, cap);
  if (blockIdx.x * kScanTile >= n) return;
  __shared__ int smem[33];
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) sum += (base + k < n) ? crop_keep(__ldg(in + base + k), bounds) : 0;
  const int total = block_sum(sum, smem);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) crop_scatter_kernel(const P4* __restrict__ in, const int* __restrict__ d_n, const int* __restrict__ d_extra, int cap,
                                                                    const float* __restrict__ bounds, const int* __restrict__ tile_sums, P4* __restrict__ out,
                                                                    int* d_nout, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  const int n = crop_count(d_n, d_extra, cap);
  if (n == 0) { if (blockIdx.x == 0 && threadIdx.x == 0) *d_nout = 0; return; }
  if (blockIdx.x * kScanTile >= n) return;
  __shared__ int smem[33];
  const int offset = tile_offset(tile_sums, blockIdx.x, smem);
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  float4 p[kScanItems];
  int keep[kScanItems];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    keep[k] = 0;
    if (base + k < n) { p[k] = __ldg(in + base + k); keep[k] = crop_keep(p[k], bounds); }
    sum += keep[k];
  }
  int total;
  int pos = block_excl_scan(sum, smem, &total) + offset;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (keep[k]) out[pos++] = p[k];
  if (blockIdx.x == (n - 1) / kScanTile && threadIdx.x == 0) *d_nout = offset + total;
}

__global__ void __launch_bounds__(kThreads) repack_kernel(const char* __restrict__ in, const int* __restrict__ d_n, P4* __restrict__ out) {
  pdl_prologue();
  const int n = *d_n;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) out[i] = load_xyzi(in, 32, i);
}

// The kernels stride; empty CTAs of a capacity-sized grid are not free. Default two CTAs of 256 threads per SM: measured against four,
// one sequence runs as fast (5.43k frames/s either way) and four sequences sharing the GPU gain 6.5 % (9.6k -> 10.2k): fewer resident
// CTAs of one sequence's kernels in the way of the others'. Large inputs (the >= 250k-point maps that also take the merge path) keep
// four per SM: at configs[3] two cost 6.5 % (0.653 -> 0.696 ms per frame).
inline int grid_for(int n_max, int ctas_per_sm = 2) {
  int g = (n_max + kThreads - 1) / kThreads;
  const int cap = kNumSMs * ctas_per_sm;
  return g < 1 ? 1 : (g > cap ? cap : g);
}

}  // namespace

size_t voxel_workspace_bytes(int n_max) {
  return (size_t)n_max * 4 * 4 + ((size_t)n_max + 1) * 4 + 1024 + sort_workspace_bytes(n_max) + scan_workspace_bytes(n_max + 1) +
         ((size_t)n_max / kScanTile + 2) * 16 + 8192;
}

void voxel_workspace_bind(VoxelWorkspace& ws, void* mem, int n_max) {
  char* p = (char*)mem;
  auto take = [&](size_t bytes) { void* r = p; p += (bytes + 255) / 256 * 256; return r; };
  ws.n_max = n_max;
  ws.keys = (unsigned int*)take((size_t)n_max * 4);
  ws.vals = (int*)take((size_t)n_max * 4);
  ws.flags = (int*)take(((size_t)n_max + 1) * 4);
  ws.main_keys = (unsigned int*)take((size_t)n_max * 4);
  ws.main_vals = (int*)take((size_t)n_max * 4);
  ws.merge_state = (unsigned long long*)take(((size_t)n_max / kScanTile + 2) * 16);
  ws.bbox = (unsigned int*)take(32);
  ws.d_nbits = (int*)take(4);
  ws.d_passthrough = (int*)take(4);
  ws.d_counts = (int*)take(32);
  sort_workspace_bind(ws.sort, take(sort_workspace_bytes(n_max)), n_max);
  scan_workspace_bind(ws.scan, take(scan_workspace_bytes(n_max + 1)), n_max + 1);
}

int voxel_workspace_arm(VoxelWorkspace& ws, cudaStream_t s) {
  const unsigned int bb[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  FLOAM_CUDA_OK(cudaMemcpyAsync(ws.bbox, bb, sizeof(bb), cudaMemcpyHostToDevice, s));
  if (sort_workspace_arm(ws.sort, s)) return FLOAM_ERR_CUDA;
  FLOAM_CUDA_OK(cudaMemsetAsync(ws.d_counts, 0, 32, s));
  FLOAM_CUDA_OK(cudaFuncSetAttribute(voxel_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * kMaxSamples * sizeof(unsigned int))));
  int per_sm = 0;
  FLOAM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, voxel_rank_kernel, kScanThreads, 0));
  g_rank_grid_limit = std::max(1, per_sm) * kNumSMs;
  FLOAM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, voxel_classify_kernel, kScanThreads, 0));
  g_classify_grid_limit = std::max(1, per_sm) * kNumSMs;
  FLOAM_CUDA_OK(cudaStreamSynchronize(s));
  return FLOAM_OK;
}

void voxel_grid_device(const void* d_in, int stride_bytes, const int* d_n, int n_max, float leaf, P4* d_out, int* d_nout, VoxelWorkspace& ws,
                       const int* d_skip, cudaStream_t s, const float* d_crop, const int* d_extra, int cap, const VoxelAppend* append,
                       unsigned int* out_bbox) {
  const VoxelAppend app = append ? *append : VoxelAppend{nullptr, nullptr, nullptr};
  if (n_max > ws.n_max) n_max = ws.n_max;
  const char* in = (const char*)d_in;
  const int g = grid_for(n_max, ws.large_input ? 4 : 2);
  int gt = (n_max + kScanTile - 1) / kScanTile;
  if (gt > g_rank_grid_limit) gt = g_rank_grid_limit;   // voxel_rank_kernel loops over tiles; all of its CTAs must be able to be resident together
  int* counts = ws.d_counts;   // [0] input points, [1] points kept by the crop
  FLOAM_LAUNCH(K_VOXEL_BBOX, voxel_bbox_kernel, g, kThreads, s, in, stride_bytes, d_n, d_extra, cap, d_crop, ws.bbox, counts, app, (unsigned long long*)nullptr, d_skip);
  FLOAM_LAUNCH(K_VOXEL_KEYS, voxel_keys_kernel, g, kThreads, s, in, stride_bytes, counts, leaf, d_crop, ws.bbox, ws.keys, ws.vals, ws.d_nbits, ws.d_passthrough,
               ws.scan.block_sums, d_skip);
  unsigned int* skeys = nullptr;
  int* svals = nullptr;
  radix_sort_pairs(ws.keys, ws.vals, counts, ws.d_nbits, n_max, ws.sort, d_skip, s, &skeys, &svals);
  FLOAM_LAUNCH(K_VOXEL_RANK, voxel_rank_kernel, gt, kScanThreads, s, skeys, counts + 1, ws.scan.block_sums, ws.flags, d_nout, d_skip);
  const int gr = g < 3 * kNumSMs ? g : 3 * kNumSMs;   // 80 registers + 33 KB shared memory: three CTAs per SM make one wave; the kernel strides
  FLOAM_LAUNCH(K_VOXEL_REDUCE, voxel_reduce_kernel, gr, kThreads, s, in, stride_bytes, svals, ws.flags, d_nout, d_out, ws.bbox, counts, out_bbox, d_skip);
}

void voxel_grid_merge_device(const void* d_in, int stride_bytes, const int* d_n, int n_max, float leaf, P4* d_out, int* d_nout, VoxelWorkspace& ws,
                             const int* d_skip, cudaStream_t s, const float* d_crop, const int* d_extra, int cap, const VoxelAppend* append,
                             unsigned int* out_bbox) {
  const VoxelAppend app = append ? *append : VoxelAppend{nullptr, nullptr, nullptr};
  if (n_max > ws.n_max) n_max = ws.n_max;
  const char* in = (const char*)d_in;
  const int g = grid_for(n_max, ws.large_input ? 4 : 2);
  int gt = (n_max + kScanTile - 1) / kScanTile, gc = gt;
  if (gt > g_rank_grid_limit) gt = g_rank_grid_limit;         // the look-back kernels loop over tiles; all CTAs of a grid must be able to be
  if (gc > g_classify_grid_limit) gc = g_classify_grid_limit; // resident together (a waiting tile depends on lower-numbered ones)
  int* counts = ws.d_counts;   // [0] input points, [1] points kept by the crop, [2] MAIN, [3] SIDE
  FLOAM_LAUNCH(K_VOXEL_BBOX, voxel_bbox_kernel, g, kThreads, s, in, stride_bytes, d_n, d_extra, cap, d_crop, ws.bbox, counts, app, ws.merge_state, d_skip);
  FLOAM_LAUNCH(K_VOXEL_CLASSIFY, voxel_classify_kernel, gc, kScanThreads, s, in, stride_bytes, d_n, counts, leaf, d_crop, ws.bbox, ws.main_keys, ws.main_vals,
               ws.keys, ws.vals, counts + 2, ws.d_nbits, ws.d_passthrough, ws.merge_state, ws.scan.block_sums, d_skip);
  unsigned int* skeys = nullptr;
  int* svals = nullptr;
  radix_sort_pairs(ws.keys, ws.vals, counts + 3, ws.d_nbits, n_max, ws.sort, d_skip, s, &skeys, &svals);   // the SIDE list only
  const int gm = g < 2 * kNumSMs ? g : 2 * kNumSMs;
  FLOAM_LAUNCH_DYN(K_VOXEL_MERGE, voxel_merge_kernel, gm, kMergeThreads, 2 * kMaxSamples * sizeof(unsigned int), s, ws.main_keys, ws.main_vals, skeys, svals,
                   counts + 2, ws.keys, ws.vals, d_skip);   // the sort left its result in the alternate buffers: ws.keys / ws.vals are free again
  FLOAM_LAUNCH(K_VOXEL_RANK, voxel_rank_kernel, gt, kScanThreads, s, ws.keys, counts + 1, ws.scan.block_sums, ws.flags, d_nout, d_skip);
  const int gr = g < 3 * kNumSMs ? g : 3 * kNumSMs;
  FLOAM_LAUNCH(K_VOXEL_REDUCE, voxel_reduce_kernel, gr, kThreads, s, in, stride_bytes, ws.vals, ws.flags, d_nout, d_out, ws.bbox, counts, out_bbox, d_skip);
}

void repack_xyzi_device(const void* d_in32, const int* d_n, int n_max, P4* d_out, cudaStream_t s) {
  FLOAM_LAUNCH(K_REPACK, repack_kernel, grid_for(n_max), kThreads, s, (const char*)d_in32, d_n, d_out);
}

void crop_box_device(const P4* d_in, const int* d_n, int n_max, const float* d_bounds, P4* d_out, int* d_nout, VoxelWorkspace& ws,
                     const int* d_skip, cudaStream_t s, const int* d_extra, int cap) {
  if (n_max > ws.n_max) n_max = ws.n_max;
  const int gt = (n_max + kScanTile - 1) / kScanTile;
  FLOAM_LAUNCH(K_CROP_FLAGS, crop_flags_kernel, gt, kScanThreads, s, d_in, d_n, d_extra, cap, d_bounds, ws.scan.block_sums, d_skip);
  FLOAM_LAUNCH(K_CROP_SCATTER, crop_scatter_kernel, gt, kScanThreads, s, d_in, d_n, d_extra, cap, d_bounds, ws.scan.block_sums, d_out, d_nout, d_skip);
}

}  // namespace floam
