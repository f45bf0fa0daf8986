// Subsystem (2): pcl::VoxelGrid and pcl::CropBox on the device.
// VoxelGrid (reference call sites src/odomEstimationClass.cpp:137-142,289-292, src/laserMappingClass.cpp:175-184; PCL 1.8.1
// applyFilter restated in SURVEY.md Appendix A.1): bbox reduction -> voxel key per point (same float floor/int arithmetic)
// -> stable radix sort of (key, index) -> segment heads -> exclusive scan -> one thread per voxel accumulates xyz and
// intensity in float, in ascending point index, and divides by the count.  CropBox (src/odomEstimationClass.cpp:270-287,
// Appendix A.2): predicate -> scan -> order-preserving scatter.
#include "voxel.cuh"

#include "odom_math.cuh"

namespace floam {
namespace {

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
                                                               const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  const int n = total_count(d_n, d_extra, cap);
  const int n_old = *d_n;
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
  // one wave of CTAs (the grid is sized for the hardware, not the capacity), each taking tiles b, b + gridDim.x, ...: a tile only
  // ever waits for lower-numbered tiles, which belong to CTAs that are resident or done
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int base = tile * kScanTile + threadIdx.x * kScanItems;
    int h[kScanItems];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { h[k] = (base + k < n) ? is_head(keys, base + k) : 0; sum += h[k]; }
    int total;
    int rank = block_excl_scan(sum, smem, &total);
    if (threadIdx.x == 0) {
      volatile int* st = tile_state;
      int before = 0;
      if (tile > 0) {
        st[tile] = (total << 2) | 1;
        for (int t = tile - 1; t >= 0; --t) {
          int v;
          while (((v = st[t]) & 3) == 0) __nanosleep(20);
          before += v >> 2;
          if ((v & 3) == 2) break;
        }
      }
      __threadfence();
      st[tile] = ((before + total) << 2) | 2;
      s_offset = before;
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

inline int grid_for(int n_max) {
  int g = (n_max + kThreads - 1) / kThreads;
  const int cap = kNumSMs * 4;   // the kernels stride; empty CTAs of a capacity-sized grid are not free
  return g < 1 ? 1 : (g > cap ? cap : g);
}

}  // namespace

size_t voxel_workspace_bytes(int n_max) {
  return (size_t)n_max * 4 * 2 + ((size_t)n_max + 1) * 4 + 1024 + sort_workspace_bytes(n_max) + scan_workspace_bytes(n_max + 1) + 4096;
}

void voxel_workspace_bind(VoxelWorkspace& ws, void* mem, int n_max) {
  char* p = (char*)mem;
  auto take = [&](size_t bytes) { void* r = p; p += (bytes + 255) / 256 * 256; return r; };
  ws.n_max = n_max;
  ws.keys = (unsigned int*)take((size_t)n_max * 4);
  ws.vals = (int*)take((size_t)n_max * 4);
  ws.flags = (int*)take(((size_t)n_max + 1) * 4);
  ws.bbox = (unsigned int*)take(32);
  ws.d_nbits = (int*)take(4);
  ws.d_passthrough = (int*)take(4);
  ws.d_counts = (int*)take(16);
  sort_workspace_bind(ws.sort, take(sort_workspace_bytes(n_max)), n_max);
  scan_workspace_bind(ws.scan, take(scan_workspace_bytes(n_max + 1)), n_max + 1);
}

int voxel_workspace_arm(VoxelWorkspace& ws, cudaStream_t s) {
  const unsigned int bb[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  FLOAM_CUDA_OK(cudaMemcpyAsync(ws.bbox, bb, sizeof(bb), cudaMemcpyHostToDevice, s));
  if (sort_workspace_arm(ws.sort, s)) return FLOAM_ERR_CUDA;
  FLOAM_CUDA_OK(cudaMemsetAsync(ws.d_counts, 0, 16, s));
  FLOAM_CUDA_OK(cudaStreamSynchronize(s));
  return FLOAM_OK;
}

void voxel_grid_device(const void* d_in, int stride_bytes, const int* d_n, int n_max, float leaf, P4* d_out, int* d_nout, VoxelWorkspace& ws,
                       const int* d_skip, cudaStream_t s, const float* d_crop, const int* d_extra, int cap, const VoxelAppend* append,
                       unsigned int* out_bbox) {
  const VoxelAppend app = append ? *append : VoxelAppend{nullptr, nullptr, nullptr};
  if (n_max > ws.n_max) n_max = ws.n_max;
  const char* in = (const char*)d_in;
  const int g = grid_for(n_max);
  int gt = (n_max + kScanTile - 1) / kScanTile;
  if (gt > 2 * kNumSMs) gt = 2 * kNumSMs;   // voxel_rank_kernel loops over tiles
  int* counts = ws.d_counts;   // [0] input points, [1] points kept by the crop
  FLOAM_LAUNCH(K_VOXEL_BBOX, voxel_bbox_kernel, g, kThreads, s, in, stride_bytes, d_n, d_extra, cap, d_crop, ws.bbox, counts, app, d_skip);
  FLOAM_LAUNCH(K_VOXEL_KEYS, voxel_keys_kernel, g, kThreads, s, in, stride_bytes, counts, leaf, d_crop, ws.bbox, ws.keys, ws.vals, ws.d_nbits, ws.d_passthrough,
               ws.scan.block_sums, d_skip);
  unsigned int* skeys = nullptr;
  int* svals = nullptr;
  radix_sort_pairs(ws.keys, ws.vals, counts, ws.d_nbits, n_max, ws.sort, d_skip, s, &skeys, &svals);
  FLOAM_LAUNCH(K_VOXEL_RANK, voxel_rank_kernel, gt, kScanThreads, s, skeys, counts + 1, ws.scan.block_sums, ws.flags, d_nout, d_skip);
  const int gr = g < 3 * kNumSMs ? g : 3 * kNumSMs;   // 80 registers + 33 KB shared memory: three CTAs per SM make one wave; the kernel strides
  FLOAM_LAUNCH(K_VOXEL_REDUCE, voxel_reduce_kernel, gr, kThreads, s, in, stride_bytes, svals, ws.flags, d_nout, d_out, ws.bbox, counts, out_bbox, d_skip);
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
