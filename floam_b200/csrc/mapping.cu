// LaserMappingClass on the device (reference src/laserMappingClass.cpp:7-32,106-200).
//
// The reference keeps a vector<vector<vector<cloud>>> of 50 m cells and, per frame, pushes the transformed scan into the cells and
// runs an in-place pcl::VoxelGrid over each of the 5x5x5 cells around the sensor.  Here the global map is ONE flat cloud with a
// packed cell id per point.  A frame
//   1. splits the map into "rest" and "block" (points whose cell lies in the 5x5x5 block) with a stable partition,
//   2. appends the float-transformed scan (z-based intensity, :165) to the block,
//   3. voxelises the block with two stable radix sorts, (ky,kx) then (cell,kz) — i.e. per cell in ascending VoxelGrid index,
//      accumulating in arrival order exactly like the 125 separate filters (old centroid first, then the new points),
//   4. writes rest ++ block back.
// getMap() (:188-200) orders cells x-major, y, z like the reference's triple loop with one more stable sort by cell id.
// Points that land in a cell allocated by an earlier block but outside the current one are appended unfiltered, like the
// reference's push_back; points in cells that were never allocated make the reference dereference a null cloud (undefined
// behaviour) — here they are dropped and counted in d_counts[4].
#include "mapping.cuh"

#include <cmath>
#include <vector>

namespace floam {
namespace {

constexpr int kThreads = 256;
constexpr double kCell = 50.0;     // LASER_CELL_WIDTH/HEIGHT/DEPTH, include/laserMappingClass.h:26-28
constexpr int kRange = 2;          // LASER_CELL_RANGE_HORIZONTAL/VERTICAL, :32-33
constexpr unsigned int kNoCell = 0xffffffffu;

struct BlockGeom {
  int cx, cy, cz;       // cell of the sensor position
  int kx0, ky0, kz0;    // voxel coordinate origin of the block
  int DX, DZ;           // key strides: key1 = kx + ky * DX ; key2 = kz + local_cell * DZ
  float inv_leaf;
  float R[9], t[3];     // pose.cast<float>()
};

inline int grid_for(int n_max) {
  int g = (n_max + kThreads - 1) / kThreads;
  const int cap = kNumSMs * 8;
  return g < 1 ? 1 : (g > cap ? cap : g);
}

__host__ __device__ inline unsigned int pack_cell(int cx, int cy, int cz) {
  return ((unsigned int)(cx + 512) << 20) | ((unsigned int)(cy + 512) << 10) | (unsigned int)(cz + 512);
}
constexpr int kLcPass = 125;   // allocated cell outside the current block: kept as is, not filtered this frame
constexpr int kLcDrop = 126;   // unallocated cell: the reference would index a null cloud here
__device__ __forceinline__ int local_cell(unsigned int packed, const BlockGeom& g) {  // 0..124 inside the block, 125 otherwise
  if (packed == kNoCell) return kLcDrop;
  const int dx = (int)(packed >> 20) - 512 - g.cx + kRange, dy = (int)((packed >> 10) & 1023u) - 512 - g.cy + kRange,
            dz = (int)(packed & 1023u) - 512 - g.cz + kRange;
  if (dx < 0 || dx > 2 * kRange || dy < 0 || dy > 2 * kRange || dz < 0 || dz > 2 * kRange) return 125;
  return (dx * 5 + dy) * 5 + dz;
}
__device__ __forceinline__ int cell_coord(float v) { return (int)floor((double)v / kCell + 0.5); }  // :166-168

__global__ void __launch_bounds__(kThreads) classify_old_kernel(const unsigned int* __restrict__ cell, const int* __restrict__ counts, BlockGeom g,
                                                                 int* __restrict__ flags) {
  pdl_prologue();
  const int n = counts[0];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) flags[i] = local_cell(cell[i], g) < kLcPass ? 1 : 0;
}

__global__ void __launch_bounds__(kThreads) partition_kernel(const P4* __restrict__ pts, const unsigned int* __restrict__ cell, const int* __restrict__ pos,
                                                              int* __restrict__ counts, P4* __restrict__ rest, unsigned int* __restrict__ rest_cell,
                                                              P4* __restrict__ work, unsigned int* __restrict__ work_cell) {
  pdl_prologue();
  const int n = counts[0];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const int p = pos[i];
    const bool in_block = pos[i + 1] != p;
    if (in_block) { work[p] = pts[i]; work_cell[p] = cell[i]; }
    else { rest[i - p] = pts[i]; rest_cell[i - p] = cell[i]; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int nin = n > 0 ? pos[n] : 0;
    counts[2] = nin;
    counts[1] = n - nin;
  }
}

// pcl::transformPointCloud(pose.cast<float>()) + intensity rewrite + cell id (:158-171)
__global__ void __launch_bounds__(kThreads) transform_new_kernel(const char* __restrict__ in, int stride, const int* __restrict__ d_nin, int* __restrict__ counts,
                                                                  BlockGeom g, int cap, P4* __restrict__ work, unsigned int* __restrict__ work_cell,
                                                                  const unsigned int* __restrict__ allocated, int n_allocated,
                                                                  unsigned char* __restrict__ dirty) {
  pdl_prologue();
  const int nin = *d_nin, base = counts[2];
  const bool fits = counts[1] + base + nin <= cap;
  if (fits) {
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < nin; i += gridDim.x * kThreads) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(in + (size_t)i * stride));
      const float x = fadd(fadd(fadd(fmul(g.R[0], p.x), fmul(g.R[1], p.y)), fmul(g.R[2], p.z)), g.t[0]);
      const float y = fadd(fadd(fadd(fmul(g.R[3], p.x), fmul(g.R[4], p.y)), fmul(g.R[5], p.z)), g.t[1]);
      const float z = fadd(fadd(fadd(fmul(g.R[6], p.x), fmul(g.R[7], p.y)), fmul(g.R[8], p.z)), g.t[2]);
      const float inten = (float)fmin(1.0, fmax((double)p.z + 2.0, 0.0) / 5);
      const int cx = cell_coord(x), cy = cell_coord(y), cz = cell_coord(z);
      unsigned int pc = kNoCell;
      if (abs(cx - g.cx) <= kRange && abs(cy - g.cy) <= kRange && abs(cz - g.cz) <= kRange) {
        pc = pack_cell(cx, cy, cz);
      } else if (abs(cx) < 512 && abs(cy) < 512 && abs(cz) < 512) {
        const unsigned int want = pack_cell(cx, cy, cz);
        int lo = 0, hi = n_allocated;   // sorted list of every cell some earlier block allocated
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (allocated[mid] < want) lo = mid + 1; else hi = mid; }
        if (lo < n_allocated && allocated[lo] == want) { pc = want; dirty[lo] = 1; }   // appended to a cell outside the block: that cell changed too
      }
      if (pc == kNoCell) atomicAdd(&counts[4], 1);
      work[base + i] = make_float4(x, y, z, inten);
      work_cell[base + i] = pc;
    }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts[3] = fits ? base + nin : base;
    if (!fits) counts[7] = 1;
  }
}

__device__ __forceinline__ void voxel_of(const float4 p, const BlockGeom& g, int& kx, int& ky, int& kz) {
  kx = (int)floorf(fmul(p.x, g.inv_leaf)) - g.kx0;
  ky = (int)floorf(fmul(p.y, g.inv_leaf)) - g.ky0;
  kz = (int)floorf(fmul(p.z, g.inv_leaf)) - g.kz0;
}

__global__ void __launch_bounds__(kThreads) keys1_kernel(const P4* __restrict__ work, const unsigned int* __restrict__ work_cell, const int* __restrict__ counts,
                                                          BlockGeom g, unsigned int* __restrict__ keys, int* __restrict__ vals) {
  pdl_prologue();
  const int n = counts[3];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    unsigned int key = 0;
    if (local_cell(work_cell[i], g) < kLcPass) {
      int kx, ky, kz;
      voxel_of(work[i], g, kx, ky, kz);
      key = (unsigned int)(kx + ky * g.DX);
    }
    keys[i] = key;
    vals[i] = i;
  }
}

__global__ void __launch_bounds__(kThreads) keys2_kernel(const P4* __restrict__ work, const unsigned int* __restrict__ work_cell, const int* __restrict__ counts,
                                                          BlockGeom g, const int* __restrict__ vals, unsigned int* __restrict__ keys) {
  pdl_prologue();
  const int n = counts[3];
  for (int j = blockIdx.x * kThreads + threadIdx.x; j < n; j += gridDim.x * kThreads) {
    const int i = vals[j];
    const int lc = local_cell(work_cell[i], g);
    int kz = 0;
    if (lc < kLcPass) { int kx, ky; voxel_of(work[i], g, kx, ky, kz); }
    keys[j] = (unsigned int)(kz + lc * g.DZ);
  }
}

__global__ void __launch_bounds__(kThreads) heads_kernel(const P4* __restrict__ work, const unsigned int* __restrict__ work_cell, const int* __restrict__ counts,
                                                          BlockGeom g, const int* __restrict__ vals, int* __restrict__ flags) {
  pdl_prologue();
  const int n = counts[3];
  for (int j = blockIdx.x * kThreads + threadIdx.x; j < n; j += gridDim.x * kThreads) {
    const int i = vals[j];
    const unsigned int c = work_cell[i];
    int head = 0;
    if (c != kNoCell) {
      head = 1;
      if (j > 0 && local_cell(c, g) < kLcPass) {   // pass-through points are never merged
        const int ip = vals[j - 1];
        if (work_cell[ip] == c) {
          int ax, ay, az, bx, by, bz;
          voxel_of(work[i], g, ax, ay, az);
          voxel_of(work[ip], g, bx, by, bz);
          if (ax == bx && ay == by && az == bz) head = 0;
        }
      }
    }
    flags[j] = head;
  }
}

__global__ void __launch_bounds__(kThreads) reduce_kernel(const P4* __restrict__ work, const unsigned int* __restrict__ work_cell, int* __restrict__ counts,
                                                           BlockGeom g, const int* __restrict__ vals, const int* __restrict__ seg, P4* __restrict__ out,
                                                           unsigned int* __restrict__ out_cell) {
  pdl_prologue();
  const int n = counts[3], base = counts[1];
  for (int j = blockIdx.x * kThreads + threadIdx.x; j < n; j += gridDim.x * kThreads) {
    if (seg[j + 1] == seg[j]) continue;  // not a voxel head (or a dropped point)
    const unsigned int c = work_cell[vals[j]];
    int hx, hy, hz;
    voxel_of(work[vals[j]], g, hx, hy, hz);
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    // pcl::CentroidPoint: float sums in run order, then / n.  The ORDER of the additions is fixed, the loads are not: eight elements
    // of the run are fetched at a time (index and head flag, then point and cell), so a long run (a ground voxel next to the sensor
    // holds hundreds of returns) costs one round trip per eight points instead of three per point.
    int k = j;
    bool more = true;
    while (more) {
      int idx[8], sg[9];
#pragma unroll
      for (int u = 0; u < 8; ++u) idx[u] = (k + u < n) ? __ldg(vals + k + u) : -1;
#pragma unroll
      for (int u = 0; u < 9; ++u) sg[u] = (k + u <= n) ? __ldg(seg + k + u) : 0;
      float4 p[8];
      unsigned int cc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (idx[u] >= 0) { p[u] = work[idx[u]]; cc[u] = work_cell[idx[u]]; }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (!more) break;
        const bool first = (k + u == j);   // the head itself always belongs to the run
        if (!first && !(idx[u] >= 0 && sg[u + 1] == sg[u] && cc[u] == c)) { more = false; k += u; break; }
        sx = fadd(sx, p[u].x); sy = fadd(sy, p[u].y); sz = fadd(sz, p[u].z); si = fadd(si, p[u].w);
      }
      if (more) k += 8;
    }
    const float cnt = (float)(k - j);
    out[base + seg[j]] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), __fdiv_rn(si, cnt));
    out_cell[base + seg[j]] = c;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts[5] = n > 0 ? seg[n] : 0;
    counts[6] = base + (n > 0 ? seg[n] : 0);  // new map size, committed by commit_kernel
  }
}

__global__ void commit_kernel(int* counts) {
  pdl_prologue();
  if (threadIdx.x == 0) counts[0] = counts[6];
}

__global__ void __launch_bounds__(kThreads) cell_keys_kernel(const unsigned int* __restrict__ cell, const int* __restrict__ counts, unsigned int* __restrict__ keys,
                                                              int* __restrict__ vals) {
  pdl_prologue();
  const int n = counts[0];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) { keys[i] = cell[i]; vals[i] = i; }
}
__global__ void __launch_bounds__(kThreads) gather_kernel(const P4* __restrict__ pts, const int* __restrict__ vals, const int* __restrict__ counts,
                                                           P4* __restrict__ out) {
  pdl_prologue();
  const int n = counts[0];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) out[i] = pts[vals[i]];
}

// incremental getMap(): flag = 1 for every point whose cell carries a change mark
__global__ void __launch_bounds__(kThreads) dirty_flags_kernel(const unsigned int* __restrict__ cell, const int* __restrict__ counts,
                                                                const unsigned int* __restrict__ allocated, int n_allocated,
                                                                const unsigned char* __restrict__ dirty, int* __restrict__ flags) {
  pdl_prologue();
  const int n = counts[0];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const unsigned int want = cell[i];
    int lo = 0, hi = n_allocated;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(allocated + mid) < want) lo = mid + 1; else hi = mid; }
    flags[i] = (lo < n_allocated && __ldg(allocated + lo) == want && dirty[lo]) ? 1 : 0;
  }
}
__global__ void __launch_bounds__(kThreads) dirty_scatter_kernel(const P4* __restrict__ pts, const unsigned int* __restrict__ cell, int* __restrict__ counts,
                                                                  const int* __restrict__ pos, P4* __restrict__ out, unsigned int* __restrict__ out_cell) {
  pdl_prologue();
  const int n = counts[0];
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const int p = pos[i];
    if (pos[i + 1] != p) { out[p] = pts[i]; out_cell[p] = cell[i]; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[11] = n > 0 ? pos[n] : 0;
}
__global__ void __launch_bounds__(kThreads) mark_dirty_kernel(const unsigned int* __restrict__ cells, int n_cells, const unsigned int* __restrict__ allocated,
                                                               int n_allocated, unsigned char* __restrict__ dirty) {
  pdl_prologue();
  for (int k = blockIdx.x * kThreads + threadIdx.x; k < n_cells; k += gridDim.x * kThreads) {
    const unsigned int want = cells[k];
    int lo = 0, hi = n_allocated;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (allocated[mid] < want) lo = mid + 1; else hi = mid; }
    if (lo < n_allocated && allocated[lo] == want) dirty[lo] = 1;
  }
}

int bits_for(long long v) {
  int b = 1;
  while (b < 32 && (1ll << b) < v) ++b;
  return b;
}

}  // namespace

int mapping_device_init(MappingDevice& md, int cap, double map_resolution, VoxelWorkspace* vws, void* (*alloc)(void*, size_t), void* actx, cudaStream_t s) {
  md.cap = cap;
  md.leaf = (float)map_resolution;  // downSizeFilter.setLeafSize(map_resolution, ...) :31
  md.vws = vws;
  if (cap > vws->n_max) md.cap = vws->n_max;  // the sorts run in the shared voxel workspace
  md.pts = (P4*)alloc(actx, (size_t)md.cap * 16);
  md.pts_alt = (P4*)alloc(actx, (size_t)md.cap * 16);
  md.work = (P4*)alloc(actx, (size_t)md.cap * 16);
  md.cell = (unsigned int*)alloc(actx, (size_t)md.cap * 4);
  md.cell_alt = (unsigned int*)alloc(actx, (size_t)md.cap * 4);
  md.work_cell = (unsigned int*)alloc(actx, (size_t)md.cap * 4);
  md.d_counts = (int*)alloc(actx, 64);
  md.alloc_cap = 1 << 16;
  md.d_allocated = (unsigned int*)alloc(actx, (size_t)md.alloc_cap * 4);
  md.d_dirty = (unsigned char*)alloc(actx, (size_t)md.alloc_cap);
  if (md.d_dirty) FLOAM_CUDA_OK(cudaMemsetAsync(md.d_dirty, 0, (size_t)md.alloc_cap, s));
  md.d_nbits = md.d_counts ? md.d_counts + 8 : nullptr;
  if (!md.pts || !md.pts_alt || !md.work || !md.cell || !md.cell_alt || !md.work_cell || !md.d_counts || !md.d_allocated) return FLOAM_ERR_CUDA;
  FLOAM_CUDA_OK(cudaMemsetAsync(md.d_counts, 0, 64, s));
  for (int dx = -kRange; dx <= kRange; ++dx)   // LaserMappingClass::init allocates the block around the origin (:12-29)
    for (int dy = -kRange; dy <= kRange; ++dy)
      for (int dz = -kRange; dz <= kRange; ++dz) md.allocated.insert(pack_cell(dx, dy, dz));
  md.allocated_prev.assign(md.allocated.begin(), md.allocated.end());   // the device list is valid from the start (sorted: std::set order)
  FLOAM_CUDA_OK(cudaMemcpyAsync(md.d_allocated, md.allocated_prev.data(), md.allocated_prev.size() * 4, cudaMemcpyHostToDevice, s));
  FLOAM_CUDA_OK(cudaStreamSynchronize(s));
  md.enabled = true;
  return FLOAM_OK;
}

int mapping_update_device(MappingDevice& md, const void* d_in, int stride, const int* d_n, int n_max, const double T[16], cudaStream_t s) {
  BlockGeom g;
  g.cx = (int)std::floor(T[3] / kCell + 0.5);   // :150-152
  g.cy = (int)std::floor(T[7] / kCell + 0.5);
  g.cz = (int)std::floor(T[11] / kCell + 0.5);
  if (std::abs(g.cx) > 500 || std::abs(g.cy) > 500 || std::abs(g.cz) > 500) return FLOAM_ERR_CAPACITY;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) g.R[r * 3 + c] = (float)T[r * 4 + c];
    g.t[r] = (float)T[r * 4 + 3];
  }
  g.inv_leaf = 1.0f / md.leaf;
  // the block spans cells [c-2, c+2] -> coordinates [(c-2.5)*50, (c+2.5)*50); one metre of slack on each side
  const double span = (2 * kRange + 1) * kCell + 2.0;
  g.kx0 = (int)std::floor(((g.cx - kRange - 0.5) * kCell - 1.0) * g.inv_leaf) - 1;
  g.ky0 = (int)std::floor(((g.cy - kRange - 0.5) * kCell - 1.0) * g.inv_leaf) - 1;
  g.kz0 = (int)std::floor(((g.cz - kRange - 0.5) * kCell - 1.0) * g.inv_leaf) - 1;
  const long long D = (long long)std::ceil(span * g.inv_leaf) + 4;
  if (D * D >= (1ll << 32) || D * 128 >= (1ll << 32)) return FLOAM_ERR_ARG;  // leaf too small for 32-bit keys
  g.DX = (int)D;
  g.DZ = (int)D;
  const int nbits[2] = {bits_for(D * D), bits_for(D * 128)};
  FLOAM_CUDA_OK(cudaMemcpyAsync(md.d_nbits, nbits, 8, cudaMemcpyHostToDevice, s));

  // checkPoints (:106-145): the 5x5x5 block around the sensor is allocated; remember every cell that ever was
  const size_t before = md.allocated.size();
  for (int dx = -kRange; dx <= kRange; ++dx)
    for (int dy = -kRange; dy <= kRange; ++dy)
      for (int dz = -kRange; dz <= kRange; ++dz) md.allocated.insert(pack_cell(g.cx + dx, g.cy + dy, g.cz + dz));
  if (md.allocated.size() != before) {
    if ((int)md.allocated.size() > md.alloc_cap) return FLOAM_ERR_CAPACITY;
    // the sorted cell list is about to change, and with it the index of every cell in the dirty array: bring the device-side marks
    // (cells that received appended points) back into the host set first; the next hand-out re-applies them
    if (!md.allocated_prev.empty()) {
      std::vector<unsigned char> marks(md.allocated_prev.size());
      FLOAM_CUDA_OK(cudaMemcpyAsync(marks.data(), md.d_dirty, marks.size(), cudaMemcpyDeviceToHost, s));
      FLOAM_CUDA_OK(cudaStreamSynchronize(s));
      for (size_t i = 0; i < marks.size(); ++i) if (marks[i]) md.dirty_host.insert(md.allocated_prev[i]);
      FLOAM_CUDA_OK(cudaMemsetAsync(md.d_dirty, 0, (size_t)md.alloc_cap, s));
    }
    md.allocated_prev.assign(md.allocated.begin(), md.allocated.end());
    std::vector<unsigned int> sorted(md.allocated.begin(), md.allocated.end());
    FLOAM_CUDA_OK(cudaMemcpyAsync(md.d_allocated, sorted.data(), sorted.size() * 4, cudaMemcpyHostToDevice, s));
    FLOAM_CUDA_OK(cudaStreamSynchronize(s));
  }
  for (int dx = -kRange; dx <= kRange; ++dx)   // every cell of the block is re-filtered in place (:175-184): all of them count as changed
    for (int dy = -kRange; dy <= kRange; ++dy)
      for (int dz = -kRange; dz <= kRange; ++dz) md.dirty_host.insert(pack_cell(g.cx + dx, g.cy + dy, g.cz + dz));
  VoxelWorkspace& ws = *md.vws;
  int* counts = md.d_counts;
  const int gmap = grid_for(md.cap), gin = grid_for(n_max);
  FLOAM_LAUNCH(K_CLASSIFY_OLD, classify_old_kernel, gmap, kThreads, s, md.cell, counts, g, ws.flags);
  exclusive_scan_i32(ws.flags, ws.flags, counts, 0, md.cap, ws.scan, nullptr, s);
  FLOAM_LAUNCH(K_PARTITION, partition_kernel, gmap, kThreads, s, md.pts, md.cell, ws.flags, counts, md.pts_alt, md.cell_alt, md.work, md.work_cell);
  FLOAM_LAUNCH(K_TRANSFORM_NEW, transform_new_kernel, gin, kThreads, s, (const char*)d_in, stride, d_n, counts, g, md.cap, md.work, md.work_cell, md.d_allocated, (int)md.allocated.size(),
               md.d_dirty);
  FLOAM_LAUNCH(K_KEYS1, keys1_kernel, gmap, kThreads, s, md.work, md.work_cell, counts, g, ws.keys, ws.vals);
  unsigned int* k1 = nullptr; int* v1 = nullptr;
  radix_sort_pairs(ws.keys, ws.vals, counts + 3, md.d_nbits, md.cap, ws.sort, nullptr, s, &k1, &v1);
  // second-level keys are written next to the first-level order (k1/v1 live in the workspace's alternate buffers; the sort below
  // ping-pongs between them and the primary buffers again)
  FLOAM_LAUNCH(K_KEYS2, keys2_kernel, gmap, kThreads, s, md.work, md.work_cell, counts, g, v1, k1);
  unsigned int* k2 = nullptr; int* v2 = nullptr;
  radix_sort_pairs_from(k1, v1, ws.keys, ws.vals, counts + 3, md.d_nbits + 1, md.cap, ws.sort, nullptr, s, &k2, &v2);
  FLOAM_LAUNCH(K_HEADS, heads_kernel, gmap, kThreads, s, md.work, md.work_cell, counts, g, v2, ws.flags);
  exclusive_scan_i32(ws.flags, ws.flags, counts + 3, 0, md.cap, ws.scan, nullptr, s);
  FLOAM_LAUNCH(K_REDUCE, reduce_kernel, gmap, kThreads, s, md.work, md.work_cell, counts, g, v2, ws.flags, md.pts_alt, md.cell_alt);
  FLOAM_LAUNCH(K_COMMIT, commit_kernel, 1, 32, s, counts);
  std::swap(md.pts, md.pts_alt);
  std::swap(md.cell, md.cell_alt);
  return FLOAM_OK;
}

int mapping_get_map_device(MappingDevice& md, P4** d_out, int** d_out_n, cudaStream_t s) {
  VoxelWorkspace& ws = *md.vws;
  const int nbits = 30;
  FLOAM_CUDA_OK(cudaMemcpyAsync(md.d_nbits + 2, &nbits, 4, cudaMemcpyHostToDevice, s));
  const int g = grid_for(md.cap);
  FLOAM_LAUNCH(K_CELL_KEYS, cell_keys_kernel, g, kThreads, s, md.cell, md.d_counts, ws.keys, ws.vals);
  unsigned int* sk = nullptr; int* sv = nullptr;
  radix_sort_pairs(ws.keys, ws.vals, md.d_counts, md.d_nbits + 2, md.cap, ws.sort, nullptr, s, &sk, &sv);
  FLOAM_LAUNCH(K_GATHER, gather_kernel, g, kThreads, s, md.pts, sv, md.d_counts, md.pts_alt);
  *d_out = md.pts_alt;
  *d_out_n = md.d_counts;
  return FLOAM_OK;
}

int mapping_get_dirty_device(MappingDevice& md, P4** d_out, unsigned int** d_out_cell, int** d_out_n, cudaStream_t s) {
  VoxelWorkspace& ws = *md.vws;
  const int g = grid_for(md.cap);
  const int n_alloc = (int)md.allocated.size();
  if (!md.dirty_host.empty()) {   // the block cells of the updates since the last call (the host knows them without asking the device)
    std::vector<unsigned int> cells(md.dirty_host.begin(), md.dirty_host.end());
    if (cells.size() > (size_t)md.alloc_cap) return FLOAM_ERR_CAPACITY;
    unsigned int* d_cells = reinterpret_cast<unsigned int*>(md.work_cell);   // idle between updates
    FLOAM_CUDA_OK(cudaMemcpyAsync(d_cells, cells.data(), cells.size() * 4, cudaMemcpyHostToDevice, s));
    FLOAM_LAUNCH(K_CLASSIFY_OLD, mark_dirty_kernel, grid_for((int)cells.size()), kThreads, s, d_cells, (int)cells.size(), md.d_allocated, n_alloc, md.d_dirty);
    FLOAM_CUDA_OK(cudaStreamSynchronize(s));   // `cells` is a pageable temporary
    md.dirty_host.clear();
  }
  FLOAM_LAUNCH(K_CLASSIFY_OLD, dirty_flags_kernel, g, kThreads, s, md.cell, md.d_counts, md.d_allocated, n_alloc, md.d_dirty, ws.flags);
  exclusive_scan_i32(ws.flags, ws.flags, md.d_counts, 0, md.cap, ws.scan, nullptr, s);
  FLOAM_LAUNCH(K_PARTITION, dirty_scatter_kernel, g, kThreads, s, md.pts, md.cell, md.d_counts, ws.flags, md.pts_alt, md.cell_alt);
  *d_out = md.pts_alt;
  *d_out_cell = md.cell_alt;
  *d_out_n = md.d_counts + 11;
  return FLOAM_OK;
}
int mapping_clear_dirty_device(MappingDevice& md, cudaStream_t s) {
  FLOAM_CUDA_OK(cudaMemsetAsync(md.d_dirty, 0, (size_t)md.alloc_cap, s));
  return FLOAM_OK;
}

}  // namespace floam
