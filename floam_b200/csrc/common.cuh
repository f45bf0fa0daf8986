// Shared device/host definitions for the B200-native FLOAM odometry path (sm_100a only).
// Data layout and kernel inventory are described in DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "floam_b200.h"

#define FLOAM_CUDA_OK(expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      std::fprintf(stderr, "[floam_b200] CUDA error %s at %s:%d: %s\n", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return FLOAM_ERR_CUDA;                                                                         \
    }                                                                                                \
  } while (0)

namespace floam {

constexpr int kNumSMs = 148;  // B200

// ---- launch accounting and per-kernel-class device timing (sort_scan.cu) ----
// Every kernel goes through FLOAM_LAUNCH: it counts the launch (bench.py's gpu_launches) and, when a LaunchTimer is active on
// the calling thread, brackets the kernel with CUDA events on its stream (bench.py's roofline leg; graphs are off then).
enum KernelSlot {
  K_RING_COUNT,
  K_RING_SCATTER,
  K_SECTOR,
  K_FEATURE_OFFSETS,
  K_FEATURE_GATHER,
  K_DESKEW_ALIGN,
  K_CLASSIFY_OLD,
  K_PARTITION,
  K_TRANSFORM_NEW,
  K_KEYS1,
  K_KEYS2,
  K_HEADS,
  K_REDUCE,
  K_COMMIT,
  K_CELL_KEYS,
  K_GATHER,
  K_GRID_BBOX,
  K_GRID_COUNT,
  K_GRID_SCATTER,
  K_STATE_INIT,
  K_MAP_APPEND_RAW,
  K_MAP_BUMP,
  K_PREDICT,
  K_ASSOC_KNN,
  K_ASSOC_EVAL,
  K_LM_CLUSTER,
  K_FINISH,
  K_COMPENSATE_VELOCITY,
  K_KNN5,
  K_RADIX_HIST,
  K_SINGLE_BLOCK_SCAN,
  K_RADIX_SCATTER,
  K_SCAN_TILES,
  K_SCAN_ADD,
  K_VOXEL_RANK,
  K_VOXEL_BBOX,
  K_VOXEL_KEYS,
  K_VOXEL_REDUCE,
  K_REPACK,
  K_CROP_FLAGS,
  K_CROP_SCATTER,
  K_RECORD_POSE,
  K_MAIL_STATE,
  K_UNPACK_PC2,
  K_VOXEL_CLASSIFY,
  K_VOXEL_MERGE,
  K_NOOP,
  K_NUM_SLOTS
};
const char* kernel_slot_name(int slot);
extern thread_local long long g_launches;

struct LaunchTimer {
  static constexpr int kPairs = 2048;
  bool enabled = false;
  cudaEvent_t ev[2 * kPairs];
  int slot_of[kPairs];
  int used = 0;      // pairs handed out so far
  int persist = 0;   // pairs [0, persist) belong to captured frame graphs (event-record nodes) and are re-read after every replay
  bool created = false;
  double total_ms[K_NUM_SLOTS];
  long long launches[K_NUM_SLOTS];
};
extern thread_local LaunchTimer* g_timer;
void launch_timer_begin(int slot, cudaStream_t s);
void launch_timer_end(cudaStream_t s);
int launch_timer_collect(LaunchTimer* t, cudaStream_t s);  // synchronises the stream and folds the pending (non-graph) event pairs into the totals
// empty kernel launched once per frame in timing mode: its measured duration is the cost of the event-pair bracketing itself
void launch_noop(cudaStream_t s);
void launch_timer_fold(LaunchTimer* t, int first, int last);  // folds the pairs [first, last) of a replayed graph (after a synchronise)

// Programmatic dependent launch: every kernel starts with pdl_prologue() — it lets the NEXT kernel of the stream be scheduled
// right away (griddepcontrol.launch_dependents) and then waits until the PREVIOUS kernel has completed and flushed
// (griddepcontrol.wait) before touching memory. With the launch attribute below, the launch latency of kernel k+1 overlaps the
// execution of kernel k; without it both instructions are no-ops. The frame is a chain of ~60 dependent, microsecond-sized
// kernels, so this is where the time goes.
extern bool g_use_pdl;         // FLOAM_PDL=1: every launch (measured slower: early-launched CTAs of wide kernels crowd the SMs)
extern bool g_pdl_solve;       // the serial solve chain only (prediction -> kNN -> fit -> LM, narrow kernels): FLOAM_PDL_SOLVE, default on
extern thread_local bool t_pdl_scope;   // set around the launches of that chain (PdlSolveScope)
struct PdlSolveScope {
  bool saved;
  PdlSolveScope() : saved(t_pdl_scope) { t_pdl_scope = g_pdl_solve; }
  ~PdlSolveScope() { t_pdl_scope = saved; }
};
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_kernel_dyn(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t dyn_smem, cudaStream_t s, Args&&... args) {
  if (!g_use_pdl && !t_pdl_scope) {
    kern<<<grid, block, dyn_smem, s>>>(KArgs(args)...);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = dyn_smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args&&... args) {
  launch_kernel_dyn(kern, grid, block, 0, s, static_cast<Args&&>(args)...);
}

// one thread-block cluster of `cluster` CTAs (runtime cluster dimension; 16 needs cudaFuncAttributeNonPortableClusterSizeAllowed)
template <typename... KArgs, typename... Args>
inline void launch_cluster_kernel(void (*kern)(KArgs...), int cluster, dim3 block, size_t dyn_smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster); cfg.blockDim = block; cfg.dynamicSmemBytes = dyn_smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  int n = 1;
  if (g_use_pdl || t_pdl_scope) {
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    n = 2;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#define FLOAM_LAUNCH_CLUSTER(slot, kern, cluster, block, smem, stream, ...) \
  do {                                                                     \
    ::floam::g_launches++;                                                 \
    if (::floam::g_timer) ::floam::launch_timer_begin(slot, stream);       \
    ::floam::launch_cluster_kernel(kern, cluster, dim3(block), smem, stream, __VA_ARGS__); \
    if (::floam::g_timer) ::floam::launch_timer_end(stream);               \
  } while (0)

// dynamic shared memory variant (the kernel's cudaFuncAttributeMaxDynamicSharedMemorySize is raised by its owner, once)
#define FLOAM_LAUNCH_DYN(slot, kern, grid, block, smem, stream, ...)       \
  do {                                                                     \
    ::floam::g_launches++;                                                 \
    if (::floam::g_timer) ::floam::launch_timer_begin(slot, stream);       \
    ::floam::launch_kernel_dyn(kern, dim3(grid), dim3(block), smem, stream, __VA_ARGS__); \
    if (::floam::g_timer) ::floam::launch_timer_end(stream);               \
  } while (0)

#define FLOAM_LAUNCH(slot, kern, grid, block, stream, ...)                 \
  do {                                                                     \
    ::floam::g_launches++;                                                 \
    if (::floam::g_timer) ::floam::launch_timer_begin(slot, stream);       \
    ::floam::launch_kernel(kern, dim3(grid), dim3(block), stream, __VA_ARGS__); \
    if (::floam::g_timer) ::floam::launch_timer_end(stream);               \
  } while (0)

// 32-byte scan/feature point, byte-identical to vel_point::PointXYZIRT (reference include/lidar.h:14-32).
struct __align__(16) PointIRT {
  float x, y, z, pad0;
  float intensity;
  unsigned short ring, pad1;
  float time;
  float pad2;
};
static_assert(sizeof(PointIRT) == 32, "PointXYZIRT layout");

// Odometry-side clouds are kept as float4 (x, y, z, intensity): half the bytes of pcl::PointXYZI, same information.
typedef float4 P4;

// 32-byte pcl::PointXYZI for the boundary
struct __align__(16) PointI {
  float x, y, z, pad0;
  float intensity, p1, p2, p3;
};
static_assert(sizeof(PointI) == 32, "PointXYZI layout");

// ---- arithmetic that must match the reference's non-contracted x86 code bit for bit (SURVEY.md §7 FP discipline) ----
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// order-preserving float <-> uint mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ unsigned int float_flip(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_unflip(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= o) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32). smem: 33 ints. Returns exclusive
// prefix; *total receives the block sum (valid in every thread).
__device__ __forceinline__ int block_excl_scan(int v, int* smem, int* total) {
  int incl = warp_incl_scan(v);
  const int w = warp_id(), l = lane_id(), nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 31) smem[w] = incl;
  __syncthreads();
  if (w == 0) {
    int s = (l < nw) ? smem[l] : 0;
    int si = warp_incl_scan(s);
    smem[l] = si - s;
    if (l == 31) smem[32] = si;
  }
  __syncthreads();
  int r = incl - v + smem[w];
  *total = smem[32];
  return r;
}

// block-wide sum of one int per thread (every thread gets the result). smem: 33 ints.
__device__ __forceinline__ int block_sum(int v, int* smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = warp_id(), l = lane_id(), nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    int s = (l < nw) ? smem[l] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (l == 0) smem[32] = s;
  }
  __syncthreads();
  return smem[32];
}

// Two-kernel scans: kernel 1 leaves one total per tile in tile_sums; in kernel 2 every CTA adds up the totals of the tiles before
// it by itself (a few thousand L2-resident ints at most) instead of waiting for a serial single-CTA scan in between.
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;
__device__ __forceinline__ int tile_offset(const int* __restrict__ tile_sums, int tile, int* smem) {
  int s = 0;
  for (int t = threadIdx.x; t < tile; t += blockDim.x) s += tile_sums[t];
  return block_sum(s, smem);
}

// ---- device-wide primitives (sort_scan.cu). All sizes are read from device memory so a frame needs no host sync. ----
struct SortWorkspace {
  unsigned int* keys_alt;   // capacity n_max
  int* vals_alt;            // capacity n_max
  int* hist;                // three digit-count tables (one per pass), table_stride ints apart
  int table_stride;
  unsigned int* ticket;     // last-CTA-done counter of the count kernel
  int max_blocks;
  int n_max;
};
size_t sort_workspace_bytes(int n_max);
void sort_workspace_bind(SortWorkspace& ws, void* mem, int n_max);
int sort_workspace_arm(SortWorkspace& ws, cudaStream_t s);
// Stable LSD radix sort of (key,value) pairs in three passes of ceil(*d_nbits / 3) bits each (decided on the device). n is read from
// *d_n; launch geometry is sized for n_max. The sorted arrays are returned through sorted_keys / sorted_vals (workspace buffers).
void radix_sort_pairs(unsigned int* keys, int* vals, const int* d_n, const int* d_nbits, int n_max, SortWorkspace& ws, const int* d_skip, cudaStream_t s,
                      unsigned int** sorted_keys, int** sorted_vals);
// same, with the ping-pong partner buffers given explicitly (the result lands in keys_alt / vals_alt)
void radix_sort_pairs_from(unsigned int* keys, int* vals, unsigned int* keys_alt, int* vals_alt, const int* d_n, const int* d_nbits, int n_max,
                           SortWorkspace& ws, const int* d_skip, cudaStream_t s, unsigned int** sorted_keys, int** sorted_vals);

// in-place exclusive scan of a small device array (n known on the host) by a single CTA
void exclusive_scan_small(int* data, int n, cudaStream_t s);

struct ScanWorkspace {
  int* block_sums;  // one total per 4096-element tile
  int n_max;
};
// second half of exclusive_scan_i32 alone, for a caller that already holds the sums of every kScanTile-element tile of `in`
void exclusive_scan_with_tile_sums(const int* in, int* out, const int* d_n, int n_max, const int* tile_sums, const int* d_skip, cudaStream_t s);
size_t scan_workspace_bytes(int n_max);
void scan_workspace_bind(ScanWorkspace& ws, void* mem, int n_max);
// out[i] = sum_{j<i} in[j] for i in [0, n]; out has n+1 entries (out[n] = total). n read from *d_n (or n_fixed if d_n==nullptr).
void exclusive_scan_i32(const int* in, int* out, const int* d_n, int n_fixed, int n_max, ScanWorkspace& ws, const int* d_skip, cudaStream_t s);

}  // namespace floam
