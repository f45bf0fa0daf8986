// Device-wide building blocks: stable LSD radix sort of (key,value) pairs and exclusive scan.
// Element counts live in device memory, so the callers (voxel filter, grid build) never synchronise with the host.
#include "common.cuh"

namespace floam {

thread_local long long g_launches = 0;
thread_local LaunchTimer* g_timer = nullptr;

static const char* const kSlotNames[K_NUM_SLOTS] = {
  "ring_count",
  "ring_scatter",
  "sector",
  "feature_offsets",
  "feature_gather",
  "deskew_align",
  "classify_old",
  "partition",
  "transform_new",
  "keys1",
  "keys2",
  "heads",
  "reduce",
  "commit",
  "cell_keys",
  "gather",
  "grid_bbox",
  "grid_dims",
  "grid_count",
  "grid_scatter",
  "state_init",
  "map_append_raw",
  "map_bump",
  "predict",
  "assoc_eval",
  "cand_eval",
  "finish",
  "map_append",
  "compensate_velocity",
  "knn5",
  "radix_hist",
  "single_block_scan",
  "radix_scatter",
  "scan_tiles",
  "scan_add",
  "voxel_init",
  "voxel_bbox",
  "voxel_keys",
  "voxel_heads",
  "voxel_reduce",
  "repack",
  "crop_flags",
  "crop_scatter",
  "record_pose"
};
const char* kernel_slot_name(int slot) { return (slot >= 0 && slot < K_NUM_SLOTS) ? kSlotNames[slot] : "?"; }

void launch_timer_begin(int slot, cudaStream_t s) {
  LaunchTimer* t = g_timer;
  if (!t || !t->enabled) return;
  if (t->used == LaunchTimer::kPairs) launch_timer_collect(t, s);
  t->slot_of[t->used] = slot;
  cudaEventRecord(t->ev[2 * t->used], s);
}
void launch_timer_end(cudaStream_t s) {
  LaunchTimer* t = g_timer;
  if (!t || !t->enabled) return;
  cudaEventRecord(t->ev[2 * t->used + 1], s);
  t->used++;
}
int launch_timer_collect(LaunchTimer* t, cudaStream_t s) {
  if (!t) return FLOAM_OK;
  FLOAM_CUDA_OK(cudaStreamSynchronize(s));
  for (int i = 0; i < t->used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t->ev[2 * i], t->ev[2 * i + 1]) == cudaSuccess) {
      t->total_ms[t->slot_of[i]] += ms;
      t->launches[t->slot_of[i]]++;
    }
  }
  t->used = 0;
  return FLOAM_OK;
}

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItemsPerWarp = 512;
constexpr int kSortTile = kSortWarps * kSortItemsPerWarp;  // 4096 keys per block

// Per-warp digit histogram of the warp's contiguous 512-key chunk. cnt points at this warp's 256 counters.
__device__ __forceinline__ void warp_digit_count(const unsigned int* __restrict__ keys, int begin, int n, int shift, int* cnt) {
  const int l = lane_id();
  for (int r = 0; r < kSortItemsPerWarp / 32; ++r) {
    const int i = begin + r * 32 + l;
    const bool valid = i < n;
    const unsigned int d = valid ? ((keys[i] >> shift) & 0xffu) : 0xffffffffu;
    const unsigned int m = __match_any_sync(0xffffffffu, d);
    if (valid && (__ffs(m) - 1) == l) cnt[d] += __popc(m);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const unsigned int* __restrict__ keys, const int* __restrict__ d_n,
                                                                  const int* __restrict__ d_nbits, int shift, int* __restrict__ hist, int nblocks, const int* d_skip) {
  if (d_skip && *d_skip) return;
  if (shift >= *d_nbits) return;
  const int n = *d_n;
  const int tile0 = blockIdx.x * kSortTile;
  if (tile0 >= n) return;
  nblocks = (n + kSortTile - 1) / kSortTile;  // histogram rows are laid out for the live element count, not the capacity
  __shared__ int s_cnt[kSortWarps][256];
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
  __syncthreads();
  warp_digit_count(keys, tile0 + warp_id() * kSortItemsPerWarp, n, shift, s_cnt[warp_id()]);
  __syncthreads();
  for (int d = threadIdx.x; d < 256; d += kSortThreads) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) t += s_cnt[w][d];
    hist[d * nblocks + blockIdx.x] = t;
  }
}

// in-place exclusive scan of a small array by one block
// d_sort_n (optional): the array is the radix histogram of *d_sort_n keys -> n = 256 * ceil(*d_sort_n / kSortTile)
__global__ void __launch_bounds__(1024) single_block_scan_kernel(int* __restrict__ data, int n, const int* __restrict__ d_nbits, int shift, const int* d_skip,
                                                                 const int* __restrict__ d_sort_n) {
  if (d_skip && *d_skip) return;
  if (d_nbits && shift >= *d_nbits) return;
  if (d_sort_n) n = 256 * ((*d_sort_n + kSortTile - 1) / kSortTile);
  __shared__ int smem[33];
  int carry = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < n) ? data[i] : 0;
    int total;
    const int ex = block_excl_scan(v, smem, &total);
    if (i < n) data[i] = ex + carry;
    carry += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const unsigned int* __restrict__ keys_in, const int* __restrict__ vals_in,
                                                                     unsigned int* __restrict__ keys_out, int* __restrict__ vals_out,
                                                                     const int* __restrict__ d_n, const int* __restrict__ d_nbits, int shift,
                                                                     const int* __restrict__ hist, int nblocks, const int* d_skip) {
  if (d_skip && *d_skip) return;
  const int n = *d_n;
  const int tile0 = blockIdx.x * kSortTile;
  if (tile0 >= n) return;
  nblocks = (n + kSortTile - 1) / kSortTile;
  if (shift >= *d_nbits) {  // pass not needed for this key width: keep ping-pong parity with a straight copy
    for (int i = tile0 + threadIdx.x; i < min(n, tile0 + kSortTile); i += kSortThreads) {
      keys_out[i] = keys_in[i];
      vals_out[i] = vals_in[i];
    }
    return;
  }
  __shared__ int s_cnt[kSortWarps][256];
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
  __syncthreads();
  const int w = warp_id(), l = lane_id();
  const int begin = tile0 + w * kSortItemsPerWarp;
  warp_digit_count(keys_in, begin, n, shift, s_cnt[w]);
  __syncthreads();
  // per digit: exclusive prefix over warps + global base of (digit, block)
  for (int d = threadIdx.x; d < 256; d += kSortThreads) {
    int run = hist[d * nblocks + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < kSortWarps; ++ww) {
      const int c = s_cnt[ww][d];
      s_cnt[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
  int* base = s_cnt[w];
  for (int r = 0; r < kSortItemsPerWarp / 32; ++r) {
    const int i = begin + r * 32 + l;
    const bool valid = i < n;
    unsigned int k = 0;
    int v = 0;
    if (valid) { k = keys_in[i]; v = vals_in[i]; }
    const unsigned int d = valid ? ((k >> shift) & 0xffu) : 0xffffffffu;
    const unsigned int m = __match_any_sync(0xffffffffu, d);
    int pos = 0;
    if (valid) pos = base[d] + __popc(m & ((1u << l) - 1u));
    __syncwarp();
    if (valid && (__ffs(m) - 1) == l) base[d] += __popc(m);
    __syncwarp();
    if (valid) { keys_out[pos] = k; vals_out[pos] = v; }
  }
}

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ d_n,
                                                                  int n_fixed, int* __restrict__ block_sums, const int* d_skip) {
  if (d_skip && *d_skip) return;
  const int n = d_n ? *d_n : n_fixed;
  __shared__ int smem[33];
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; sum += v[k]; }
  int total;
  int ex = block_excl_scan(sum, smem, &total);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(int* __restrict__ out, const int* __restrict__ d_n, int n_fixed,
                                                                const int* __restrict__ block_sums, int nblocks, const int* d_skip) {
  if (d_skip && *d_skip) return;
  const int n = d_n ? *d_n : n_fixed;
  const int off = block_sums[blockIdx.x];
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  if (blockIdx.x > 0) {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) if (base + k < n) out[base + k] += off;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = block_sums[nblocks];  // grand total
}

}  // namespace

size_t sort_workspace_bytes(int n_max) {
  const int nblocks = (n_max + kSortTile - 1) / kSortTile;
  return (size_t)n_max * 8 + (size_t)256 * nblocks * 4 + 256;
}
void sort_workspace_bind(SortWorkspace& ws, void* mem, int n_max) {
  char* p = (char*)mem;
  ws.n_max = n_max;
  ws.max_blocks = (n_max + kSortTile - 1) / kSortTile;
  ws.keys_alt = (unsigned int*)p; p += (size_t)n_max * 4;
  ws.vals_alt = (int*)p; p += (size_t)n_max * 4;
  ws.hist = (int*)p;
}

void radix_sort_pairs(unsigned int* keys, int* vals, const int* d_n, const int* d_nbits, int n_max, SortWorkspace& ws, const int* d_skip, cudaStream_t s) {
  if (n_max > ws.n_max) n_max = ws.n_max;
  const int nblocks = (n_max + kSortTile - 1) / kSortTile;
  unsigned int* kin = keys; int* vin = vals;
  unsigned int* kout = ws.keys_alt; int* vout = ws.vals_alt;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = pass * 8;
    FLOAM_LAUNCH(K_RADIX_HIST, radix_hist_kernel, nblocks, kSortThreads, s, kin, d_n, d_nbits, shift, ws.hist, nblocks, d_skip);
    FLOAM_LAUNCH(K_SINGLE_BLOCK_SCAN, single_block_scan_kernel, 1, 1024, s, ws.hist, 256 * nblocks, d_nbits, shift, d_skip, d_n);
    FLOAM_LAUNCH(K_RADIX_SCATTER, radix_scatter_kernel, nblocks, kSortThreads, s, kin, vin, kout, vout, d_n, d_nbits, shift, ws.hist, nblocks, d_skip);
    unsigned int* tk = kin; kin = kout; kout = tk;
    int* tv = vin; vin = vout; vout = tv;
  }
}

void exclusive_scan_small(int* data, int n, cudaStream_t s) {
  FLOAM_LAUNCH(K_SINGLE_BLOCK_SCAN, single_block_scan_kernel, 1, 1024, s, data, n, nullptr, 0, nullptr, nullptr);
}

size_t scan_workspace_bytes(int n_max) { return ((size_t)(n_max + kScanTile - 1) / kScanTile + 2) * 4 + 256; }
void scan_workspace_bind(ScanWorkspace& ws, void* mem, int n_max) { ws.block_sums = (int*)mem; ws.n_max = n_max; }

void exclusive_scan_i32(const int* in, int* out, const int* d_n, int n_fixed, int n_max, ScanWorkspace& ws, const int* d_skip, cudaStream_t s) {
  const int nblocks = (n_max + kScanTile - 1) / kScanTile;
  FLOAM_LAUNCH(K_SCAN_TILES, scan_tiles_kernel, nblocks, kScanThreads, s, in, out, d_n, n_fixed, ws.block_sums, d_skip);
  // exclusive scan of nblocks+1 entries: entry nblocks becomes the grand total
  FLOAM_LAUNCH(K_SINGLE_BLOCK_SCAN, single_block_scan_kernel, 1, 1024, s, ws.block_sums, nblocks + 1, nullptr, 0, d_skip, nullptr);
  FLOAM_LAUNCH(K_SCAN_ADD, scan_add_kernel, nblocks, kScanThreads, s, out, d_n, n_fixed, ws.block_sums, nblocks, d_skip);
}

}  // namespace floam
