// Device-wide building blocks: stable LSD radix sort of (key,value) pairs and exclusive scan.
// Element counts live in device memory, so the callers (voxel filter, grid build) never synchronise with the host.
// Everything here is sized for latency on SMALL inputs (5k - 200k elements per frame, all L2-resident): tiles of 1024 keys so
// that even a 20k-element sort spreads over 20+ CTAs, every global load issued before its first use, and — up to 262,144 keys —
// no separate scan kernel: each scatter CTA derives its digit bases straight from the [tile][digit] count table.
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
  "assoc_knn",
  "assoc_eval",
  "lm_cluster",
  "finish",
  "map_append",
  "compensate_velocity",
  "knn5",
  "radix_hist",
  "single_block_scan",
  "radix_scatter",
  "scan_tiles",
  "scan_add",
  "voxel_rank",
  "voxel_bbox",
  "voxel_keys",
  "voxel_heads",
  "voxel_reduce",
  "repack",
  "crop_flags",
  "crop_scatter",
  "record_pose",
  "noop"
};
const char* kernel_slot_name(int slot) { return (slot >= 0 && slot < K_NUM_SLOTS) ? kSlotNames[slot] : "?"; }

__global__ void noop_kernel() {}
void launch_noop(cudaStream_t s) { FLOAM_LAUNCH(K_NOOP, noop_kernel, 1, 32, s); }

static void timer_record(cudaEvent_t ev, cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &st);
  // inside a capture the record becomes an event-record node of the graph; "external" lets the host read it after the replay
  if (st == cudaStreamCaptureStatusActive) cudaEventRecordWithFlags(ev, s, cudaEventRecordExternal);
  else cudaEventRecord(ev, s);
}
void launch_timer_begin(int slot, cudaStream_t s) {
  LaunchTimer* t = g_timer;
  if (!t || !t->enabled || t->used >= LaunchTimer::kPairs) return;
  t->slot_of[t->used] = slot;
  timer_record(t->ev[2 * t->used], s);
}
void launch_timer_end(cudaStream_t s) {
  LaunchTimer* t = g_timer;
  if (!t || !t->enabled || t->used >= LaunchTimer::kPairs) return;
  timer_record(t->ev[2 * t->used + 1], s);
  t->used++;
}
void launch_timer_fold(LaunchTimer* t, int first, int last) {
  for (int i = first; i < last; ++i) {
    float ms = 0.f;
    cudaEventSynchronize(t->ev[2 * i + 1]);
    if (cudaEventElapsedTime(&ms, t->ev[2 * i], t->ev[2 * i + 1]) == cudaSuccess) {
      t->total_ms[t->slot_of[i]] += ms;
      t->launches[t->slot_of[i]]++;
    }
  }
  cudaGetLastError();
}
int launch_timer_collect(LaunchTimer* t, cudaStream_t s) {
  if (!t) return FLOAM_OK;
  FLOAM_CUDA_OK(cudaStreamSynchronize(s));
  launch_timer_fold(t, t->persist, t->used);
  t->used = t->persist;
  return FLOAM_OK;
}

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = 4;                                 // keys per lane
constexpr int kSortItemsPerWarp = 32 * kSortRounds;            // 128: a warp owns a contiguous chunk (stability)
constexpr int kSortTile = kSortWarps * kSortItemsPerWarp;      // 1024 keys per CTA
constexpr int kDirectTiles = 256;                              // <= 262,144 keys: scatter CTAs scan the count table themselves

__device__ __forceinline__ int sort_tiles(int n) { return (n + kSortTile - 1) / kSortTile; }

// counts of this tile's digits -> hist. Layout [tile][digit] (direct mode) or [digit][tile] (scanned by single_block_scan_kernel).
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const unsigned int* __restrict__ keys, const int* __restrict__ d_n,
                                                                  const int* __restrict__ d_nbits, int shift, int* __restrict__ hist, const int* d_skip) {
  if (d_skip && *d_skip) return;
  if (shift >= *d_nbits) return;
  const int n = *d_n;
  const int tile0 = blockIdx.x * kSortTile;
  if (tile0 >= n) return;
  const int nb = sort_tiles(n);
  __shared__ int s_hist[256];
  unsigned int k[kSortRounds];
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int i = tile0 + r * kSortThreads + threadIdx.x;
    k[r] = (i < n) ? keys[i] : 0xffffffffu;
  }
  s_hist[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int i = tile0 + r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&s_hist[(k[r] >> shift) & 0xffu], 1);
  }
  __syncthreads();
  const int d = threadIdx.x;
  if (nb <= kDirectTiles) hist[blockIdx.x * 256 + d] = s_hist[d];
  else hist[d * nb + blockIdx.x] = s_hist[d];
}

// in-place exclusive scan of a small array by one block
// d_sort_n (optional): the array is the [digit][tile] radix count table of *d_sort_n keys; nothing to do in direct mode
__global__ void __launch_bounds__(1024) single_block_scan_kernel(int* __restrict__ data, int n, const int* __restrict__ d_nbits, int shift, const int* d_skip,
                                                                 const int* __restrict__ d_sort_n) {
  if (d_skip && *d_skip) return;
  if (d_nbits && shift >= *d_nbits) return;
  if (d_sort_n) {
    const int nb = sort_tiles(*d_sort_n);
    if (nb <= kDirectTiles) return;
    n = 256 * nb;
  }
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
                                                                     const int* __restrict__ hist, const int* d_skip) {
  if (d_skip && *d_skip) return;
  const int n = *d_n;
  const int tile0 = blockIdx.x * kSortTile;
  if (tile0 >= n) return;
  const int w = warp_id(), l = lane_id();
  const int begin = tile0 + w * kSortItemsPerWarp;
  // every load of the tile is in flight before anything is ranked
  unsigned int k[kSortRounds];
  int v[kSortRounds];
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int i = begin + r * 32 + l;
    k[r] = 0; v[r] = 0;
    if (i < n) { k[r] = keys_in[i]; v[r] = vals_in[i]; }
  }
  if (shift >= *d_nbits) {  // pass not needed for this key width: keep ping-pong parity with a straight copy
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
      const int i = begin + r * 32 + l;
      if (i < n) { keys_out[i] = k[r]; vals_out[i] = v[r]; }
    }
    return;
  }
  const int nb = sort_tiles(n);
  __shared__ int s_cnt[kSortWarps][256];
  __shared__ int s_scan[33];
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
  __syncthreads();
  unsigned int mask[kSortRounds];
  int* cnt = s_cnt[w];
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const bool valid = begin + r * 32 + l < n;
    const unsigned int d = valid ? ((k[r] >> shift) & 0xffu) : 0xffffffffu;
    mask[r] = __match_any_sync(0xffffffffu, d);
    if (valid && (__ffs(mask[r]) - 1) == l) cnt[d] += __popc(mask[r]);
    __syncwarp();
  }
  __syncthreads();
  {  // digit d = threadIdx.x: global base of (digit, tile), then the exclusive prefix over this tile's warps
    const int d = threadIdx.x;
    int base;
    if (nb <= kDirectTiles) {
      int total = 0, before = 0;
      const int b = blockIdx.x;
#pragma unroll 4
      for (int t = 0; t < nb; ++t) {
        const int c = hist[t * 256 + d];
        total += c;
        before += (t < b) ? c : 0;
      }
      int grand;
      base = block_excl_scan(total, s_scan, &grand) + before;
    } else {
      base = hist[d * nb + blockIdx.x];
    }
#pragma unroll
    for (int ww = 0; ww < kSortWarps; ++ww) {
      const int c = s_cnt[ww][d];
      s_cnt[ww][d] = base;
      base += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const bool valid = begin + r * 32 + l < n;
    const unsigned int d = (k[r] >> shift) & 0xffu;
    int pos = 0;
    if (valid) pos = cnt[d] + __popc(mask[r] & ((1u << l) - 1u));
    __syncwarp();
    if (valid && (__ffs(mask[r]) - 1) == l) cnt[d] += __popc(mask[r]);
    __syncwarp();
    if (valid) { keys_out[pos] = k[r]; vals_out[pos] = v[r]; }
  }
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
    FLOAM_LAUNCH(K_RADIX_HIST, radix_hist_kernel, nblocks, kSortThreads, s, kin, d_n, d_nbits, shift, ws.hist, d_skip);
    if (nblocks > kDirectTiles)  // only inputs that can exceed 262,144 keys need the separate scan of the count table
      FLOAM_LAUNCH(K_SINGLE_BLOCK_SCAN, single_block_scan_kernel, 1, 1024, s, ws.hist, 256 * nblocks, d_nbits, shift, d_skip, d_n);
    FLOAM_LAUNCH(K_RADIX_SCATTER, radix_scatter_kernel, nblocks, kSortThreads, s, kin, vin, kout, vout, d_n, d_nbits, shift, ws.hist, d_skip);
    unsigned int* tk = kin; kin = kout; kout = tk;
    int* tv = vin; vin = vout; vout = tv;
  }
}

void exclusive_scan_small(int* data, int n, cudaStream_t s) {
  FLOAM_LAUNCH(K_SINGLE_BLOCK_SCAN, single_block_scan_kernel, 1, 1024, s, data, n, nullptr, 0, nullptr, nullptr);
}

// ---- generic exclusive scan: two kernels, no serial single-CTA step -------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(const int* __restrict__ in, const int* __restrict__ d_n, int n_fixed,
                                                                  int* __restrict__ block_sums, const int* d_skip) {
  if (d_skip && *d_skip) return;
  const int n = d_n ? *d_n : n_fixed;
  if (blockIdx.x * kScanTile >= n) return;
  __shared__ int smem[33];
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int sum = 0;
  if (base + kScanItems <= n) {
    const int4 q = *reinterpret_cast<const int4*>(in + base);
    sum = q.x + q.y + q.z + q.w;
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) sum += (base + k < n) ? in[base + k] : 0;
  }
  const int total = block_sum(sum, smem);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ d_n, int n_fixed,
                                                                const int* __restrict__ block_sums, const int* d_skip) {
  if (d_skip && *d_skip) return;
  const int n = d_n ? *d_n : n_fixed;
  if (n == 0) { if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0; return; }
  if (blockIdx.x * kScanTile >= n) return;
  __shared__ int smem[33];
  const int offset = tile_offset(block_sums, blockIdx.x, smem);
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; sum += v[k]; }
  int total;
  int ex = block_excl_scan(sum, smem, &total) + offset;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
  if (blockIdx.x == (n - 1) / kScanTile && threadIdx.x == 0) out[n] = offset + total;  // grand total
}

}  // namespace

size_t scan_workspace_bytes(int n_max) { return ((size_t)(n_max + kScanTile - 1) / kScanTile + 2) * 4 + 256; }
void scan_workspace_bind(ScanWorkspace& ws, void* mem, int n_max) { ws.block_sums = (int*)mem; ws.n_max = n_max; }

void exclusive_scan_i32(const int* in, int* out, const int* d_n, int n_fixed, int n_max, ScanWorkspace& ws, const int* d_skip, cudaStream_t s) {
  const int nblocks = (n_max + kScanTile - 1) / kScanTile;
  FLOAM_LAUNCH(K_SCAN_TILES, scan_tiles_kernel, nblocks, kScanThreads, s, in, d_n, n_fixed, ws.block_sums, d_skip);
  FLOAM_LAUNCH(K_SCAN_ADD, scan_add_kernel, nblocks, kScanThreads, s, in, out, d_n, n_fixed, ws.block_sums, d_skip);
}

}  // namespace floam
