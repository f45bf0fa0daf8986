// Device-wide building blocks: stable LSD radix sort of (key,value) pairs and exclusive scan.
// Element counts live in device memory, so the callers (voxel filter, grid build) never synchronise with the host.
// Everything here is sized for latency on SMALL inputs (5k - 200k elements per frame, all L2-resident): tiles of 1024 keys so
// that even a 20k-element sort spreads over 20+ CTAs, every global load issued before its first use, and — up to 262,144 keys —
// no separate scan kernel: each scatter CTA derives its digit bases straight from the [tile][digit] count table.
#include <cstdint>

#include "common.cuh"

namespace floam {

thread_local long long g_launches = 0;
thread_local LaunchTimer* g_timer = nullptr;
bool g_use_pdl = false;
bool g_pdl_solve = false;
thread_local bool t_pdl_scope = false;

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
  "voxel_reduce",
  "repack",
  "crop_flags",
  "crop_scatter",
  "record_pose",
  "mail_state",
  "unpack_pc2",
  "voxel_classify",
  "voxel_merge",
  "noop"
};
const char* kernel_slot_name(int slot) { return (slot >= 0 && slot < K_NUM_SLOTS) ? kSlotNames[slot] : "?"; }

__global__ void noop_kernel() {}
void launch_noop(cudaStream_t s) { g_launches++; if (g_timer) launch_timer_begin(K_NOOP, s); noop_kernel<<<1, 32, 0, s>>>(); if (g_timer) launch_timer_end(s); }

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

// Stable LSD radix sort in exactly THREE passes whatever the key width: the digit width is ceil(nbits / 3) bits, decided on the
// device (nbits <= 24 -> the usual 8-bit digits or narrower; up to 11 bits = 2048 bins for 31-bit keys, on a slower but correct
// path). A pass is two kernels: per-tile digit counts (+ the scan of the count table by the last CTA to finish, only needed above
// 262,144 keys) and the ranking scatter.
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortMinRounds = 4;                              // keys per lane: 4 (tile = 1024 keys) for small inputs ...
constexpr int kSortMaxRounds = 32;                             // ... up to 32 (tile = 8192 keys) for large ones
constexpr int kSortTile = kSortThreads * kSortMinRounds;       // smallest tile: the launch grid is sized for it
constexpr int kDirectTiles = 256;                              // up to this many tiles the scatter CTAs scan the count table themselves
constexpr int kMaxBins = 2048;
constexpr int kStageTableBytes = 128 * 1024;                    // count table staged in shared memory when it fits (128 rows x 256 bins)
constexpr int kNarrowBins = 1024;                              // digits up to 10 bits (keys up to 30 bits) rank with per-warp counters                                 // 11-bit digits at most (3 x 11 >= 31 key bits)

// Keys per lane, decided on the device from the live element count: the tile grows with the input so that the [tile][digit] count
// table stays around 128 rows — every scatter CTA reads all of it, which is quadratic in the number of tiles (at a fixed 1024-key
// tile a 236k-key sort spent 47 us per pass there).
// Wide digits (keys above 24 bits: 512 or 1024 bins) widen the rows, so they get proportionally fewer of them.
__device__ __forceinline__ int sort_rounds(int n, int nbits) {
  const int bits = (nbits + 2) / 3;
  const int rows = bits <= 8 ? 128 : 64;
  int r = kSortMinRounds;
  while (r < kSortMaxRounds && n > rows * kSortThreads * r) r <<= 1;
  return r;
}
__device__ __forceinline__ int sort_tiles(int n, int rounds) { const int tile = kSortThreads * rounds; return (n + tile - 1) / tile; }
__device__ __forceinline__ int digit_bits(int nbits) {
  const int b = (nbits + 2) / 3;
  return b < 1 ? 1 : b;
}

// The last CTA to finish (ticket) turns a [digit][tile] count table into exclusive offsets in place: large inputs only.
__device__ __forceinline__ void scan_table_by_last_cta(int* __restrict__ hist, int total_entries, unsigned int* ticket, int nb, int* s_scan, int& s_last) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned int)nb - 1u) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  int carry = 0;
  for (int base = 0; base < total_entries; base += kSortThreads * 4) {
    const int i = base + threadIdx.x * 4;
    int v[4], sum = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) { v[u] = (i + u < total_entries) ? __ldcg(hist + i + u) : 0; sum += v[u]; }
    int total;
    int ex = block_excl_scan(sum, s_scan, &total) + carry;
#pragma unroll
    for (int u = 0; u < 4; ++u) { if (i + u < total_entries) hist[i + u] = ex; ex += v[u]; }
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *ticket = 0;
}

// Pass 0 only: counts of this tile's digits -> hist. Layout [tile][digit] (direct mode) or [digit][tile], which the last CTA then scans
// in place. The count tables of passes 1 and 2 (hist + table_stride, hist + 2 * table_stride) are filled by the scatter kernel of the
// pass before them — every key adds itself to the (tile, digit) cell of the position it is written to — so this kernel zeroes this
// tile's cells in both.
template <int MAXR>
__device__ __forceinline__ void radix_hist_body(const unsigned int* __restrict__ keys, const int* __restrict__ d_n,
                                                                  const int* __restrict__ d_nbits, int pass, int* __restrict__ hist, int table_stride,
                                                                  unsigned int* ticket, int* s_hist, int* s_scan, int& s_last, int tile) {
  const int n = *d_n;
  const int R = MAXR;
  const int tile0 = tile * kSortThreads * R;
  if (tile0 >= n) return;
  const int bits = digit_bits(*d_nbits), shift = pass * bits, nbins = 1 << bits;
  const unsigned int dmask = (unsigned int)nbins - 1u;
  const int nb = sort_tiles(n, R);
  unsigned int k[MAXR];
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    const int i = tile0 + r * kSortThreads + threadIdx.x;
    k[r] = (r < R && i < n) ? keys[i] : 0xffffffffu;
  }
  for (int d = threadIdx.x; d < nbins; d += kSortThreads) s_hist[d] = 0;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    const int i = tile0 + r * kSortThreads + threadIdx.x;
    if (r < R && i < n) atomicAdd(&s_hist[(k[r] >> shift) & dmask], 1);
  }
  __syncthreads();
  if (nb <= kDirectTiles) {
    for (int d = threadIdx.x; d < nbins; d += kSortThreads) {
      const int cell = tile * nbins + d;
      hist[cell] = s_hist[d];
      hist[table_stride + cell] = 0;
      hist[2 * table_stride + cell] = 0;
    }
    return;
  }
  for (int d = threadIdx.x; d < nbins; d += kSortThreads) {
    const int cell = d * nb + tile;
    hist[cell] = s_hist[d];
    hist[table_stride + cell] = 0;
    hist[2 * table_stride + cell] = 0;
  }
  scan_table_by_last_cta(hist, nbins * nb, ticket, nb, s_scan, s_last);
}

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const unsigned int* __restrict__ keys, const int* __restrict__ d_n,
                                                                  const int* __restrict__ d_nbits, int pass, int* __restrict__ hist, int table_stride,
                                                                  unsigned int* ticket, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  __shared__ int s_hist[kMaxBins];
  __shared__ int s_scan[33];
  __shared__ int s_last;
  // The grid is sized for the hardware (one wave), not for the capacity: CTA b takes tiles b, b + gridDim.x, ... of the live input.
  // (A capacity-sized grid spent most of the kernel draining thousands of CTAs that load two words and exit.)
  const int n = *d_n;
  const int rounds = sort_rounds(n, *d_nbits);
  const int nb = sort_tiles(n, rounds);
  for (int tile = blockIdx.x; tile < nb; tile += gridDim.x) {
    switch (rounds) {   // one fully unrolled variant per tile size; the small-input one stays as tight as a fixed-size kernel
      case 4: radix_hist_body<4>(keys, d_n, d_nbits, pass, hist, table_stride, ticket, s_hist, s_scan, s_last, tile); break;
      case 8: radix_hist_body<8>(keys, d_n, d_nbits, pass, hist, table_stride, ticket, s_hist, s_scan, s_last, tile); break;
      case 16: radix_hist_body<16>(keys, d_n, d_nbits, pass, hist, table_stride, ticket, s_hist, s_scan, s_last, tile); break;
      default: radix_hist_body<32>(keys, d_n, d_nbits, pass, hist, table_stride, ticket, s_hist, s_scan, s_last, tile); break;
    }
    __syncthreads();   // the shared arrays are reused by the next tile
  }
}

// in-place exclusive scan of a small array by one block
__global__ void __launch_bounds__(1024) single_block_scan_kernel(int* __restrict__ data, int n) {
  pdl_prologue();
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

template <int MAXR>
__device__ __forceinline__ void radix_scatter_body(const unsigned int* __restrict__ keys_in, const int* __restrict__ vals_in,
                                                                     unsigned int* __restrict__ keys_out, int* __restrict__ vals_out,
                                                                     const int* __restrict__ d_n, const int* __restrict__ d_nbits, int pass,
                                                                     const int* __restrict__ hist, int* __restrict__ hist_next, unsigned int* ticket,
                                                                     int (*s_cnt)[kNarrowBins], int* s_base, int* s_scan, int& s_last, int tile,
                                                                     const int* s_table, unsigned int bar_addr, bool& table_ready) {
  const int n = *d_n;
  const int R = MAXR;
  const int tile0 = tile * kSortThreads * R;
  if (tile0 >= n) return;
  constexpr int kTileShift = MAXR == 4 ? 10 : MAXR == 8 ? 11 : MAXR == 16 ? 12 : 13;   // log2(kSortThreads * MAXR)
  const int w = warp_id(), l = lane_id();
  const int begin = tile0 + w * 32 * R;   // a warp owns a contiguous chunk of the tile (stability)
  // every load of the tile is in flight before anything is ranked
  unsigned int k[MAXR];
  int v[MAXR];
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    const int i = begin + r * 32 + l;
    k[r] = 0; v[r] = 0;
    if (r < R && i < n) { k[r] = keys_in[i]; v[r] = vals_in[i]; }
  }
  const int bits = digit_bits(*d_nbits), shift = pass * bits, nbins = 1 << bits;
  const unsigned int dmask = (unsigned int)nbins - 1u;
  const int nb = sort_tiles(n, R);
  const bool direct = nb <= kDirectTiles;
  const bool narrow = nbins <= kNarrowBins;
  unsigned int mask[MAXR];
  int* cnt = s_cnt[w];
  if (narrow) {
    for (int ww = 0; ww < kSortWarps; ++ww)
      for (int d = threadIdx.x; d < nbins; d += kSortThreads) s_cnt[ww][d] = 0;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      if (r >= R) break;
      const bool valid = begin + r * 32 + l < n;
      const unsigned int d = valid ? ((k[r] >> shift) & dmask) : 0xffffffffu;
      mask[r] = __match_any_sync(0xffffffffu, d);
      if (valid && (__ffs(mask[r]) - 1) == l) cnt[d] += __popc(mask[r]);
      __syncwarp();
    }
    __syncthreads();
  }
  // the count table staged by the bulk copy (issued before the tile's loads) must have landed by now
  if (s_table && !table_ready) {
    asm volatile("{\n .reg .pred p;\n WAIT_TABLE:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra TABLE_DONE;\n bra WAIT_TABLE;\n TABLE_DONE:\n}"
                 :: "r"(bar_addr) : "memory");
    table_ready = true;
  }
  // global base of (digit, this tile): everything with a smaller digit, plus the same digit in earlier tiles
  int carry = 0;
  for (int d0 = 0; d0 < nbins; d0 += kSortThreads) {
    const int d = d0 + threadIdx.x;
    int base = 0, total = 0;
    if (nb <= kDirectTiles) {
      if (d < nbins) {
        int before = 0;
        const int b = tile;
        // nb independent L2 reads per thread: unrolled 16-deep so that 16 are in flight (at 4 this loop WAS the kernel: 57 tiles ->
        // 14 dependent round trips)
        if (s_table) {   // shared-memory copy of the whole table: consecutive digits, no bank conflicts
#pragma unroll 8
          for (int t = 0; t < nb; ++t) {
            const int c = s_table[t * nbins + d];
            total += c;
            before += (t < b) ? c : 0;
          }
        } else {
#pragma unroll 16
          for (int t = 0; t < nb; ++t) {
            const int c = __ldg(hist + t * nbins + d);
            total += c;
            before += (t < b) ? c : 0;
          }
        }
        base = before;
      }
      int grand;
      base += block_excl_scan(total, s_scan, &grand) + carry;
      carry += grand;
      __syncthreads();
    } else if (d < nbins) {
      base = hist[d * nb + tile];
    }
    if (d < nbins) {
      if (narrow) {
#pragma unroll
        for (int ww = 0; ww < kSortWarps; ++ww) {
          const int c = s_cnt[ww][d];
          s_cnt[ww][d] = base;
          base += c;
        }
      } else {
        s_base[d] = base;
      }
    }
  }
  __syncthreads();
  if (narrow) {
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      if (r >= R) break;
      const bool valid = begin + r * 32 + l < n;
      const unsigned int d = (k[r] >> shift) & dmask;
      int pos = 0;
      if (valid) pos = cnt[d] + __popc(mask[r] & ((1u << l) - 1u));
      __syncwarp();
      if (valid && (__ffs(mask[r]) - 1) == l) cnt[d] += __popc(mask[r]);
      __syncwarp();
      if (valid) {
        keys_out[pos] = k[r]; vals_out[pos] = v[r];
        if (hist_next) {   // this key counts itself into the next pass's table, at the tile it lands in
          const int dn = (int)((k[r] >> (shift + bits)) & dmask), tn = pos >> kTileShift;
          atomicAdd(hist_next + (direct ? tn * nbins + dn : dn * nb + tn), 1);
        }
      }
    }
  } else {
    // wide digits (keys above 24 bits): the warps of the tile rank one after the other against the shared running offsets
    for (int ww = 0; ww < kSortWarps; ++ww) {
      if (w == ww) {
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
          if (r >= R) break;
          const bool valid = begin + r * 32 + l < n;
          const unsigned int d = valid ? ((k[r] >> shift) & dmask) : 0xffffffffu;
          const unsigned int m = __match_any_sync(0xffffffffu, d);
          int pos = 0;
          if (valid) pos = s_base[d] + __popc(m & ((1u << l) - 1u));
          __syncwarp();
          if (valid && (__ffs(m) - 1) == l) s_base[d] += __popc(m);
          __syncwarp();
          if (valid) {
            keys_out[pos] = k[r]; vals_out[pos] = v[r];
            if (hist_next) {
              const int dn = (int)((k[r] >> (shift + bits)) & dmask), tn = pos >> kTileShift;
              atomicAdd(hist_next + (direct ? tn * nbins + dn : dn * nb + tn), 1);
            }
          }
        }
      }
      __syncthreads();
    }
  }
  if (hist_next && !direct) scan_table_by_last_cta(hist_next, nbins * nb, ticket, nb, s_scan, s_last);
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const unsigned int* __restrict__ keys_in, const int* __restrict__ vals_in,
                                                                     unsigned int* __restrict__ keys_out, int* __restrict__ vals_out,
                                                                     const int* __restrict__ d_n, const int* __restrict__ d_nbits, int pass,
                                                                     const int* __restrict__ hist, int* __restrict__ hist_next, unsigned int* ticket,
                                                                     const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  __shared__ int s_cnt[kSortWarps][kNarrowBins];   // digits up to 10 bits: per-warp counters / running offsets (32 KB)
  __shared__ int s_base[kMaxBins];         // wide digits: one running offset per bin, warps take turns
  __shared__ int s_scan[33];
  __shared__ int s_last;
  const int n = *d_n;
  const int nbits = *d_nbits;
  const int rounds = sort_rounds(n, nbits);
  const int nb = sort_tiles(n, rounds);
  // The [tile][digit] count table of this pass (every scatter CTA needs all of it) comes in with ONE bulk copy (TMA, cp.async.bulk
  // global -> shared, completion on an mbarrier) issued before the tile's own loads, instead of nb dependent batches of L2 reads.
  extern __shared__ __align__(128) unsigned char s_dyn[];
  __shared__ __align__(8) unsigned long long s_bar;
  const int table_bytes = (nb * (1 << digit_bits(nbits)) * 4 + 15) & ~15;
  const bool staged = nb <= kDirectTiles && table_bytes <= kStageTableBytes && (int)blockIdx.x < nb;
  const unsigned int bar_addr = (unsigned int)__cvta_generic_to_shared(&s_bar);
  if (staged) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_addr));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int dst = (unsigned int)__cvta_generic_to_shared(s_dyn);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_addr), "r"(table_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(dst), "l"(hist), "r"(table_bytes), "r"(bar_addr) : "memory");
    }
  }
  const int* s_table = staged ? reinterpret_cast<const int*>(s_dyn) : nullptr;
  bool table_ready = false;
  for (int tile = blockIdx.x; tile < nb; tile += gridDim.x) {   // one wave of CTAs, each looping over its tiles (see radix_hist_kernel)
    switch (rounds) {
      case 4: radix_scatter_body<4>(keys_in, vals_in, keys_out, vals_out, d_n, d_nbits, pass, hist, hist_next, ticket, s_cnt, s_base, s_scan, s_last, tile, s_table, bar_addr, table_ready); break;
      case 8: radix_scatter_body<8>(keys_in, vals_in, keys_out, vals_out, d_n, d_nbits, pass, hist, hist_next, ticket, s_cnt, s_base, s_scan, s_last, tile, s_table, bar_addr, table_ready); break;
      case 16: radix_scatter_body<16>(keys_in, vals_in, keys_out, vals_out, d_n, d_nbits, pass, hist, hist_next, ticket, s_cnt, s_base, s_scan, s_last, tile, s_table, bar_addr, table_ready); break;
      default: radix_scatter_body<32>(keys_in, vals_in, keys_out, vals_out, d_n, d_nbits, pass, hist, hist_next, ticket, s_cnt, s_base, s_scan, s_last, tile, s_table, bar_addr, table_ready); break;
    }
    __syncthreads();
  }
}

}  // namespace

// rows of one count table: the tile grows with the input (sort_rounds), so at most 128 tiles up to 128 * 8192 keys, n / 8192 beyond
static int table_rows(int n_max) {
  const int big = (n_max + kSortThreads * kSortMaxRounds - 1) / (kSortThreads * kSortMaxRounds);
  return (big > 128 ? big : 128) + 1;
}
size_t sort_workspace_bytes(int n_max) {
  return (size_t)n_max * 8 + (size_t)3 * kMaxBins * table_rows(n_max) * 4 + 1024;
}
void sort_workspace_bind(SortWorkspace& ws, void* mem, int n_max) {
  char* p = (char*)mem;
  ws.n_max = n_max;
  ws.max_blocks = (n_max + kSortTile - 1) / kSortTile;
  ws.keys_alt = (unsigned int*)p; p += (size_t)n_max * 4;
  ws.vals_alt = (int*)p; p += (size_t)n_max * 4;
  p = (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
  ws.ticket = (unsigned int*)p; p += 256;
  ws.hist = (int*)p;   // 256-byte aligned: the scatter kernel bulk-copies whole tables
  ws.table_stride = kMaxBins * table_rows(n_max);
}
int sort_workspace_arm(SortWorkspace& ws, cudaStream_t s) {
  // per device: the scatter kernel stages its count table in up to 128 KB of dynamic shared memory
  FLOAM_CUDA_OK(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageTableBytes));
  FLOAM_CUDA_OK(cudaMemsetAsync(ws.ticket, 0, 256, s));
  return FLOAM_OK;
}

void radix_sort_pairs_from(unsigned int* keys, int* vals, unsigned int* keys_alt, int* vals_alt, const int* d_n, const int* d_nbits, int n_max,
                           SortWorkspace& ws, const int* d_skip, cudaStream_t s, unsigned int** sorted_keys, int** sorted_vals) {
  if (n_max > ws.n_max) n_max = ws.n_max;
  int nblocks = (n_max + kSortTile - 1) / kSortTile;
  if (nblocks > kNumSMs) nblocks = kNumSMs;   // one wave (the scatter kernel's registers allow one CTA per SM); CTAs loop over tiles
  unsigned int* kin = keys; int* vin = vals;
  unsigned int* kout = keys_alt; int* vout = vals_alt;
  // four launches: digit counts of pass 0, then three scatters, each of which also counts the digits of the pass after it
  FLOAM_LAUNCH(K_RADIX_HIST, radix_hist_kernel, nblocks, kSortThreads, s, kin, d_n, d_nbits, 0, ws.hist, ws.table_stride, ws.ticket, d_skip);
  for (int pass = 0; pass < 3; ++pass) {
    FLOAM_LAUNCH_DYN(K_RADIX_SCATTER, radix_scatter_kernel, nblocks, kSortThreads, kStageTableBytes, s, kin, vin, kout, vout, d_n, d_nbits, pass, ws.hist + (size_t)pass * ws.table_stride,
                 pass < 2 ? ws.hist + (size_t)(pass + 1) * ws.table_stride : (int*)nullptr, ws.ticket, d_skip);
    unsigned int* tk = kin; kin = kout; kout = tk;
    int* tv = vin; vin = vout; vout = tv;
  }
  *sorted_keys = kin;   // three passes: the result sits in the alternate buffers
  *sorted_vals = vin;
}

void radix_sort_pairs(unsigned int* keys, int* vals, const int* d_n, const int* d_nbits, int n_max, SortWorkspace& ws, const int* d_skip, cudaStream_t s,
                      unsigned int** sorted_keys, int** sorted_vals) {
  radix_sort_pairs_from(keys, vals, ws.keys_alt, ws.vals_alt, d_n, d_nbits, n_max, ws, d_skip, s, sorted_keys, sorted_vals);
}

void exclusive_scan_small(int* data, int n, cudaStream_t s) {
  FLOAM_LAUNCH(K_SINGLE_BLOCK_SCAN, single_block_scan_kernel, 1, 1024, s, data, n);
}

// ---- generic exclusive scan: two kernels, no serial single-CTA step -------------------------------------------------------------
namespace {

// Both kernels run one wave of CTAs that loop over the live tiles (grids sized for the hardware, not for the capacity).
__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(const int* __restrict__ in, const int* __restrict__ d_n, int n_fixed,
                                                                  int* __restrict__ block_sums, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  const int n = d_n ? *d_n : n_fixed;
  __shared__ int smem[33];
  const int ntiles = (n + kScanTile - 1) / kScanTile;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int base = tile * kScanTile + threadIdx.x * kScanItems;
    int sum = 0;
    if (base + kScanItems <= n) {
      const int4 q = *reinterpret_cast<const int4*>(in + base);
      sum = q.x + q.y + q.z + q.w;
    } else {
#pragma unroll
      for (int k = 0; k < kScanItems; ++k) sum += (base + k < n) ? in[base + k] : 0;
    }
    const int total = block_sum(sum, smem);
    if (threadIdx.x == 0) block_sums[tile] = total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ d_n, int n_fixed,
                                                                const int* __restrict__ block_sums, const int* d_skip) {
  pdl_prologue();
  if (d_skip && *d_skip) return;
  const int n = d_n ? *d_n : n_fixed;
  if (n == 0) { if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0; return; }
  __shared__ int smem[33];
  const int ntiles = (n + kScanTile - 1) / kScanTile;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int offset = tile_offset(block_sums, tile, smem);
    const int base = tile * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; sum += v[k]; }
    int total;
    int ex = block_excl_scan(sum, smem, &total) + offset;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
    if (tile == ntiles - 1 && threadIdx.x == 0) out[n] = offset + total;  // grand total
    __syncthreads();
  }
}

}  // namespace

size_t scan_workspace_bytes(int n_max) { return ((size_t)(n_max + kScanTile - 1) / kScanTile + 2) * 4 + 256; }
void scan_workspace_bind(ScanWorkspace& ws, void* mem, int n_max) { ws.block_sums = (int*)mem; ws.n_max = n_max; }

static int scan_grid(int n_max) {
  const int nblocks = (n_max + kScanTile - 1) / kScanTile;
  return nblocks > 2 * kNumSMs ? 2 * kNumSMs : nblocks;
}
void exclusive_scan_i32(const int* in, int* out, const int* d_n, int n_fixed, int n_max, ScanWorkspace& ws, const int* d_skip, cudaStream_t s) {
  const int nblocks = scan_grid(n_max);
  FLOAM_LAUNCH(K_SCAN_TILES, scan_tiles_kernel, nblocks, kScanThreads, s, in, d_n, n_fixed, ws.block_sums, d_skip);
  FLOAM_LAUNCH(K_SCAN_ADD, scan_add_kernel, nblocks, kScanThreads, s, in, out, d_n, n_fixed, ws.block_sums, d_skip);
}

void exclusive_scan_with_tile_sums(const int* in, int* out, const int* d_n, int n_max, const int* tile_sums, const int* d_skip, cudaStream_t s) {
  const int nblocks = scan_grid(n_max);
  FLOAM_LAUNCH(K_SCAN_ADD, scan_add_kernel, nblocks, kScanThreads, s, in, out, d_n, 0, tile_sums, d_skip);
}

}  // namespace floam
