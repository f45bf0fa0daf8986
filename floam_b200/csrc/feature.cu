// Subsystem (1): LaserProcessingClass::featureExtraction on the device.
// Replaces reference src/laserProcessingClass.cpp:11-22 (RingExtractionVelodyne), :72-118 (curvature + sectors) and
// :121-231 (featureExtractionFromSector) with six kernels and no host round trip:
//   ring_count -> scan -> ring_scatter   : stable bucketing by ring with the range gate (order inside a ring preserved)
//   sector                               : one CTA per (ring, sector): ring segment staged in shared memory, curvature in the
//                                          reference's exact float/double operation order, bitonic sort by (value, id),
//                                          greedy edge pick with the +-5 neighbour suppression, surf = the rest
//   offsets -> gather                    : sector counts -> output offsets, points copied out as 2x float4
#include "feature.cuh"

namespace floam {
namespace {

constexpr int kTile = 1024;         // points per CTA in the bucketing kernels (32 warps, one point per thread)
constexpr int kMaxRings = 128;
constexpr int kSectorThreads = 256;
constexpr int kSectorCap = 1024;    // max curvature entries per sector (ring <= ~6150 points)
constexpr int kHalo = 10;

__device__ __forceinline__ bool range_gate(const PointIRT& p, double min_d, double max_d, int num_lines) {
  // double distance = sqrt(x*x + y*y) with float products/sum and the float sqrt overload (laserProcessingClass.cpp:14-15)
  const float d = __fsqrt_rn(fadd(fmul(p.x, p.x), fmul(p.y, p.y)));
  const double dd = (double)d;
  if (dd < min_d || dd > max_d) return false;
  return (int)p.ring < num_lines;
}

__device__ __forceinline__ PointIRT load_point(const PointIRT* __restrict__ pts, int i) {
  const float4* q = reinterpret_cast<const float4*>(pts + i);
  float4 a = __ldg(q), b = __ldg(q + 1);
  PointIRT p;
  *reinterpret_cast<float4*>(&p) = a;
  *(reinterpret_cast<float4*>(&p) + 1) = b;
  return p;
}
__device__ __forceinline__ void store_point(PointIRT* pts, int i, const PointIRT& p) {
  float4* q = reinterpret_cast<float4*>(pts + i);
  q[0] = *reinterpret_cast<const float4*>(&p);
  q[1] = *(reinterpret_cast<const float4*>(&p) + 1);
}

// tile_cnt[ring * ntiles + tile] = number of gated points of that ring in the tile
__global__ void __launch_bounds__(kTile) ring_count_kernel(const PointIRT* __restrict__ pts, const int* __restrict__ d_n, FeatureParams prm,
                                                            int ntiles, int* __restrict__ tile_cnt, int* __restrict__ d_flags) {
  pdl_prologue();
  __shared__ int s_cnt[kMaxRings];
  const int n = *d_n;
  if (threadIdx.x < kMaxRings) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * kTile + threadIdx.x;
  unsigned int ring = 0xffffffffu;
  if (i < n) {
    PointIRT p = load_point(pts, i);
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) atomicOr(d_flags, 1);  // Q9: undefined in the reference
    if (range_gate(p, prm.min_distance, prm.max_distance, prm.num_lines)) ring = p.ring;
  }
  const unsigned int m = __match_any_sync(0xffffffffu, ring);
  if (ring != 0xffffffffu && (__ffs(m) - 1) == lane_id()) atomicAdd(&s_cnt[ring], __popc(m));
  __syncthreads();
  if (threadIdx.x < prm.num_lines) tile_cnt[threadIdx.x * ntiles + blockIdx.x] = s_cnt[threadIdx.x];
  if (blockIdx.x == 0 && threadIdx.x == 0) tile_cnt[prm.num_lines * ntiles] = 0;  // becomes the grand total after the scan
}

__global__ void __launch_bounds__(kTile) ring_scatter_kernel(const PointIRT* __restrict__ pts, const int* __restrict__ d_n, FeatureParams prm,
                                                              int ntiles, const int* __restrict__ tile_off, PointIRT* __restrict__ ring_pts,
                                                              int* __restrict__ ring_src) {
  pdl_prologue();
  __shared__ int s_cnt[32][kMaxRings];  // per-warp counts -> exclusive prefix over warps
  const int n = *d_n;
  if (blockIdx.x * kTile >= n) return;
  for (int k = threadIdx.x; k < 32 * kMaxRings; k += kTile) (&s_cnt[0][0])[k] = 0;
  __syncthreads();
  const int i = blockIdx.x * kTile + threadIdx.x;
  const int w = warp_id(), l = lane_id();
  unsigned int ring = 0xffffffffu;
  PointIRT p;
  if (i < n) {
    p = load_point(pts, i);
    if (range_gate(p, prm.min_distance, prm.max_distance, prm.num_lines)) ring = p.ring;
  }
  const unsigned int m = __match_any_sync(0xffffffffu, ring);
  const int rank = __popc(m & ((1u << l) - 1u));
  if (ring != 0xffffffffu && rank == 0) s_cnt[w][ring] = __popc(m);
  __syncthreads();
  if (threadIdx.x < prm.num_lines) {
    int run = tile_off[threadIdx.x * ntiles + blockIdx.x];
    for (int ww = 0; ww < 32; ++ww) { const int c = s_cnt[ww][threadIdx.x]; s_cnt[ww][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
  if (ring != 0xffffffffu) {
    const int pos = s_cnt[w][ring] + rank;
    p.pad0 = 1.0f; p.pad1 = 0; p.pad2 = 0.0f;
    store_point(ring_pts, pos, p);
    ring_src[pos] = i;
  }
}

struct SortKey {
  double v;
  int id;
};
// (value, id) order. The values are sums of squares (never negative, never -0, never NaN: non-finite points do not reach this kernel), so
// the order of the bit patterns is the order of the values and the comparison runs on the integer pipe.
__device__ __forceinline__ bool key_less(double va, int ia, double vb, int ib) {
  const long long a = __double_as_longlong(va), b = __double_as_longlong(vb);
  return a < b || (a == b && ia < ib);
}

// One CTA per (ring, sector).
__global__ void __launch_bounds__(kSectorThreads) sector_kernel(const PointIRT* __restrict__ ring_pts, const int* __restrict__ tile_off, int ntiles,
                                                                FeatureParams prm, int* __restrict__ edge_tmp, int* __restrict__ surf_tmp,
                                                                int* __restrict__ edge_cnt, int* __restrict__ surf_cnt, int* __restrict__ d_flags) {
  pdl_prologue();
  __shared__ float sx[kSectorCap + kHalo], sy[kSectorCap + kHalo], sz[kSectorCap + kHalo];
  __shared__ double sval[kSectorCap], sval2[kSectorCap];   // curvature entries (value, id); the second pair is the sort's other exchange buffer
  __shared__ short sid[kSectorCap], sid2[kSectorCap];
  __shared__ unsigned char spicked[kSectorCap + kHalo], sgap[kSectorCap + kHalo];
  __shared__ int s_scan[33];
  __shared__ int s_nedge;

  const int ring = blockIdx.x / 6, sec = blockIdx.x % 6;
  const int ring_start = tile_off[ring * ntiles];
  const int ring_n = tile_off[(ring + 1) * ntiles] - ring_start;
  const int tid = threadIdx.x;
  if (ring_n < 131) {  // laserProcessingClass.cpp:89-91
    if (tid == 0) { edge_cnt[blockIdx.x] = 0; surf_cnt[blockIdx.x] = 0; }
    return;
  }
  const int total_points = ring_n - 10;
  const int sector_length = total_points / 6;
  const int sector_start = sector_length * sec;
  const int sector_end = (sec == 5) ? (total_points - 1) : (sector_length * (sec + 1) - 1);  // exclusive end (Q5)
  const int m = sector_end - sector_start;
  if (m > kSectorCap) {
    if (tid == 0) { atomicOr(d_flags, 2); edge_cnt[blockIdx.x] = 0; surf_cnt[blockIdx.x] = 0; }
    return;
  }
  // stage ring points [sector_start, sector_end + 10): local index q <-> ring index sector_start + q;
  // curvature entry c (0 <= c < m) is ring point j = sector_start + c + 5 = local c + 5
  const int nload = m + kHalo;
  const PointIRT* base = ring_pts + ring_start + sector_start;
  for (int q = tid; q < nload; q += kSectorThreads) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(base + q));
    sx[q] = a.x; sy[q] = a.y; sz[q] = a.z;
    spicked[q] = 0;
  }
  __syncthreads();
  int mpad = 32;
  while (mpad < m) mpad <<= 1;
  for (int c = tid; c < mpad; c += kSectorThreads) {
    if (c < m) {
      const int j = c + 5;
      // float left-to-right: p[j-5]+...+p[j-1] - 10*p[j] + p[j+1]+...+p[j+5]  (laserProcessingClass.cpp:96-98)
      float dx = fadd(sx[j - 5], sx[j - 4]); dx = fadd(dx, sx[j - 3]); dx = fadd(dx, sx[j - 2]); dx = fadd(dx, sx[j - 1]);
      dx = fsub(dx, fmul(10.0f, sx[j]));
      dx = fadd(dx, sx[j + 1]); dx = fadd(dx, sx[j + 2]); dx = fadd(dx, sx[j + 3]); dx = fadd(dx, sx[j + 4]); dx = fadd(dx, sx[j + 5]);
      float dy = fadd(sy[j - 5], sy[j - 4]); dy = fadd(dy, sy[j - 3]); dy = fadd(dy, sy[j - 2]); dy = fadd(dy, sy[j - 1]);
      dy = fsub(dy, fmul(10.0f, sy[j]));
      dy = fadd(dy, sy[j + 1]); dy = fadd(dy, sy[j + 2]); dy = fadd(dy, sy[j + 3]); dy = fadd(dy, sy[j + 4]); dy = fadd(dy, sy[j + 5]);
      float dz = fadd(sz[j - 5], sz[j - 4]); dz = fadd(dz, sz[j - 3]); dz = fadd(dz, sz[j - 2]); dz = fadd(dz, sz[j - 1]);
      dz = fsub(dz, fmul(10.0f, sz[j]));
      dz = fadd(dz, sz[j + 1]); dz = fadd(dz, sz[j + 2]); dz = fadd(dz, sz[j + 3]); dz = fadd(dz, sz[j + 4]); dz = fadd(dz, sz[j + 5]);
      const double X = dx, Y = dy, Z = dz;
      sval[c] = dadd(dadd(dmul(X, X), dmul(Y, Y)), dmul(Z, Z));
      sid[c] = (short)c;
    } else {
      sval[c] = __longlong_as_double(0x7ff0000000000000LL);  // +inf padding sorts last
      sid[c] = (short)c;
    }
  }
  // gap flag between local points q and q+1: squared distance (double, float differences) > 0.05  (:151-168)
  for (int q = tid; q < nload - 1; q += kSectorThreads) {
    const double ax = (double)fsub(sx[q + 1], sx[q]), ay = (double)fsub(sy[q + 1], sy[q]), az = (double)fsub(sz[q + 1], sz[q]);
    sgap[q] = dadd(dadd(dmul(ax, ax), dmul(ay, ay)), dmul(az, az)) > 0.05 ? 1 : 0;
  }
  __syncthreads();
  // Bitonic sort ascending by (value, id): the total order that stands in for std::sort's tie behaviour (Q8). The network runs on
  // registers: thread t holds entries t, t + 256, t + 512, t + 768, so a compare-exchange at distance j < 32 is a warp shuffle, at
  // distance j >= 256 it stays inside the thread, and only the distances 32, 64, 128 go through shared memory (double-buffered: one
  // barrier each). For the usual 512-entry sector that is 9 barriers instead of 45.
  // (A barrier-free rank sort — every entry counting its predecessors — was measured slower: m^2 64-bit compares per sector.)
  {
    constexpr int kSlots = kSectorCap / kSectorThreads;
    double v[kSlots];
    int id[kSlots];
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
      const int e = tid + sl * kSectorThreads;
      v[sl] = e < mpad ? sval[e] : __longlong_as_double(0x7ff0000000000000LL);
      id[sl] = e < mpad ? (int)sid[e] : e;
    }
    __syncthreads();   // every entry is in registers: the shared arrays become exchange buffers
    int buf = 0;
    for (int k = 2; k <= mpad; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        if (j >= kSectorThreads) {               // partner entry lives in another slot of the same thread
          static_assert(kSlots == 4, "slot pairs below are written out for four slots");
#define FLOAM_CX(A, B)                                                                                              \
  do {                                                                                                              \
    const int i = tid + (A) * kSectorThreads;                                                                       \
    if (i < mpad) {                                                                                                 \
      const bool up = (i & k) == 0;                                                                                 \
      const bool a_gt_b = key_less(v[B], id[B], v[A], id[A]);                                                       \
      if (a_gt_b == up) { const double tv = v[A]; v[A] = v[B]; v[B] = tv; const int ti = id[A]; id[A] = id[B]; id[B] = ti; } \
    }                                                                                                               \
  } while (0)
          if (j == kSectorThreads) { FLOAM_CX(0, 1); FLOAM_CX(2, 3); }
          else { FLOAM_CX(0, 2); FLOAM_CX(1, 3); }
#undef FLOAM_CX
        } else if (j >= 32) {                    // partner in another warp: through shared memory
          double* bv = buf ? sval2 : sval;
          short* bi = buf ? sid2 : sid;
#pragma unroll
          for (int sl = 0; sl < kSlots; ++sl) {
            const int i = tid + sl * kSectorThreads;
            if (i < mpad) { bv[i] = v[sl]; bi[i] = (short)id[sl]; }
          }
          __syncthreads();
#pragma unroll
          for (int sl = 0; sl < kSlots; ++sl) {
            const int i = tid + sl * kSectorThreads;
            if (i < mpad) {
              const double pv = bv[i ^ j];
              const int pi = bi[i ^ j];
              const bool keep_min = (((i & j) == 0) == ((i & k) == 0));     // the lower index of an ascending pair keeps the smaller key
              if (key_less(pv, pi, v[sl], id[sl]) == keep_min) { v[sl] = pv; id[sl] = pi; }   // keys are distinct (ids are): not less = greater
            }
          }
          buf ^= 1;   // the next shared-memory stage writes the other buffer; this one is rewritten only after that stage's barrier
        } else {                                 // partner in the same warp
#pragma unroll
          for (int sl = 0; sl < kSlots; ++sl) {
            const int i = tid + sl * kSectorThreads;
            if (sl * kSectorThreads < mpad) {    // warp-uniform: whole slots take part or not (mpad is a multiple of 32)
              const double pv = __shfl_xor_sync(0xffffffffu, v[sl], j);
              const int pi = __shfl_xor_sync(0xffffffffu, id[sl], j);
              const bool keep_min = (((i & j) == 0) == ((i & k) == 0));
              if (key_less(pv, pi, v[sl], id[sl]) == keep_min && i < mpad) { v[sl] = pv; id[sl] = pi; }
            }
          }
        }
      }
    }
    __syncthreads();   // the last exchange buffer may still be being read
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
      const int i = tid + sl * kSectorThreads;
      if (i < mpad) { sval[i] = v[sl]; sid[i] = (short)id[sl]; }
    }
    __syncthreads();
  }
  // greedy pick from the largest curvature down (:132-170); local point index of entry c is c + 5. One warp walks the sorted list 32
  // candidates at a time: a ballot finds the next un-suppressed candidate, lanes 0..9 apply the +-5 neighbour suppression (which
  // stops at the first gap > 0.05 m^2 on either side), and the batch is re-examined because the suppression may have hit it.
  if (tid < 32) {
    const int l = tid;
    int largestPickedNum = 0, nedge = 0;
    bool stop = false;
    for (int top = m - 1; top >= 0 && !stop; top -= 32) {
      const int i = top - l;                       // this lane's sorted position (descending curvature)
      const int q = i >= 0 ? sid[i] + 5 : 0;
      const bool above = i >= 0 && sval[i] > 0.1;
      unsigned int done_mask = 0;                  // lanes of this batch already handled
      for (;;) {
        const bool cand = i >= 0 && !((done_mask >> l) & 1u) && !spicked[q];
        const unsigned int cm = __ballot_sync(0xffffffffu, cand);
        if (cm == 0) break;
        const int src = __ffs(cm) - 1;             // largest remaining curvature of the batch
        if (!__shfl_sync(0xffffffffu, (int)above, src)) { stop = true; break; }   // sorted: nothing below qualifies either
        const int qs = __shfl_sync(0xffffffffu, q, src);
        largestPickedNum++;
        if (l == 0) spicked[qs] = 1;
        if (largestPickedNum <= 20) {
          if (l == 0) edge_tmp[blockIdx.x * 20 + nedge] = ring_start + sector_start + qs;
          nedge++;
        } else {
          stop = true;                             // the 21st candidate stays picked: neither edge nor surf (Q6)
          break;
        }
        // lanes 0..4: forward neighbours k = 1..5 (gap between qs+k-1 and qs+k); lanes 5..9: backward k = -1..-5 (gap at qs+k)
        const bool fwd = l < 5, bwd = l >= 5 && l < 10;
        const int k = fwd ? l + 1 : l - 4;
        bool gap = false;
        if (fwd) gap = sgap[qs + k - 1] != 0;
        if (bwd) gap = sgap[qs - k] != 0;
        const unsigned int gm = __ballot_sync(0xffffffffu, gap);
        const int first_f = (gm & 0x1fu) ? __ffs(gm & 0x1fu) - 1 : 5;          // suppression stops before the first gap
        const int first_b = ((gm >> 5) & 0x1fu) ? __ffs((gm >> 5) & 0x1fu) - 1 : 5;
        if (fwd && l < first_f) spicked[qs + k] = 1;
        if (bwd && (l - 5) < first_b) spicked[qs - k] = 1;
        __syncwarp();
        done_mask |= (src == 31) ? 0xffffffffu : ((2u << src) - 1u);           // everything up to and including src is settled
      }
    }
    if (l == 0) s_nedge = nedge;
  }
  __syncthreads();
  // surf = every non-picked entry in ascending curvature order (:220-227): stable compaction of the sorted list
  int written = 0;
  int* surf_out = surf_tmp + ring_start + sector_start + 5;
  for (int b0 = 0; b0 < m; b0 += kSectorThreads) {
    const int i = b0 + tid;
    int keep = 0, q = 0;
    if (i < m) { q = sid[i] + 5; keep = spicked[q] ? 0 : 1; }
    int tot;
    const int ex = block_excl_scan(keep, s_scan, &tot);
    if (keep) surf_out[written + ex] = ring_start + sector_start + q;
    written += tot;
    __syncthreads();
  }
  if (tid == 0) { edge_cnt[blockIdx.x] = s_nedge; surf_cnt[blockIdx.x] = written; }
}

// exclusive offsets over <= 1024 sectors; totals to d_ne / d_ns
__global__ void __launch_bounds__(1024) feature_offsets_kernel(const int* __restrict__ edge_cnt, const int* __restrict__ surf_cnt, int nsectors,
                                                                int* __restrict__ edge_off, int* __restrict__ surf_off, int* d_ne, int* d_ns) {
  pdl_prologue();
  __shared__ int smem[33];
  const int t = threadIdx.x;
  int e = t < nsectors ? edge_cnt[t] : 0, s = t < nsectors ? surf_cnt[t] : 0;
  int te, ts;
  const int ee = block_excl_scan(e, smem, &te);
  __syncthreads();
  const int se = block_excl_scan(s, smem, &ts);
  if (t < nsectors) { edge_off[t] = ee; surf_off[t] = se; }
  if (t == 0) { *d_ne = te; *d_ns = ts; }
}

__global__ void __launch_bounds__(256) feature_gather_kernel(const PointIRT* __restrict__ ring_pts, const int* __restrict__ ring_src,
                                                              const int* __restrict__ tile_off, int ntiles, const int* __restrict__ edge_tmp,
                                                              const int* __restrict__ surf_tmp, const int* __restrict__ edge_cnt,
                                                              const int* __restrict__ surf_cnt, const int* __restrict__ edge_off,
                                                              const int* __restrict__ surf_off, PointIRT* __restrict__ edge_out,
                                                              PointIRT* __restrict__ surf_out, int* __restrict__ edge_src, int* __restrict__ surf_src) {
  pdl_prologue();
  const int sector = blockIdx.x;
  const int ne = edge_cnt[sector], ns = surf_cnt[sector];
  if (ne == 0 && ns == 0) return;
  const int ring = sector / 6, sec = sector % 6;
  const int ring_start = tile_off[ring * ntiles];
  const int ring_n = tile_off[(ring + 1) * ntiles] - ring_start;
  const int sector_start = ((ring_n - 10) / 6) * sec;
  const int eo = edge_off[sector], so = surf_off[sector];
  // each point is two float4: even threads move the first half, odd threads the second
  const int half = threadIdx.x & 1;
  for (int k = threadIdx.x >> 1; k < ne; k += 128) {
    const int src = edge_tmp[sector * 20 + k];
    reinterpret_cast<float4*>(edge_out + eo + k)[half] = __ldg(reinterpret_cast<const float4*>(ring_pts + src) + half);
    if (!half) edge_src[eo + k] = ring_src[src];
  }
  const int* st = surf_tmp + ring_start + sector_start + 5;
  for (int k = threadIdx.x >> 1; k < ns; k += 128) {
    const int src = st[k];
    reinterpret_cast<float4*>(surf_out + so + k)[half] = __ldg(reinterpret_cast<const float4*>(ring_pts + src) + half);
    if (!half) surf_src[so + k] = ring_src[src];
  }
}

}  // namespace

size_t feature_workspace_bytes(int max_scan_points, int num_lines) {
  const int ntiles = (max_scan_points + kTile - 1) / kTile;
  size_t b = 0;
  b += (size_t)max_scan_points * sizeof(PointIRT);      // ring_pts
  b += (size_t)max_scan_points * 4 * 2;                 // ring_src, surf_tmp
  b += ((size_t)num_lines * ntiles + 1) * 4;            // tile_cnt/off
  b += (size_t)num_lines * 6 * (20 + 4) * 4;            // edge_tmp, counts, offsets
  return b + 1024;
}

void feature_workspace_bind(FeatureWorkspace& ws, void* mem, int max_scan_points, int num_lines) {
  char* p = (char*)mem;
  auto take = [&](size_t bytes) { void* r = p; p += (bytes + 255) / 256 * 256; return r; };
  ws.max_scan_points = max_scan_points;
  ws.ntiles = (max_scan_points + kTile - 1) / kTile;
  ws.nsectors = num_lines * 6;
  ws.ring_pts = (PointIRT*)take((size_t)max_scan_points * sizeof(PointIRT));
  ws.ring_src = (int*)take((size_t)max_scan_points * 4);
  ws.surf_tmp = (int*)take((size_t)max_scan_points * 4);
  ws.tile_off = (int*)take(((size_t)num_lines * ws.ntiles + 1) * 4);
  ws.edge_tmp = (int*)take((size_t)ws.nsectors * 20 * 4);
  ws.edge_cnt = (int*)take((size_t)ws.nsectors * 4);
  ws.surf_cnt = (int*)take((size_t)ws.nsectors * 4);
  ws.edge_off = (int*)take((size_t)ws.nsectors * 4);
  ws.surf_off = (int*)take((size_t)ws.nsectors * 4);
}

size_t feature_workspace_bytes_padded(int max_scan_points, int num_lines) { return feature_workspace_bytes(max_scan_points, num_lines) + 16 * 256; }

void feature_extract_device(const PointIRT* d_scan, const int* d_n, const FeatureParams& prm, FeatureWorkspace& ws, PointIRT* d_edge, int* d_ne,
                            PointIRT* d_surf, int* d_ns, int* d_edge_src, int* d_surf_src, int* d_flags, cudaStream_t s) {
  const int ntiles = ws.ntiles;
  FLOAM_LAUNCH(K_RING_COUNT, ring_count_kernel, ntiles, kTile, s, d_scan, d_n, prm, ntiles, ws.tile_off, d_flags);
  exclusive_scan_small(ws.tile_off, prm.num_lines * ntiles + 1, s);
  FLOAM_LAUNCH(K_RING_SCATTER, ring_scatter_kernel, ntiles, kTile, s, d_scan, d_n, prm, ntiles, ws.tile_off, ws.ring_pts, ws.ring_src);
  FLOAM_LAUNCH(K_SECTOR, sector_kernel, ws.nsectors, kSectorThreads, s, ws.ring_pts, ws.tile_off, ntiles, prm, ws.edge_tmp, ws.surf_tmp, ws.edge_cnt, ws.surf_cnt, d_flags);
  FLOAM_LAUNCH(K_FEATURE_OFFSETS, feature_offsets_kernel, 1, 1024, s, ws.edge_cnt, ws.surf_cnt, ws.nsectors, ws.edge_off, ws.surf_off, d_ne, d_ns);
  FLOAM_LAUNCH(K_FEATURE_GATHER, feature_gather_kernel, ws.nsectors, 256, s, ws.ring_pts, ws.ring_src, ws.tile_off, ntiles, ws.edge_tmp, ws.surf_tmp, ws.edge_cnt, ws.surf_cnt,
                                                    ws.edge_off, ws.surf_off, d_edge, d_surf, d_edge_src, d_surf_src);
}

}  // namespace floam
