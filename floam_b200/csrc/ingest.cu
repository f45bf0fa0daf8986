#include "ingest.cuh"

namespace floam {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ unsigned int load_u32(const unsigned char* __restrict__ p, bool big) {
  const unsigned int b0 = __ldg(p), b1 = __ldg(p + 1), b2 = __ldg(p + 2), b3 = __ldg(p + 3);
  return big ? (b0 << 24 | b1 << 16 | b2 << 8 | b3) : (b3 << 24 | b2 << 16 | b1 << 8 | b0);
}
__device__ __forceinline__ unsigned short load_u16(const unsigned char* __restrict__ p, bool big) {
  const unsigned int b0 = __ldg(p), b1 = __ldg(p + 1);
  return (unsigned short)(big ? (b0 << 8 | b1) : (b1 << 8 | b0));
}

__global__ void __launch_bounds__(kThreads) unpack_pc2_kernel(const unsigned char* __restrict__ raw, floam_pc2_layout L, PointIRT* __restrict__ out) {
  pdl_prologue();
  const long long n = (long long)L.width * L.height;
  const bool big = L.is_bigendian != 0;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const unsigned int row = (unsigned int)(i / L.width), col = (unsigned int)(i % L.width);
    const unsigned char* p = raw + (size_t)row * L.row_step + (size_t)col * L.point_step;
    const float x = L.off_x >= 0 ? __uint_as_float(load_u32(p + L.off_x, big)) : 0.f;
    const float y = L.off_y >= 0 ? __uint_as_float(load_u32(p + L.off_y, big)) : 0.f;
    const float z = L.off_z >= 0 ? __uint_as_float(load_u32(p + L.off_z, big)) : 0.f;
    const float it = L.off_intensity >= 0 ? __uint_as_float(load_u32(p + L.off_intensity, big)) : 0.f;
    const unsigned int ring = L.off_ring >= 0 ? load_u16(p + L.off_ring, big) : 0u;
    const float t = L.off_time >= 0 ? __uint_as_float(load_u32(p + L.off_time, big)) : 0.f;
    // two 16-byte stores: (x, y, z, 0) and (intensity, ring | pad, time, 0)
    float4* o = reinterpret_cast<float4*>(out + i);
    o[0] = make_float4(x, y, z, 0.f);
    o[1] = make_float4(it, __uint_as_float(ring), t, 0.f);
  }
}

}  // namespace

void unpack_pointcloud2_device(const unsigned char* d_raw, const floam_pc2_layout& layout, PointIRT* d_out, cudaStream_t s) {
  const long long n = (long long)layout.width * layout.height;
  long long g = (n + kThreads - 1) / kThreads;
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  if (g < 1) g = 1;
  FLOAM_LAUNCH(K_UNPACK_PC2, unpack_pc2_kernel, (int)g, kThreads, s, d_raw, layout, d_out);
}

}  // namespace floam
