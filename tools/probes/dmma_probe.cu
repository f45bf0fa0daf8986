// Does mma.sync m8n8k4 f64 (DMMA) assemble for sm_100a, and what does a Gram-matrix accumulation look like?  (development probe)
// G = sum over rows of v^T v for 8-wide rows: lane (g = lane >> 2, t = lane & 3) feeds A[g][t] = B[t][g] = row_t[g].
#include <cstdio>
__global__ void gram(const double* __restrict__ rows, int n, double* __restrict__ out) {
  const int l = threadIdx.x & 31, g = l >> 2, t = l & 3;
  double c0 = 0.0, c1 = 0.0;
  for (int base = 0; base < n; base += 4) {
    const double v = base + t < n ? rows[(base + t) * 8 + g] : 0.0;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(v), "d"(v));
  }
  // C fragment: lane holds C[g][2t], C[g][2t+1]
  out[g * 8 + 2 * t] = c0; out[g * 8 + 2 * t + 1] = c1;
}
__global__ void latency(double* out, long long* cyc) {
  const int l = threadIdx.x & 31;
  double c0 = 0.0, c1 = 0.0, v = 1e-3 * l;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(v), "d"(v));
  long long t1 = clock64();
  double s = c0;
#pragma unroll
  for (int i = 0; i < 64; ++i) s = fma(s, 1.0000001, v);     // dependent DFMA chain for comparison
  long long t2 = clock64();
  double w = s;
#pragma unroll
  for (int i = 0; i < 64; ++i) w += __shfl_xor_sync(0xffffffffu, w, 1 + (i & 15));   // dependent 64-bit shuffle + add chain
  long long t3 = clock64();
  out[threadIdx.x] = c0 + c1 + s + w;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
}
int main() {
  const int n = 37;
  double h[n * 8], *d, *o, r[64], ref[64] = {0};
  for (int i = 0; i < n * 8; ++i) h[i] = (i % 8 == 7) ? 0.0 : 0.01 * ((i * 37) % 101) - 0.3;
  for (int i = 0; i < n; ++i) for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) ref[a * 8 + b] += h[i * 8 + a] * h[i * 8 + b];
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, sizeof(r));
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  gram<<<1, 32>>>(d, n, o);
  cudaMemcpy(r, o, sizeof(r), cudaMemcpyDeviceToHost);
  double e = 0; for (int i = 0; i < 64; ++i) e = fmax(e, fabs(r[i] - ref[i]));
  long long* dc; long long hc[3]; cudaMalloc(&dc, 24);
  for (int warps = 1; warps <= 8; warps *= 8) {
    latency<<<1, 32 * warps>>>(o, dc); latency<<<1, 32 * warps>>>(o, dc);
    cudaMemcpy(hc, dc, 24, cudaMemcpyDeviceToHost);
    printf("%d warp(s): 64 dependent DMMA %lld cycles (%.1f each), 64 dependent DFMA %lld (%.1f), 64 shuffle+add %lld (%.1f)\n", warps, hc[0], hc[0] / 64.0, hc[1], hc[1] / 64.0, hc[2], hc[2] / 64.0);
  }
  printf("dmma gram max abs err %.3e (%s)\n", e, cudaGetErrorString(cudaGetLastError()));
  return e < 1e-12 ? 0 : 1;
}
