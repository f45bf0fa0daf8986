// Development probe: conditional (IF) graph nodes inserted into a stream capture.  Measures what a skipped body costs compared
// with the same kernels launched unconditionally and exiting early on a device flag (the d_skip scheme of the keyframe map update).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cond_graph_probe cond_graph_probe.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void decide(const int* flag, cudaGraphConditionalHandle h, int* trace) { if (threadIdx.x == 0) { cudaGraphSetConditional(h, *flag ? 1u : 0u); atomicAdd(trace, 1); } }
__global__ void body(const int* skip, int* trace, int v) { if (skip && *skip) return; if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(trace, v); }
__global__ void tail(int* trace) { if (threadIdx.x == 0) atomicAdd(trace, 1000000); }

int main() {
  int *d_flag, *d_trace, *d_skip;
  CK(cudaMalloc(&d_flag, 4)); CK(cudaMalloc(&d_trace, 4)); CK(cudaMalloc(&d_skip, 4));
  cudaStream_t s, s2;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  const int NBODY = 12;
  // ---- graph A: decide -> IF { 12 kernels } -> tail, built by stream capture with the conditional node spliced in ----
  cudaGraph_t g = nullptr;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  cudaStreamCaptureStatus st; cudaGraph_t cg; const cudaGraphNode_t* deps; size_t ndeps;
  CK(cudaStreamGetCaptureInfo_v2(s, &st, nullptr, &cg, &deps, &ndeps));
  cudaGraphConditionalHandle h;
  CK(cudaGraphConditionalHandleCreate(&h, cg, 0, cudaGraphCondAssignDefault));
  decide<<<1, 32, 0, s>>>(d_flag, h, d_trace);
  CK(cudaStreamGetCaptureInfo_v2(s, &st, nullptr, &cg, &deps, &ndeps));
  cudaGraphNodeParams p = {};
  p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeIf; p.conditional.size = 1;
  cudaGraphNode_t cnode;
  CK(cudaGraphAddNode(&cnode, cg, deps, ndeps, &p));
  cudaGraph_t bodyg = p.conditional.phGraph_out[0];
  CK(cudaStreamBeginCaptureToGraph(s2, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  for (int k = 0; k < NBODY; ++k) body<<<148, 256, 0, s2>>>(nullptr, d_trace, 1);
  CK(cudaStreamEndCapture(s2, nullptr));
  CK(cudaStreamUpdateCaptureDependencies(s, &cnode, 1, cudaStreamSetCaptureDependencies));
  tail<<<1, 32, 0, s>>>(d_trace);
  CK(cudaStreamEndCapture(s, &g));
  cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
  // ---- graph B: the same with early-exit kernels ----
  cudaGraph_t g2; CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  body<<<1, 32, 0, s>>>(nullptr, d_trace, 1);
  for (int k = 0; k < NBODY; ++k) body<<<148, 256, 0, s>>>(d_skip, d_trace, 1);
  tail<<<1, 32, 0, s>>>(d_trace);
  CK(cudaStreamEndCapture(s, &g2));
  cudaGraphExec_t ex2; CK(cudaGraphInstantiate(&ex2, g2, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int flag = 0; flag < 2; ++flag) {
    int skip = flag ? 0 : 1;
    CK(cudaMemcpy(d_flag, &flag, 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_skip, &skip, 4, cudaMemcpyHostToDevice));
    for (int which = 0; which < 2; ++which) {
      CK(cudaMemset(d_trace, 0, 4));
      cudaGraphExec_t x = which ? ex2 : ex;
      for (int i = 0; i < 20; ++i) CK(cudaGraphLaunch(x, s));
      CK(cudaStreamSynchronize(s));
      CK(cudaMemset(d_trace, 0, 4));
      CK(cudaEventRecord(e0, s));
      const int R = 200;
      for (int i = 0; i < R; ++i) CK(cudaGraphLaunch(x, s));
      CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      int tr; CK(cudaMemcpy(&tr, d_trace, 4, cudaMemcpyDeviceToHost));
      printf("%s body %s: %.2f us per graph launch, trace %d (per launch: tail %d, others %d)\n", which ? "early-exit kernels" : "IF node", flag ? "taken" : "skipped",
             ms * 1e3 / R, tr, tr / 1000000 / R, (tr % 1000000) / R);
    }
  }
  return 0;
}
