// C ABI of the B200-native FLOAM odometry path (include/floam_b200.h). Thin host layer: argument checks, uploads, kernel
// sequencing (captured into CUDA graphs for the per-frame path) and the small pinned mailboxes results come back through.
// There is no CPU fallback anywhere in this file: without a usable sm_100 device floam_create fails.
#include <chrono>
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "context.cuh"
#include "ingest.cuh"
#include "odom_math.cuh"

using namespace floam;

namespace floam {
void* ctx_alloc(void* vctx, size_t bytes) {
  floam_ctx* c = (floam_ctx*)vctx;
  void* p = nullptr;
  if (bytes == 0) bytes = 256;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    std::fprintf(stderr, "[floam_b200] cudaMalloc(%zu) failed: %s\n", bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  c->allocs.push_back(p);
  return p;
}
}  // namespace floam

namespace {

void* host_alloc(floam_ctx* c, size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  c->host_allocs.push_back(p);
  return p;
}

int set_device(floam_ctx* c) {
  FLOAM_CUDA_OK(cudaSetDevice(c->device));
  return FLOAM_OK;
}

int check_async(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    std::fprintf(stderr, "[floam_b200] CUDA error after %s: %s\n", where, cudaGetErrorString(e));
    return FLOAM_ERR_CUDA;
  }
  return FLOAM_OK;
}

void iso12_to_rowmajor16(const double* T, double* M) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) M[r * 4 + c] = T[r * 3 + c];
    M[r * 4 + 3] = T[9 + r];
  }
  M[12] = M[13] = M[14] = 0.0; M[15] = 1.0;
}
void rowmajor16_to_iso12(const double* M, double* T) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[r * 3 + c] = M[r * 4 + c];
    T[9 + r] = M[r * 4 + 3];
  }
}

// upload a host cloud of 32-byte points into a device buffer and its count into a device int
int upload_cloud(floam_ctx* c, const void* host, int n, void* d_buf, int* d_n, int slot_int) {
  if (n > 0) FLOAM_CUDA_OK(cudaMemcpyAsync(d_buf, host, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
  c->h_ints[slot_int] = n;
  FLOAM_CUDA_OK(cudaMemcpyAsync(d_n, &c->h_ints[slot_int], sizeof(int), cudaMemcpyHostToDevice, c->stream));
  return FLOAM_OK;
}

// point the unsuffixed buffer names at the set of frame parity p (host-side aliases read at enqueue time)
void select_parity(floam_ctx* c, int p) {
  c->d_edge = c->d_edge_b[p]; c->d_surf = c->d_surf_b[p];
  c->d_ne = c->d_ne_b[p]; c->d_ns = c->d_ns_b[p];
  c->d_edge_src = c->d_edge_src_b[p]; c->d_surf_src = c->d_surf_src_b[p];
  c->d_flags = c->d_flags_b[p];
  odom_select_buffers(c->odom, p);
}

int fetch_state(floam_ctx* c, int slot) {
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_state[slot], c->odom.state, sizeof(PoseState), cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_flags[slot], c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  return FLOAM_OK;
}

int status_from_flags(floam_ctx* c, int slot) {
  const int sf = c->h_state[slot]->error_flags, ff = *c->h_flags[slot];
  if (ff & 1) return FLOAM_ERR_NONFINITE;
  if ((ff & 2) || sf) return FLOAM_ERR_CAPACITY;
  return FLOAM_OK;
}

void pose_out_from_state(const PoseState* S, double pose[7]) {
  for (int k = 0; k < 7; ++k) pose[k] = S->x[k];
}

int next_outer(int count) { return count > 2 ? count - 1 : count; }

// Size hints for the keyframe map filters (they only pick a kernel variant, odom_map_update_merges): the local map sizes the last
// mailed update started from.
void note_map_sizes(floam_ctx* c, const PoseState* S) {
  c->odom.edge_map_hint = S->map_points[0];
  c->odom.surf_map_hint = S->map_points[1];
}

// OdomEstimationClass::UpdatePointsToMapSelector (src/odomEstimationClass.cpp:34-50) on device-resident feature clouds
void enqueue_selector(floam_ctx* c, PointIRT* d_edge, const int* d_ne, PointIRT* d_surf, const int* d_ns, int n_max, int deskew, bool ds_ready = false) {
  OdomDevice& od = c->odom;
  if (!deskew) {
    od.optimization_count = next_outer(od.optimization_count);
    odom_update_device(od, d_edge, d_ne, d_surf, d_ns, 32, n_max, FLOAM_VANILLA, ds_ready ? 1 : 0, c->stream);
  } else {
    od.optimization_count = next_outer(od.optimization_count);
    odom_update_device(od, d_edge, d_ne, d_edge, d_ne, 32, n_max, FLOAM_INITIAL_ITERATION, 0, c->stream);  // Q3: edge as both clouds
    compensate_velocity_device(od, d_edge, d_ne, n_max, c->stream);
    compensate_velocity_device(od, d_surf, d_ns, n_max, c->stream);
    od.optimization_count = next_outer(od.optimization_count);
    odom_update_device(od, d_edge, d_ne, d_surf, d_ns, 32, n_max, FLOAM_REFINEMENT_AND_UPDATE, 0, c->stream);
  }
}

// A frame is enqueued as two halves on two stream pairs (context.cuh): FRONT = [IMU deskew] + feature extraction + (no-deskew mode)
// downSamplingToMap — everything that depends on neither the pose nor the map; BACK = prediction, association + solve, write-back,
// keyframe map update. FRONT(k+1) overlaps BACK(k); the buffers the two halves exchange exist twice (frame parity).
void enqueue_front(floam_ctx* c, PointIRT* d_scan, const int* d_scan_n, int deskew, int slot, bool first, bool imu) {
  cudaStream_t s = c->front_stream;
  cudaMemsetAsync(c->d_flags, 0, 4, s);   // error flags are per frame: a NaN point or an over-long ring in one scan does not poison the next
  // CenterTime + Compensate + IMU alignment folded into the frame (src/laserProcessingNode.cpp:100-116): in place on the uploaded scan
  if (imu) deskew_launch(c->imu, c->d_plan[slot], d_scan, d_scan_n, c->prm.max_scan_points, s);
  feature_extract_device(d_scan, d_scan_n, c->fprm, c->fws, c->d_edge, c->d_ne, c->d_surf, c->d_ns, c->d_edge_src, c->d_surf_src, c->d_flags, s);
  if (!first && !deskew)   // the two-pass deskew mode downsamples different clouds in each pass (Q3): it stays inside the BACK
    odom_downsample_device(c->odom, c->d_edge, c->d_ne, c->d_surf, c->d_ns, 32, c->prm.max_scan_points, c->vws_front, c->vws_front_aux, s, c->front_aux,
                           c->ev_ffork, c->ev_fjoin);
}

void enqueue_back(floam_ctx* c, int deskew, int ring, bool first) {
  if (first) {
    // odomEstimationNode.cpp:219-224: first frame only seeds the map (raw features, Q11); odom stays identity
    odom_init_map_device(c->odom, c->d_edge, c->d_ne, c->d_surf, c->d_ns, 32, c->prm.max_scan_points, 0, c->stream);
    odom_record_pose(c->odom, c->stream);
    c->odom.optimization_count = 12;
  } else {
    enqueue_selector(c, c->d_edge, c->d_ne, c->d_surf, c->d_ns, c->prm.max_scan_points, deskew, !deskew);
  }
  if (c->timer.enabled) launch_noop(c->stream);   // calibration of the event-pair overhead (kernel-timing mode only)
  odom_mail_state(c->odom, c->d_flags, c->h_state[ring], c->h_flags[ring], c->stream);
}

// Capture-or-replay of one half. Graphs are keyed by everything that changes the launch sequence or the buffers it touches.
template <class Body>
int launch_half(floam_ctx* c, const floam_graph_key& key, cudaStream_t s, Body body) {
  if (!c->use_graphs) {
    body();
    return check_async("frame");
  }
  const bool timing = c->timer.enabled;
  auto it = c->graphs.find(key);
  if (it == c->graphs.end()) {
    const long long before = g_launches;
    cudaGraph_t graph = nullptr;
    floam_graph_entry e;
    if (timing) { launch_timer_collect(&c->timer, c->stream); e.pair_first = c->timer.used; }
    FLOAM_CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    body();
    FLOAM_CUDA_OK(cudaStreamEndCapture(s, &graph));
    if (timing) { e.pair_last = c->timer.used; c->timer.persist = c->timer.used; }
    FLOAM_CUDA_OK(cudaGraphInstantiate(&e.exec, graph, 0));
    cudaGraphDestroy(graph);
    e.launches = (int)(g_launches - before);
    g_launches = before;                  // the capture launched nothing; replays are counted below
    it = c->graphs.emplace(key, e).first;
  }
  g_launches += it->second.launches;
  FLOAM_CUDA_OK(cudaGraphLaunch(it->second.exec, s));
  if (timing) {  // per-kernel event pairs live inside the graph: read them back after this replay
    FLOAM_CUDA_OK(cudaStreamSynchronize(s));
    launch_timer_fold(&c->timer, it->second.pair_first, it->second.pair_last);
  }
  return FLOAM_OK;
}

// Both halves of the frame whose scan sits in d_scan[slot] (slot = frame parity = ring & 1; ring = the frame's mailbox). The caller has
// made front_stream wait for the scan.
int launch_frame(floam_ctx* c, int deskew, int ring, bool imu) {
  const int slot = ring & 1;
  const bool first = !c->map_initialised;
  OdomDevice& od = c->odom;
  const bool timing = c->timer.enabled;
  select_parity(c, slot);
  const int outer = first ? 0 : next_outer(od.optimization_count);
  const int variant = first ? 0 : (odom_map_update_merges(od, 0) ? 16 : 0) + (odom_map_update_merges(od, 1) ? 32 : 0);   // which map filters the BACK is captured with
  const int flags = (first ? 0 : 1) + (imu ? 2 : 0) + (timing ? 4 : 0);
  // FRONT(k) may not overwrite the parity's buffers before BACK(k-2) has finished with them
  if (c->back_valid[slot]) FLOAM_CUDA_OK(cudaStreamWaitEvent(c->front_stream, c->ev_back_done[slot], 0));
  int rc = launch_half(c, floam_graph_key{flags + 8, 0, first ? 0 : (deskew ? 1 : 0), slot}, c->front_stream,
                       [&]() { enqueue_front(c, c->d_scan[slot], c->d_scan_n[slot], deskew, slot, first, imu); });
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_front_done[slot], c->front_stream));
  FLOAM_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_front_done[slot], 0));
  const int saved_count = od.optimization_count;
  rc = launch_half(c, floam_graph_key{flags + variant, outer, first ? 0 : (deskew ? 1 : 0), ring}, c->stream, [&]() { enqueue_back(c, deskew, ring, first); });
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_back_done[slot], c->stream));
  c->back_valid[slot] = true;
  // host-side transitions of the deterministic outer-iteration schedule (the enqueue applies them only when it is actually run)
  od.optimization_count = saved_count;
  if (first) {
    od.optimization_count = 12;  // initMapWithPoints, src/odomEstimationClass.cpp:31
  } else {
    od.optimization_count = next_outer(od.optimization_count);
    if (deskew) od.optimization_count = next_outer(od.optimization_count);  // second pass (Q4)
  }
  c->map_initialised = true;
  return FLOAM_OK;
}

}  // namespace

extern "C" {

void floam_params_default(floam_params* p) {
  if (!p) return;
  p->num_lines = 64;            // src/laserProcessingNode.cpp:175
  p->scan_period = 0.1;         // :177
  p->vertical_angle = 2.0;      // :176
  p->max_distance = 60.0;       // :178
  p->min_distance = 2.0;        // :179
  p->map_resolution = 0.4;      // src/odomEstimationNode.cpp:328
  p->loss = FLOAM_LOSS_HUBER;   // code default "Huber", :330
  p->max_scan_points = 300000;
  p->max_map_points = 4000000;
  p->max_global_map_points = 8000000;
  p->max_grid_cells = 1 << 23;
  p->fixes = 0;                 // reference behaviour, quirks included
}

int floam_loss_from_string(const char* loss_function) {
  std::string s = loss_function ? loss_function : "";
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char ch) { return (char)std::tolower(ch); });  // src/odomEstimationClass.cpp:23
  if (s == "huber") return FLOAM_LOSS_HUBER;
  if (s == "cauchy_true") return FLOAM_LOSS_CAUCHY_TRUE;  // opt-in extension, not a reference string
  return FLOAM_LOSS_TRIVIAL;                               // "cauchy" and anything else: loss_function stays nullptr (Q1)
}

const char* floam_status_string(int status) {
  switch (status) {
    case FLOAM_OK: return "ok";
    case FLOAM_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required; there is no CPU fallback)";
    case FLOAM_ERR_CUDA: return "CUDA runtime error";
    case FLOAM_ERR_CAPACITY: return "capacity exceeded";
    case FLOAM_ERR_ARG: return "invalid argument";
    case FLOAM_NO_IMU: return "no imu data";
    case FLOAM_ERR_NONFINITE: return "non-finite input";
    default: return "unknown status";
  }
}

const char* floam_version(void) { return "floam_b200 0.1 (sm_100a)"; }

int floam_create(const floam_params* params, int device, floam_ctx** out) {
  if (!params || !out) return FLOAM_ERR_ARG;
  *out = nullptr;
  if (params->num_lines < 1 || params->num_lines > 128 || params->max_scan_points < 1024 || params->max_map_points < 1024 ||
      params->max_grid_cells < 4096 || !(params->map_resolution > 0.0) || !(params->scan_period > 0.0) || (params->fixes & ~7))
    return FLOAM_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return FLOAM_ERR_NO_DEVICE; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) { cudaGetLastError(); return FLOAM_ERR_NO_DEVICE; }
  if (cudaSetDevice(device) != cudaSuccess) return FLOAM_ERR_NO_DEVICE;

  floam_ctx* c = new floam_ctx();
  c->prm = *params;
  c->device = device;
  c->use_graphs = std::getenv("FLOAM_NO_GRAPHS") == nullptr;
  if (const char* e = std::getenv("FLOAM_PDL")) g_use_pdl = std::atoi(e) != 0;   // programmatic dependent launch (common.cuh)
  if (const char* e = std::getenv("FLOAM_PDL_SOLVE")) g_pdl_solve = std::atoi(e) != 0;
  auto fail = [&](int rc) { floam_destroy(c); return rc; };
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(FLOAM_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(FLOAM_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(FLOAM_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&c->front_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(FLOAM_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&c->front_aux, cudaStreamNonBlocking) != cudaSuccess) return fail(FLOAM_ERR_CUDA);
  if (cudaEventCreateWithFlags(&c->ev_ffork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_fjoin, cudaEventDisableTiming) != cudaSuccess)
    return fail(FLOAM_ERR_CUDA);
  for (int k = 0; k < 4; ++k)
    if (cudaEventCreate(&c->ev_begin[k]) != cudaSuccess || cudaEventCreate(&c->ev_end[k]) != cudaSuccess) return fail(FLOAM_ERR_CUDA);
  for (int k = 0; k < 2; ++k) {
    if (cudaEventCreateWithFlags(&c->ev_upload[k], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_front_done[k], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_back_done[k], cudaEventDisableTiming) != cudaSuccess)
      return fail(FLOAM_ERR_CUDA);
  }
  if (cudaEventCreate(&c->ev_replay_begin) != cudaSuccess || cudaEventCreate(&c->ev_replay_end) != cudaSuccess) return fail(FLOAM_ERR_CUDA);
  const int ns = params->max_scan_points;
  const int nm = params->max_map_points;
  c->stage_cap = std::max(std::max(ns, nm), params->max_global_map_points);
  bool ok = true;
  auto A = [&](size_t bytes) { void* p = ctx_alloc(c, bytes); ok = ok && p; return p; };
  for (int k = 0; k < 2; ++k) {
    c->d_scan[k] = (PointIRT*)A((size_t)ns * 32);
    c->d_scan_n[k] = (int*)A(4);
  }
  for (int k = 0; k < 2; ++k) {
    c->d_edge_b[k] = (PointIRT*)A((size_t)ns * 32);
    c->d_surf_b[k] = (PointIRT*)A((size_t)ns * 32);
    c->d_edge_src_b[k] = (int*)A((size_t)ns * 4);
    c->d_surf_src_b[k] = (int*)A((size_t)ns * 4);
  }
  int* ints = (int*)A(128);
  c->d_stage_in = (char*)A((size_t)c->stage_cap * 32);
  c->d_stage_p4 = (P4*)A((size_t)c->stage_cap * 16);
  c->d_stage_out = (P4*)A((size_t)c->stage_cap * 16);
  c->d_stage_bounds = (float*)A(32);
  void* fmem = A(feature_workspace_bytes_padded(ns, params->num_lines));
  void* vmem = A(voxel_workspace_bytes(c->stage_cap));
  const int aux_cap = std::max(ns, nm);
  void* vmem_aux = A(voxel_workspace_bytes(aux_cap));
  void* vmem_front = A(voxel_workspace_bytes(ns));
  void* vmem_front_aux = A(voxel_workspace_bytes(ns));
  c->imu.slerp = (params->fixes & FLOAM_FIX_IMU_SLERP) != 0;
  c->imu.dev_cap = 1 << 20;   // device ring of IMU samples (sliding window, imu.cuh); FLOAM_IMU_RING = smaller power of two for tests
  if (const char* e = std::getenv("FLOAM_IMU_RING")) {
    const int v = std::atoi(e);
    if (v >= 256 && v <= (1 << 24) && (v & (v - 1)) == 0) c->imu.dev_cap = v;
  }
  c->imu.d_samples = (ImuSample*)A((size_t)c->imu.dev_cap * sizeof(ImuSample));
  c->imu.d_plan = (DeskewPlan*)A(sizeof(DeskewPlan));
  for (int k = 0; k < 2; ++k) c->d_plan[k] = (DeskewPlan*)A(sizeof(DeskewPlan));
  if (!ok) return fail(FLOAM_ERR_CUDA);
  c->d_ne_b[0] = ints; c->d_ns_b[0] = ints + 1; c->d_flags_b[0] = ints + 2; c->d_flags_b[1] = ints + 3; c->d_stage_n = ints + 4; c->d_staged_n = ints + 8;
  c->d_ne_b[1] = ints + 16; c->d_ns_b[1] = ints + 17;
  if (cudaMemsetAsync(ints, 0, 128, c->stream) != cudaSuccess) return fail(FLOAM_ERR_CUDA);
  select_parity(c, 0);
  c->fprm.min_distance = params->min_distance;
  c->fprm.max_distance = params->max_distance;
  c->fprm.num_lines = params->num_lines;
  feature_workspace_bind(c->fws, fmem, ns, params->num_lines);
  voxel_workspace_bind(c->vws, vmem, c->stage_cap);
  if (voxel_workspace_arm(c->vws, c->stream)) return fail(FLOAM_ERR_CUDA);
  voxel_workspace_bind(c->vws_aux, vmem_aux, aux_cap);
  if (voxel_workspace_arm(c->vws_aux, c->stream)) return fail(FLOAM_ERR_CUDA);
  voxel_workspace_bind(c->vws_front, vmem_front, ns);
  voxel_workspace_bind(c->vws_front_aux, vmem_front_aux, ns);
  if (voxel_workspace_arm(c->vws_front, c->stream) || voxel_workspace_arm(c->vws_front_aux, c->stream)) return fail(FLOAM_ERR_CUDA);

  c->h_ints = (int*)host_alloc(c, 64 * sizeof(int));
  c->h_doubles = (double*)host_alloc(c, 64 * sizeof(double));
  for (int k = 0; k < 4; ++k) {
    c->h_state[k] = (PoseState*)host_alloc(c, sizeof(PoseState));
    c->h_flags[k] = (int*)host_alloc(c, 64);
    if (!c->h_state[k] || !c->h_flags[k]) return fail(FLOAM_ERR_CUDA);
    std::memset(c->h_state[k], 0, sizeof(PoseState));
    c->h_state[k]->x[3] = 1.0;
    *c->h_flags[k] = 0;
  }
  if (!c->h_ints || !c->h_doubles) return fail(FLOAM_ERR_CUDA);

  int rc = odom_device_init(c->odom, c->prm, &c->vws, &c->vws_aux, c->aux_stream, ctx_alloc, c, c->stream);
  if (rc) return fail(rc);
  if (params->max_global_map_points > 0) {
    rc = mapping_device_init(c->mapping, params->max_global_map_points, params->map_resolution, &c->vws, ctx_alloc, c, c->stream);
    if (rc) return fail(rc);
  }
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return fail(FLOAM_ERR_CUDA);
  c->launches_base = g_launches;
  *out = c;
  return FLOAM_OK;
}

void floam_destroy(floam_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->aux_stream) cudaStreamSynchronize(c->aux_stream);
  if (c->front_stream) cudaStreamSynchronize(c->front_stream);
  if (c->front_aux) cudaStreamSynchronize(c->front_aux);
  if (g_timer == &c->timer) g_timer = nullptr;
  if (c->timer.created) for (int i = 0; i < 2 * LaunchTimer::kPairs; ++i) cudaEventDestroy(c->timer.ev[i]);
  for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second.exec);
  for (void* p : c->allocs) cudaFree(p);
  for (void* p : c->host_allocs) cudaFreeHost(p);
  for (int k = 0; k < 4; ++k) {
    if (c->ev_begin[k]) cudaEventDestroy(c->ev_begin[k]);
    if (c->ev_end[k]) cudaEventDestroy(c->ev_end[k]);
  }
  for (int k = 0; k < 2; ++k) {
    if (c->ev_upload[k]) cudaEventDestroy(c->ev_upload[k]);
    if (c->ev_front_done[k]) cudaEventDestroy(c->ev_front_done[k]);
    if (c->ev_back_done[k]) cudaEventDestroy(c->ev_back_done[k]);
  }
  if (c->ev_replay_begin) cudaEventDestroy(c->ev_replay_begin);
  if (c->ev_replay_end) cudaEventDestroy(c->ev_replay_end);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->front_stream) cudaStreamDestroy(c->front_stream);
  if (c->front_aux) cudaStreamDestroy(c->front_aux);
  if (c->ev_ffork) cudaEventDestroy(c->ev_ffork);
  if (c->ev_fjoin) cudaEventDestroy(c->ev_fjoin);
  if (c->odom.ev_fork) cudaEventDestroy(c->odom.ev_fork);
  if (c->odom.ev_join) cudaEventDestroy(c->odom.ev_join);
  cudaGetLastError();
  delete c;
}

void* floam_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void floam_free_pinned(void* p) { if (p) cudaFreeHost(p); }

// ---- IMU ----------------------------------------------------------------------------------------------------------
int floam_imu_push(floam_ctx* c, double stamp, const double q_xyzw[4]) {
  if (!c || !q_xyzw) return FLOAM_ERR_ARG;
  imu_push(c->imu, stamp, q_xyzw);
  return FLOAM_OK;
}
int floam_imu_get(floam_ctx* c, double stamp, double q_xyzw[4], int* valid) {
  if (!c || !q_xyzw) return FLOAM_ERR_ARG;
  q_xyzw[0] = q_xyzw[1] = q_xyzw[2] = q_xyzw[3] = 0.0;
  const bool ok = imu_get(c->imu, stamp, q_xyzw);
  if (valid) *valid = ok ? 1 : 0;
  return FLOAM_OK;
}
int floam_imu_time_contained(floam_ctx* c, double stamp, int* contained) {
  if (!c || !contained) return FLOAM_ERR_ARG;
  *contained = imu_time_contained(c->imu, stamp) ? 1 : 0;
  return FLOAM_OK;
}
int floam_imu_size(floam_ctx* c, int* n) {
  if (!c || !n) return FLOAM_ERR_ARG;
  *n = (int)c->imu.total();   // every sample ever admitted, like data_.size() of the reference (older ones have left the window)
  return FLOAM_OK;
}

int floam_deskew_align(floam_ctx* c, floam_point_xyzirt* pts, int n, uint64_t* stamp_us, const double extr_xyzw[4]) {
  return floam_deskew_align_ex(c, pts, n, stamp_us, extr_xyzw, FLOAM_DESKEW_CENTER_TIME | FLOAM_DESKEW_COMPENSATE | FLOAM_DESKEW_ALIGN);
}

int floam_compensate_velocity(floam_ctx* c, floam_point_xyzirt* pts, int n, const double velocity[3]) {
  if (!c || (!pts && n > 0) || n < 0 || !velocity) return FLOAM_ERR_ARG;
  if (n > c->prm.max_scan_points) return FLOAM_ERR_CAPACITY;
  if (n == 0) return FLOAM_OK;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  int rc = upload_cloud(c, pts, n, c->d_scan[0], c->d_scan_n[0], 0);
  if (rc) return rc;
  compensate_velocity_explicit_device(c->d_scan[0], c->d_scan_n[0], n, velocity, c->stream);
  FLOAM_CUDA_OK(cudaMemcpyAsync(pts, c->d_scan[0], (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  return check_async("compensate_velocity");
}

int floam_deskew_align_ex(floam_ctx* c, floam_point_xyzirt* pts, int n, uint64_t* stamp_us, const double extr_xyzw[4], int flags) {
  if (!c || !pts || !stamp_us || !extr_xyzw || n < 0) return FLOAM_ERR_ARG;
  if (n > c->prm.max_scan_points) return FLOAM_ERR_CAPACITY;
  if (n == 0) return FLOAM_NO_IMU;  // front()/back() on an empty cloud is undefined in the reference
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  DeskewPlan plan;
  deskew_plan(c->imu, *stamp_us, pts[0].time, pts[n - 1].time, extr_xyzw, flags, &plan);
  int rc = upload_cloud(c, pts, n, c->d_scan[0], c->d_scan_n[0], 0);
  if (rc) return rc;
  rc = deskew_align_device(c->imu, plan, c->d_scan[0], c->d_scan_n[0], n, c->stream);
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaMemcpyAsync(pts, c->d_scan[0], (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  *stamp_us = plan.stamp_us_new;
  return plan.can_compensate ? FLOAM_OK : FLOAM_NO_IMU;
}

// ---- feature extraction ---------------------------------------------------------------------------------------------
int floam_feature_extract(floam_ctx* c, const floam_point_xyzirt* pts, int n, floam_point_xyzirt* edge, int edge_cap, int* ne, floam_point_xyzirt* surf,
                          int surf_cap, int* ns) {
  if (!c || (!pts && n > 0) || !ne || !ns || n < 0) return FLOAM_ERR_ARG;
  if (n > c->prm.max_scan_points) return FLOAM_ERR_CAPACITY;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  FLOAM_CUDA_OK(cudaMemsetAsync(c->d_flags, 0, 4, c->stream));
  int rc = upload_cloud(c, pts, n, c->d_scan[0], c->d_scan_n[0], 0);
  if (rc) return rc;
  feature_extract_device(c->d_scan[0], c->d_scan_n[0], c->fprm, c->fws, c->d_edge, c->d_ne, c->d_surf, c->d_ns, c->d_edge_src, c->d_surf_src, c->d_flags,
                         c->stream);
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_ints + 8, c->d_ne, 12, cudaMemcpyDeviceToHost, c->stream));  // ne, ns, flags
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if ((rc = check_async("feature_extract"))) return rc;
  *ne = c->h_ints[8]; *ns = c->h_ints[9];
  const int flags = c->h_ints[10];
  if (flags & 1) return FLOAM_ERR_NONFINITE;
  if (flags & 2) return FLOAM_ERR_CAPACITY;
  if (*ne > edge_cap || *ns > surf_cap) return FLOAM_ERR_CAPACITY;
  if (edge && *ne) FLOAM_CUDA_OK(cudaMemcpyAsync(edge, c->d_edge, (size_t)*ne * 32, cudaMemcpyDeviceToHost, c->stream));
  if (surf && *ns) FLOAM_CUDA_OK(cudaMemcpyAsync(surf, c->d_surf, (size_t)*ns * 32, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  return FLOAM_OK;
}

// ---- odometry -----------------------------------------------------------------------------------------------------
static int load_maps(floam_ctx* c, const floam_point_xyzi* edge, int ne, const floam_point_xyzi* surf, int ns, int replace) {
  if (!c || ne < 0 || ns < 0 || (ne > 0 && !edge) || (ns > 0 && !surf)) return FLOAM_ERR_ARG;
  if (ne > c->stage_cap || ns > c->stage_cap) return FLOAM_ERR_CAPACITY;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  OdomDevice& od = c->odom;
  int rc = upload_cloud(c, edge, ne, c->d_stage_in, c->d_stage_n, 0);
  if (rc) return rc;
  local_map_load(od, od.edge_map, c->d_stage_in, c->d_stage_n, 32, std::max(ne, 1), replace, c->stream);
  rc = upload_cloud(c, surf, ns, c->d_stage_in, c->d_stage_n + 1, 1);
  if (rc) return rc;
  local_map_load(od, od.surf_map, c->d_stage_in, c->d_stage_n + 1, 32, std::max(ns, 1), replace, c->stream);
  if ((rc = fetch_state(c, 0))) return rc;
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if ((rc = check_async("load_maps"))) return rc;
  c->map_initialised = true;
  return c->h_state[0]->error_flags ? FLOAM_ERR_CAPACITY : FLOAM_OK;
}

int floam_odom_init_map(floam_ctx* c, const floam_point_xyzi* edge, int ne, const floam_point_xyzi* surf, int ns) {
  const int rc = load_maps(c, edge, ne, surf, ns, 0);
  if (rc == FLOAM_OK) c->odom.optimization_count = 12;  // src/odomEstimationClass.cpp:31
  return rc;
}
int floam_odom_set_map(floam_ctx* c, const floam_point_xyzi* edge, int ne, const floam_point_xyzi* surf, int ns) {
  return load_maps(c, edge, ne, surf, ns, 1);
}

static int finish_update(floam_ctx* c, double pose_out[7]) {
  int rc = fetch_state(c, 0);
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_end[0], c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if ((rc = check_async("odom_update"))) return rc;
  cudaEventElapsedTime(&c->last_frame_ms, c->ev_begin[0], c->ev_end[0]);
  note_map_sizes(c, c->h_state[0]);
  if (pose_out) pose_out_from_state(c->h_state[0], pose_out);
  return status_from_flags(c, 0);
}

int floam_odom_update(floam_ctx* c, floam_point_xyzirt* edge, int ne, floam_point_xyzirt* surf, int ns, int deskew, double pose_out[7]) {
  if (!c || ne < 0 || ns < 0 || (ne > 0 && !edge) || (ns > 0 && !surf)) return FLOAM_ERR_ARG;
  if (ne > c->prm.max_scan_points || ns > c->prm.max_scan_points) return FLOAM_ERR_CAPACITY;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_begin[0], c->stream));
  int rc = upload_cloud(c, edge, ne, c->d_edge, c->d_ne, 0);
  if (rc) return rc;
  if ((rc = upload_cloud(c, surf, ns, c->d_surf, c->d_ns, 1))) return rc;
  enqueue_selector(c, c->d_edge, c->d_ne, c->d_surf, c->d_ns, std::max(std::max(ne, ns), 1), deskew);
  if (deskew) {  // the reference compensates the caller's clouds in place (:42-43)
    if (ne) FLOAM_CUDA_OK(cudaMemcpyAsync(edge, c->d_edge, (size_t)ne * 32, cudaMemcpyDeviceToHost, c->stream));
    if (ns) FLOAM_CUDA_OK(cudaMemcpyAsync(surf, c->d_surf, (size_t)ns * 32, cudaMemcpyDeviceToHost, c->stream));
  }
  return finish_update(c, pose_out);
}

int floam_odom_update_xyzi(floam_ctx* c, const floam_point_xyzi* edge, int ne, const floam_point_xyzi* surf, int ns, int update_type, double pose_out[7]) {
  if (!c || ne < 0 || ns < 0 || (ne > 0 && !edge) || (ns > 0 && !surf) || update_type < 0 || update_type > 2) return FLOAM_ERR_ARG;
  if (ne > c->prm.max_scan_points || ns > c->prm.max_scan_points) return FLOAM_ERR_CAPACITY;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_begin[0], c->stream));
  int rc = upload_cloud(c, edge, ne, c->d_edge, c->d_ne, 0);
  if (rc) return rc;
  if ((rc = upload_cloud(c, surf, ns, c->d_surf, c->d_ns, 1))) return rc;
  c->odom.optimization_count = next_outer(c->odom.optimization_count);
  odom_update_device(c->odom, c->d_edge, c->d_ne, c->d_surf, c->d_ns, 32, std::max(std::max(ne, ns), 1), update_type, 0, c->stream);
  return finish_update(c, pose_out);
}

// Refreshes mailbox 0 with the device state.  Mailbox 0 is also a slot of the submit/wait ring, and the pose state belongs to the
// frames in flight: every caller is refused while submissions are pending (like the other stage entry points).
static int sync_state(floam_ctx* c) {
  if (c->inflight != 0) return FLOAM_ERR_ARG;
  if (set_device(c)) return FLOAM_ERR_CUDA;
  int rc = fetch_state(c, 0);
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  note_map_sizes(c, c->h_state[0]);
  return FLOAM_OK;
}

int floam_odom_get(floam_ctx* c, double odom_rowmajor[16], double velocity[3]) {
  if (!c) return FLOAM_ERR_ARG;
  int rc = sync_state(c);
  if (rc) return rc;
  const PoseState* S = c->h_state[0];
  if (odom_rowmajor) iso12_to_rowmajor16(S->odom, odom_rowmajor);
  if (velocity)  // GetVelocity(): (odom.t - last_odom.t) / scan_period, include/odomEstimationClass.h:78
    for (int a = 0; a < 3; ++a) velocity[a] = (S->odom[9 + a] - S->last_odom[9 + a]) / c->prm.scan_period;
  return FLOAM_OK;
}

int floam_odom_get_state(floam_ctx* c, double odom_rowmajor[16], double last_odom_rowmajor[16], int* optimization_count) {
  if (!c) return FLOAM_ERR_ARG;
  int rc = sync_state(c);
  if (rc) return rc;
  if (odom_rowmajor) iso12_to_rowmajor16(c->h_state[0]->odom, odom_rowmajor);
  if (last_odom_rowmajor) iso12_to_rowmajor16(c->h_state[0]->last_odom, last_odom_rowmajor);
  if (optimization_count) *optimization_count = c->odom.optimization_count;
  return FLOAM_OK;
}

int floam_odom_set_state(floam_ctx* c, const double odom_rowmajor[16], const double last_odom_rowmajor[16], int optimization_count) {
  if (!c || !odom_rowmajor || !last_odom_rowmajor || optimization_count < 0) return FLOAM_ERR_ARG;
  int rc = sync_state(c);
  if (rc) return rc;
  PoseState* S = c->h_state[0];
  rowmajor16_to_iso12(odom_rowmajor, S->odom);
  rowmajor16_to_iso12(last_odom_rowmajor, S->last_odom);
  double q[4];
  m::quat_from_matrix(S->odom, q);
  S->x[0] = q[0]; S->x[1] = q[1]; S->x[2] = q[2]; S->x[3] = q[3];
  S->x[4] = S->odom[9]; S->x[5] = S->odom[10]; S->x[6] = S->odom[11];
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->odom.state, S, sizeof(PoseState), cudaMemcpyHostToDevice, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  c->odom.optimization_count = optimization_count;
  return FLOAM_OK;
}

int floam_odom_map_sizes(floam_ctx* c, int* n_edge, int* n_surf) {
  if (!c) return FLOAM_ERR_ARG;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_ints + 16, c->odom.edge_map.d_n, 4, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_ints + 17, c->odom.surf_map.d_n, 4, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if (n_edge) *n_edge = c->h_ints[16];
  if (n_surf) *n_surf = c->h_ints[17];
  return FLOAM_OK;
}

static void p4_to_xyzi(const P4* in, floam_point_xyzi* out, int n) {
  for (int i = 0; i < n; ++i) {
    out[i].x = in[i].x; out[i].y = in[i].y; out[i].z = in[i].z; out[i]._pad0 = 1.0f;
    out[i].intensity = in[i].w; out[i]._pad1[0] = out[i]._pad1[1] = out[i]._pad1[2] = 0.0f;
  }
}

static int download_p4(floam_ctx* c, const P4* d_src, int n, floam_point_xyzi* out) {
  if (n <= 0) return FLOAM_OK;
  std::vector<P4> tmp((size_t)n);
  FLOAM_CUDA_OK(cudaMemcpyAsync(tmp.data(), d_src, (size_t)n * sizeof(P4), cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  p4_to_xyzi(tmp.data(), out, n);
  return FLOAM_OK;
}

int floam_odom_get_map(floam_ctx* c, floam_point_xyzi* edge, int edge_cap, floam_point_xyzi* surf, int surf_cap) {
  int ne = 0, ns = 0;
  int rc = floam_odom_map_sizes(c, &ne, &ns);
  if (rc) return rc;
  if ((edge && ne > edge_cap) || (surf && ns > surf_cap)) return FLOAM_ERR_CAPACITY;
  if (edge && (rc = download_p4(c, c->odom.edge_map.pts, ne, edge))) return rc;
  if (surf && (rc = download_p4(c, c->odom.surf_map.pts, ns, surf))) return rc;
  return FLOAM_OK;
}

// ---- fused frame path ---------------------------------------------------------------------------------------------
static int pc2_layout_check(const floam_pc2_layout* L, long long* n_out) {
  if (!L || L->point_step == 0 || L->width == 0) return FLOAM_ERR_ARG;
  if ((unsigned long long)L->row_step < (unsigned long long)L->width * L->point_step) return FLOAM_ERR_ARG;
  const int32_t offs[6] = {L->off_x, L->off_y, L->off_z, L->off_intensity, L->off_ring, L->off_time};
  for (int k = 0; k < 6; ++k)
    if (offs[k] >= 0 && (uint32_t)offs[k] + (k == 4 ? 2u : 4u) > L->point_step) return FLOAM_ERR_ARG;
  *n_out = (long long)L->width * L->height;
  return FLOAM_OK;
}
static int raw_reserve(floam_ctx* c, size_t bytes) {
  if (bytes <= c->raw_cap) return FLOAM_OK;
  if (c->raw_cap != 0) return FLOAM_ERR_CAPACITY;   // sized by the first message: max_scan_points of its point_step (at least 64 B)
  for (int k = 0; k < 2; ++k)
    if (!(c->d_raw[k] = (unsigned char*)ctx_alloc(c, bytes))) return FLOAM_ERR_CUDA;
  c->raw_cap = bytes;
  return FLOAM_OK;
}

// pts: packed 32-byte points, or (raw, layout): PointCloud2 bytes unpacked on the device
static int submit_common(floam_ctx* c, const floam_point_xyzirt* pts, int n, int deskew, DeskewPlan* plan, const uint8_t* raw = nullptr,
                         const floam_pc2_layout* layout = nullptr) {
  if (!c || (!pts && !raw && n > 0) || n < 0) return FLOAM_ERR_ARG;
  if (n > c->prm.max_scan_points) return FLOAM_ERR_CAPACITY;
  if (c->inflight >= 3) return FLOAM_ERR_ARG;
  if (set_device(c)) return FLOAM_ERR_CUDA;
  const int ring = c->submit_slot, slot = ring & 1;
  // upload on the copy stream once the FRONT that read this scan buffer two frames ago is done
  if (c->consumed_valid[slot]) FLOAM_CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_front_done[slot], 0));
  if (raw && n > 0) {
    const size_t bytes = (size_t)layout->row_step * layout->height;
    const size_t step = layout->point_step > 64 ? layout->point_step : 64;
    int rc = raw_reserve(c, bytes > step * c->prm.max_scan_points ? bytes : step * c->prm.max_scan_points);
    if (rc) return rc;
    FLOAM_CUDA_OK(cudaMemcpyAsync(c->d_raw[slot], raw, bytes, cudaMemcpyHostToDevice, c->copy_stream));
    unpack_pointcloud2_device(c->d_raw[slot], *layout, c->d_scan[slot], c->copy_stream);
  } else if (n > 0) FLOAM_CUDA_OK(cudaMemcpyAsync(c->d_scan[slot], pts, (size_t)n * 32, cudaMemcpyHostToDevice, c->copy_stream));
  c->h_ints[32 + ring] = n;
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->d_scan_n[slot], &c->h_ints[32 + ring], 4, cudaMemcpyHostToDevice, c->copy_stream));
  if (plan) {
    const int rc = deskew_upload(c->imu, *plan, c->d_plan[slot], c->copy_stream);
    if (rc) return rc;
  }
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_upload[slot], c->copy_stream));
  FLOAM_CUDA_OK(cudaStreamWaitEvent(c->front_stream, c->ev_upload[slot], 0));
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_begin[ring], c->front_stream));
  int rc = launch_frame(c, deskew, ring, plan != nullptr);
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_end[ring], c->stream));
  c->consumed_valid[slot] = true;
  c->submit_slot = (ring + 1) & 3;
  c->inflight++;
  return FLOAM_OK;
}

// same for a scan that already sits in device memory (staged replay): the copy into the frame's scan buffer rides on the front stream
static int enqueue_staged_frame(floam_ctx* c, int frame, int deskew) {
  const int ring = c->submit_slot, slot = ring & 1;
  const int n = (int)(c->staged_offsets[frame + 1] - c->staged_offsets[frame]);
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_begin[ring], c->front_stream));
  if (n > 0)
    FLOAM_CUDA_OK(cudaMemcpyAsync(c->d_scan[slot], c->d_staged + c->staged_offsets[frame], (size_t)n * 32, cudaMemcpyDeviceToDevice, c->front_stream));
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->d_scan_n[slot], c->d_staged_counts + frame, 4, cudaMemcpyDeviceToDevice, c->front_stream));
  const int rc = launch_frame(c, deskew, ring, false);
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_end[ring], c->stream));
  c->consumed_valid[slot] = true;
  c->submit_slot = (ring + 1) & 3;
  c->wait_slot = c->submit_slot;   // nothing stays in flight across these calls: keep the submit / wait slots aligned
  return FLOAM_OK;
}

int floam_process_submit(floam_ctx* c, const floam_point_xyzirt* pts, int n, int deskew) { return submit_common(c, pts, n, deskew, nullptr); }

int floam_process_submit_imu(floam_ctx* c, const floam_point_xyzirt* pts, int n, uint64_t* stamp_us, const double extr_xyzw[4], int deskew) {
  if (!c || !pts || n < 1 || !stamp_us || !extr_xyzw) return FLOAM_ERR_ARG;
  DeskewPlan plan;
  deskew_plan(c->imu, *stamp_us, pts[0].time, pts[n - 1].time, extr_xyzw, FLOAM_DESKEW_CENTER_TIME | FLOAM_DESKEW_COMPENSATE | FLOAM_DESKEW_ALIGN, &plan);
  *stamp_us = plan.stamp_us_new;
  if (!plan.can_compensate) return FLOAM_NO_IMU;  // "cannot compensate - no IMU data": the node drops the scan (src/laserProcessingNode.cpp:108-112)
  return submit_common(c, pts, n, deskew, &plan);
}

// the time field of point i of a raw message, read on the host (the deskew plan needs the first and the last point's)
static float pc2_time_at(const uint8_t* data, const floam_pc2_layout* L, long long i) {
  if (L->off_time < 0) return 0.f;
  const uint8_t* p = data + (size_t)(i / L->width) * L->row_step + (size_t)(i % L->width) * L->point_step + L->off_time;
  uint32_t u = L->is_bigendian ? ((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3])
                               : ((uint32_t)p[3] << 24 | (uint32_t)p[2] << 16 | (uint32_t)p[1] << 8 | p[0]);
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

int floam_process_submit_pc2(floam_ctx* c, const uint8_t* data, const floam_pc2_layout* layout, uint64_t* stamp_us, const double extr_xyzw[4], int deskew) {
  long long n = 0;
  if (!c || !data || (stamp_us == nullptr) != (extr_xyzw == nullptr)) return FLOAM_ERR_ARG;
  int rc = pc2_layout_check(layout, &n);
  if (rc) return rc;
  if (n > c->prm.max_scan_points) return FLOAM_ERR_CAPACITY;
  if (n < 1) return FLOAM_ERR_ARG;
  if (!stamp_us) return submit_common(c, nullptr, (int)n, deskew, nullptr, data, layout);
  DeskewPlan plan;
  deskew_plan(c->imu, *stamp_us, pc2_time_at(data, layout, 0), pc2_time_at(data, layout, n - 1), extr_xyzw,
              FLOAM_DESKEW_CENTER_TIME | FLOAM_DESKEW_COMPENSATE | FLOAM_DESKEW_ALIGN, &plan);
  *stamp_us = plan.stamp_us_new;
  if (!plan.can_compensate) return FLOAM_NO_IMU;
  return submit_common(c, nullptr, (int)n, deskew, &plan, data, layout);
}

int floam_unpack_pointcloud2(floam_ctx* c, const uint8_t* data, const floam_pc2_layout* layout, floam_point_xyzirt* out) {
  long long n = 0;
  if (!c || !data || !out || c->inflight != 0) return FLOAM_ERR_ARG;
  int rc = pc2_layout_check(layout, &n);
  if (rc) return rc;
  if (n > c->prm.max_scan_points) return FLOAM_ERR_CAPACITY;
  if (n < 1) return FLOAM_OK;
  if (set_device(c)) return FLOAM_ERR_CUDA;
  const size_t bytes = (size_t)layout->row_step * layout->height;
  const size_t step = layout->point_step > 64 ? layout->point_step : 64;
  if ((rc = raw_reserve(c, bytes > step * c->prm.max_scan_points ? bytes : step * c->prm.max_scan_points))) return rc;
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->d_raw[0], data, bytes, cudaMemcpyHostToDevice, c->stream));
  unpack_pointcloud2_device(c->d_raw[0], *layout, c->d_scan[0], c->stream);
  FLOAM_CUDA_OK(cudaMemcpyAsync(out, c->d_scan[0], (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  return check_async("unpack_pointcloud2");
}

int floam_process_scan_imu(floam_ctx* c, const floam_point_xyzirt* pts, int n, uint64_t* stamp_us, const double extr_xyzw[4], int deskew, double pose_out[7]) {
  if (!c || c->inflight != 0) return FLOAM_ERR_ARG;
  const int rc = floam_process_submit_imu(c, pts, n, stamp_us, extr_xyzw, deskew);
  if (rc) return rc;
  return floam_process_wait(c, pose_out);
}

int floam_process_wait(floam_ctx* c, double pose_out[7]) {
  if (!c || c->inflight <= 0) return FLOAM_ERR_ARG;
  if (set_device(c)) return FLOAM_ERR_CUDA;
  const int slot = c->wait_slot;
  FLOAM_CUDA_OK(cudaEventSynchronize(c->ev_end[slot]));
  cudaEventElapsedTime(&c->last_frame_ms, c->ev_begin[slot], c->ev_end[slot]);
  c->wait_slot = (slot + 1) & 3;
  c->inflight--;
  note_map_sizes(c, c->h_state[slot]);
  if (pose_out) pose_out_from_state(c->h_state[slot], pose_out);
  return status_from_flags(c, slot);
}

int floam_process_scan(floam_ctx* c, const floam_point_xyzirt* pts, int n, int deskew, double pose_out[7]) {
  if (!c || c->inflight != 0) return FLOAM_ERR_ARG;
  const int rc = floam_process_submit(c, pts, n, deskew);
  if (rc) return rc;
  return floam_process_wait(c, pose_out);
}

int floam_stage_scans(floam_ctx* c, const floam_point_xyzirt* pts, const int64_t* offsets, int n_frames) {
  if (!c || !pts || !offsets || n_frames < 1) return FLOAM_ERR_ARG;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  for (int f = 0; f < n_frames; ++f)
    if (offsets[f + 1] < offsets[f] || offsets[f + 1] - offsets[f] > c->prm.max_scan_points) return FLOAM_ERR_CAPACITY;
  const size_t total = (size_t)(offsets[n_frames] - offsets[0]);
  void* p = ctx_alloc(c, (total + 1) * 32);
  int* counts = (int*)ctx_alloc(c, (size_t)n_frames * 4);
  if (!p || !counts) return FLOAM_ERR_CUDA;
  c->d_staged = (PointIRT*)p;
  c->d_staged_counts = counts;
  c->staged_offsets.assign(n_frames + 1, 0);
  std::vector<int> hc(n_frames);
  for (int f = 0; f <= n_frames; ++f) c->staged_offsets[f] = (long long)(offsets[f] - offsets[0]);
  for (int f = 0; f < n_frames; ++f) hc[f] = (int)(offsets[f + 1] - offsets[f]);
  FLOAM_CUDA_OK(cudaMemcpy(p, pts + offsets[0], total * 32, cudaMemcpyHostToDevice));
  FLOAM_CUDA_OK(cudaMemcpy(counts, hc.data(), (size_t)n_frames * 4, cudaMemcpyHostToDevice));
  return FLOAM_OK;
}

int floam_process_staged(floam_ctx* c, int frame, int deskew, double pose_out[7]) {
  if (!c || !c->d_staged || frame < 0 || frame + 1 >= (int)c->staged_offsets.size() || c->inflight != 0) return FLOAM_ERR_ARG;
  if (set_device(c)) return FLOAM_ERR_CUDA;
  const int slot = c->submit_slot;
  int rc = enqueue_staged_frame(c, frame, deskew);
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if ((rc = check_async("process_staged"))) return rc;
  cudaEventElapsedTime(&c->last_frame_ms, c->ev_begin[slot], c->ev_end[slot]);
  note_map_sizes(c, c->h_state[slot]);
  if (pose_out) pose_out_from_state(c->h_state[slot], pose_out);
  return status_from_flags(c, slot);
}

// ---- stage entry points: pcl::VoxelGrid, pcl::CropBox, pcl::KdTreeFLANN ------------------------------------------------
int floam_voxel_grid(floam_ctx* c, const floam_point_xyzi* pts, int n, float leaf, floam_point_xyzi* out, int cap, int* n_out) {
  if (!c || (!pts && n > 0) || !n_out || n < 0 || !(leaf > 0.f)) return FLOAM_ERR_ARG;
  if (n > c->stage_cap) return FLOAM_ERR_CAPACITY;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  int rc = upload_cloud(c, pts, n, c->d_stage_in, c->d_stage_n, 0);
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaMemsetAsync(c->d_stage_n + 1, 0, 4, c->stream));
  voxel_grid_device(c->d_stage_in, 32, c->d_stage_n, std::max(n, 1), leaf, c->d_stage_out, c->d_stage_n + 1, c->vws, nullptr, c->stream);
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_ints + 20, c->d_stage_n + 1, 4, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if ((rc = check_async("voxel_grid"))) return rc;
  *n_out = c->h_ints[20];
  if (*n_out > cap) return FLOAM_ERR_CAPACITY;
  return out ? download_p4(c, c->d_stage_out, *n_out, out) : FLOAM_OK;
}

int floam_voxel_grid_update(floam_ctx* c, const floam_point_xyzi* map_pts, int n_map, const floam_point_xyzi* new_pts, int n_new, float leaf,
                            const float min_xyz[3], const float max_xyz[3], floam_point_xyzi* out, int cap, int* n_out) {
  if (!c || (!map_pts && n_map > 0) || (!new_pts && n_new > 0) || !n_out || n_map < 0 || n_new < 0 || !(leaf > 0.f) || (min_xyz == nullptr) != (max_xyz == nullptr))
    return FLOAM_ERR_ARG;
  if ((long long)n_map + n_new > c->stage_cap) return FLOAM_ERR_CAPACITY;
  if (c->inflight != 0) return FLOAM_ERR_ARG;
  if (set_device(c)) return FLOAM_ERR_CUDA;
  int rc = upload_cloud(c, map_pts, n_map, c->d_stage_in, c->d_stage_n, 0);
  if (rc) return rc;
  if (n_new > 0) FLOAM_CUDA_OK(cudaMemcpyAsync((char*)c->d_stage_in + (size_t)n_map * 32, new_pts, (size_t)n_new * 32, cudaMemcpyHostToDevice, c->stream));
  c->h_ints[1] = n_new;
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->d_stage_n + 2, &c->h_ints[1], 4, cudaMemcpyHostToDevice, c->stream));
  if (min_xyz) {
    float* hb = (float*)(c->h_doubles);
    for (int a = 0; a < 3; ++a) { hb[a] = min_xyz[a]; hb[3 + a] = max_xyz[a]; }
    FLOAM_CUDA_OK(cudaMemcpyAsync(c->d_stage_bounds, hb, 24, cudaMemcpyHostToDevice, c->stream));
  }
  FLOAM_CUDA_OK(cudaMemsetAsync(c->d_stage_n + 1, 0, 4, c->stream));
  voxel_grid_merge_device(c->d_stage_in, 32, c->d_stage_n, std::max(n_map + n_new, 1), leaf, c->d_stage_out, c->d_stage_n + 1, c->vws, nullptr, c->stream,
                          min_xyz ? c->d_stage_bounds : nullptr, c->d_stage_n + 2, c->stage_cap);
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_ints + 20, c->d_stage_n + 1, 4, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if ((rc = check_async("voxel_grid_update"))) return rc;
  *n_out = c->h_ints[20];
  if (*n_out > cap) return FLOAM_ERR_CAPACITY;
  return out ? download_p4(c, c->d_stage_out, *n_out, out) : FLOAM_OK;
}

int floam_crop_box(floam_ctx* c, const floam_point_xyzi* pts, int n, const float min_xyz[3], const float max_xyz[3], floam_point_xyzi* out, int cap, int* n_out) {
  if (!c || (!pts && n > 0) || !n_out || n < 0 || !min_xyz || !max_xyz) return FLOAM_ERR_ARG;
  if (n > c->stage_cap) return FLOAM_ERR_CAPACITY;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  int rc = upload_cloud(c, pts, n, c->d_stage_in, c->d_stage_n, 0);
  if (rc) return rc;
  float* hb = (float*)(c->h_doubles);
  for (int a = 0; a < 3; ++a) { hb[a] = min_xyz[a]; hb[3 + a] = max_xyz[a]; }
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->d_stage_bounds, hb, 24, cudaMemcpyHostToDevice, c->stream));
  FLOAM_CUDA_OK(cudaMemsetAsync(c->d_stage_n + 1, 0, 4, c->stream));
  repack_xyzi_device(c->d_stage_in, c->d_stage_n, std::max(n, 1), c->d_stage_p4, c->stream);
  crop_box_device(c->d_stage_p4, c->d_stage_n, std::max(n, 1), c->d_stage_bounds, c->d_stage_out, c->d_stage_n + 1, c->vws, nullptr, c->stream);
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_ints + 20, c->d_stage_n + 1, 4, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if ((rc = check_async("crop_box"))) return rc;
  *n_out = c->h_ints[20];
  if (*n_out > cap) return FLOAM_ERR_CAPACITY;
  return out ? download_p4(c, c->d_stage_out, *n_out, out) : FLOAM_OK;
}

int floam_knn5(floam_ctx* c, const floam_point_xyzi* map, int m, const floam_point_xyzi* queries, int nq, int* ids, float* sqdist) {
  if (!c || !map || !queries || m < 0 || nq < 0 || !ids || !sqdist) return FLOAM_ERR_ARG;
  if (m > c->stage_cap || nq > c->stage_cap) return FLOAM_ERR_CAPACITY;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  if (!c->knn_map_ready) {
    int rc = local_map_alloc(c->knn_map, c->prm.max_map_points, c->prm.max_grid_cells, ctx_alloc, c, c->stream);
    if (rc) return rc;
    c->d_knn_ids = (int*)ctx_alloc(c, (size_t)c->stage_cap * 5 * 4);
    c->d_knn_d2 = (float*)ctx_alloc(c, (size_t)c->stage_cap * 5 * 4);
    if (!c->d_knn_ids || !c->d_knn_d2) return FLOAM_ERR_CUDA;
    c->knn_map_ready = true;
  }
  if (m > c->knn_map.cap) return FLOAM_ERR_CAPACITY;
  int rc = upload_cloud(c, map, m, c->d_stage_in, c->d_stage_n, 0);
  if (rc) return rc;
  local_map_load(c->odom, c->knn_map, c->d_stage_in, c->d_stage_n, 32, std::max(m, 1), 1, c->stream);
  if ((rc = upload_cloud(c, queries, nq, c->d_stage_in, c->d_stage_n + 1, 1))) return rc;
  repack_xyzi_device(c->d_stage_in, c->d_stage_n + 1, std::max(nq, 1), c->d_stage_p4, c->stream);
  knn5_device(c->odom, c->knn_map, c->d_stage_p4, c->d_stage_n + 1, std::max(nq, 1), c->d_knn_ids, c->d_knn_d2, c->stream);
  if (nq) {
    FLOAM_CUDA_OK(cudaMemcpyAsync(ids, c->d_knn_ids, (size_t)nq * 5 * 4, cudaMemcpyDeviceToHost, c->stream));
    FLOAM_CUDA_OK(cudaMemcpyAsync(sqdist, c->d_knn_d2, (size_t)nq * 5 * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  if ((rc = fetch_state(c, 0))) return rc;
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if ((rc = check_async("knn5"))) return rc;
  return c->h_state[0]->error_flags ? FLOAM_ERR_CAPACITY : FLOAM_OK;
}

// ---- LaserMappingClass --------------------------------------------------------------------------------------------
int floam_mapping_update(floam_ctx* c, const floam_point_xyzi* pts, int n, const double pose_rowmajor[16]) {
  if (!c || (!pts && n > 0) || n < 0 || !pose_rowmajor) return FLOAM_ERR_ARG;
  if (!c->mapping.enabled) return FLOAM_ERR_ARG;
  if (n > c->stage_cap) return FLOAM_ERR_CAPACITY;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  int rc = upload_cloud(c, pts, n, c->d_stage_in, c->d_stage_n, 0);
  if (rc) return rc;
  if ((rc = mapping_update_device(c->mapping, c->d_stage_in, 32, c->d_stage_n, std::max(n, 1), pose_rowmajor, c->stream))) return rc;
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_ints + 24, c->mapping.d_counts, 32, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if ((rc = check_async("mapping_update"))) return rc;
  return c->h_ints[24 + 7] ? FLOAM_ERR_CAPACITY : FLOAM_OK;
}

int floam_mapping_get_map(floam_ctx* c, floam_point_xyzi* out, int cap, int* n) {
  if (!c || !n) return FLOAM_ERR_ARG;
  if (!c->mapping.enabled) return FLOAM_ERR_ARG;
  if (c->inflight != 0) return FLOAM_ERR_ARG;  // frames submitted with floam_process_submit must be waited for first
  if (set_device(c)) return FLOAM_ERR_CUDA;
  P4* d_out = nullptr;
  int* d_n = nullptr;
  int rc = mapping_get_map_device(c->mapping, &d_out, &d_n, c->stream);
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_ints + 24, d_n, 4, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if ((rc = check_async("mapping_get_map"))) return rc;
  *n = c->h_ints[24];
  if (!out) return FLOAM_OK;
  if (*n > cap) return FLOAM_ERR_CAPACITY;
  return download_p4(c, d_out, *n, out);
}

int floam_mapping_get_changed_cells(floam_ctx* c, floam_point_xyzi* out, int32_t* cells_xyz, int cap, int* n) {
  if (!c || !n || (out && !cells_xyz)) return FLOAM_ERR_ARG;
  if (!c->mapping.enabled) return FLOAM_ERR_ARG;
  if (c->inflight != 0) return FLOAM_ERR_ARG;
  if (set_device(c)) return FLOAM_ERR_CUDA;
  P4* d_out = nullptr; unsigned int* d_cell = nullptr; int* d_n = nullptr;
  int rc = mapping_get_dirty_device(c->mapping, &d_out, &d_cell, &d_n, c->stream);
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaMemcpyAsync(c->h_ints + 24, d_n, 4, cudaMemcpyDeviceToHost, c->stream));
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if ((rc = check_async("mapping_get_changed_cells"))) return rc;
  const int need = c->h_ints[24];
  *n = need;
  if (!out) return FLOAM_OK;                  // size query: the change marks stay
  if (need > cap) return FLOAM_ERR_CAPACITY;  // likewise
  if (need > 0) {
    std::vector<unsigned int> packed((size_t)need);
    FLOAM_CUDA_OK(cudaMemcpy(packed.data(), d_cell, (size_t)need * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < need; ++i) {
      cells_xyz[3 * i] = (int)(packed[i] >> 20) - 512;
      cells_xyz[3 * i + 1] = (int)((packed[i] >> 10) & 1023u) - 512;
      cells_xyz[3 * i + 2] = (int)(packed[i] & 1023u) - 512;
    }
    if ((rc = download_p4(c, d_out, need, out))) return rc;
  }
  if ((rc = mapping_clear_dirty_device(c->mapping, c->stream))) return rc;
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  return FLOAM_OK;
}

// ---- debug taps ---------------------------------------------------------------------------------------------------
int floam_debug_fetch(floam_ctx* c, int what, void* out, size_t cap_bytes, size_t* n_bytes) {
  if (!c || !n_bytes) return FLOAM_ERR_ARG;
  if (set_device(c)) return FLOAM_ERR_CUDA;
  OdomDevice& od = c->odom;
  int rc = sync_state(c);
  if (rc) return rc;
  FLOAM_CUDA_OK(cudaMemcpy(c->h_ints + 40, od.d_nds_edge, 8, cudaMemcpyDeviceToHost));
  FLOAM_CUDA_OK(cudaMemcpy(c->h_ints + 42, c->d_ne, 8, cudaMemcpyDeviceToHost));
  const int nde = c->h_ints[40], nds = c->h_ints[41], ne = c->h_ints[42], ns = c->h_ints[43];
  const PoseState* S = c->h_state[0];
  auto copy_dev = [&](const void* d_src, size_t bytes) -> int {
    *n_bytes = bytes;
    if (!out) return FLOAM_OK;
    if (bytes > cap_bytes) return FLOAM_ERR_CAPACITY;
    if (bytes) FLOAM_CUDA_OK(cudaMemcpy(out, d_src, bytes, cudaMemcpyDeviceToHost));
    return FLOAM_OK;
  };
  switch (what) {
    case FLOAM_DBG_DS_EDGE:
    case FLOAM_DBG_DS_SURF: {
      const int n = what == FLOAM_DBG_DS_EDGE ? nde : nds;
      *n_bytes = (size_t)n * 32;
      if (!out) return FLOAM_OK;
      if (*n_bytes > cap_bytes) return FLOAM_ERR_CAPACITY;
      return download_p4(c, what == FLOAM_DBG_DS_EDGE ? od.ds_edge : od.ds_surf, n, (floam_point_xyzi*)out);
    }
    case FLOAM_DBG_EDGE_KNN: return copy_dev(od.knn_ids, (size_t)nde * 5 * 4);
    case FLOAM_DBG_SURF_KNN: return copy_dev(od.knn_ids + (size_t)od.qcap * 5, (size_t)nds * 5 * 4);
    case FLOAM_DBG_EDGE_D2: return copy_dev(od.knn_d2, (size_t)nde * 5 * 4);
    case FLOAM_DBG_SURF_D2: return copy_dev(od.knn_d2 + (size_t)od.qcap * 5, (size_t)nds * 5 * 4);
    case FLOAM_DBG_EDGE_OK: return copy_dev(od.corr_ok, (size_t)nde);
    case FLOAM_DBG_SURF_OK: return copy_dev(od.corr_ok + od.qcap, (size_t)nds);
    case FLOAM_DBG_RESIDUALS: {
      // records of the accepted correspondences of the last outer iteration: kind, curr(3), a(3), b(3); edge first
      std::vector<unsigned char> ok((size_t)2 * od.qcap);
      std::vector<P4> de((size_t)std::max(nde, 1)), dsf((size_t)std::max(nds, 1));
      std::vector<double> corr((size_t)12 * od.qcap);
      FLOAM_CUDA_OK(cudaMemcpy(ok.data(), od.corr_ok, ok.size(), cudaMemcpyDeviceToHost));
      FLOAM_CUDA_OK(cudaMemcpy(corr.data(), od.corr, corr.size() * 8, cudaMemcpyDeviceToHost));
      if (nde) FLOAM_CUDA_OK(cudaMemcpy(de.data(), od.ds_edge, (size_t)nde * 16, cudaMemcpyDeviceToHost));
      if (nds) FLOAM_CUDA_OK(cudaMemcpy(dsf.data(), od.ds_surf, (size_t)nds * 16, cudaMemcpyDeviceToHost));
      std::vector<double> rec;
      const size_t cs = (size_t)2 * od.qcap;
      if (!S->skip_solve) {
        for (int i = 0; i < nde; ++i)
          if (ok[i]) {
            const double r[10] = {0.0, de[i].x, de[i].y, de[i].z, corr[0 * cs + i], corr[1 * cs + i], corr[2 * cs + i],
                                  corr[3 * cs + i], corr[4 * cs + i], corr[5 * cs + i]};
            rec.insert(rec.end(), r, r + 10);
          }
        for (int i = 0; i < nds; ++i) {
          const size_t o = (size_t)od.qcap + i;
          if (ok[o]) {
            const double r[10] = {1.0, dsf[i].x, dsf[i].y, dsf[i].z, corr[0 * cs + o], corr[1 * cs + o], corr[2 * cs + o], corr[3 * cs + o], 0.0, 0.0};
            rec.insert(rec.end(), r, r + 10);
          }
        }
      }
      *n_bytes = rec.size() * 8;
      if (!out) return FLOAM_OK;
      if (*n_bytes > cap_bytes) return FLOAM_ERR_CAPACITY;
      std::memcpy(out, rec.data(), *n_bytes);
      return FLOAM_OK;
    }
    case FLOAM_DBG_LM: {
      double lm[47];
      lm[0] = S->lm_iterations_last; lm[1] = S->lm_accepted_last; lm[2] = S->initial_cost; lm[3] = S->lm_final_cost_last; lm[4] = S->lm_termination_last;
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) {
          const int lo = std::min(a, b), hi = std::max(a, b);
          lm[5 + a * 6 + b] = S->H0[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
        }
      for (int a = 0; a < 6; ++a) lm[41 + a] = S->g0[a];
      *n_bytes = sizeof(lm);
      if (!out) return FLOAM_OK;
      if (*n_bytes > cap_bytes) return FLOAM_ERR_CAPACITY;
      std::memcpy(out, lm, sizeof(lm));
      return FLOAM_OK;
    }
    case FLOAM_DBG_SCALARS: {
      const int sc[6] = {S->outer_iterations, S->keyframe, nde, nds, S->n_corr, S->skip_solve};
      *n_bytes = sizeof(sc);
      if (!out) return FLOAM_OK;
      if (*n_bytes > cap_bytes) return FLOAM_ERR_CAPACITY;
      std::memcpy(out, sc, sizeof(sc));
      return FLOAM_OK;
    }
    case FLOAM_DBG_CLOCKS: {
      *n_bytes = sizeof(S->dbg_clk);
      if (!out) return FLOAM_OK;
      if (*n_bytes > cap_bytes) return FLOAM_ERR_CAPACITY;
      std::memcpy(out, S->dbg_clk, sizeof(S->dbg_clk));
      return FLOAM_OK;
    }
    case FLOAM_DBG_TIMELINE: {
      long long tl[8] = {S->tl_sum[0], S->tl_sum[1], S->tl_sum[2], S->tl_sum[3], S->tl_predict, S->tl_finish, S->tl_end[0], S->tl_end[1]};
      *n_bytes = sizeof(tl);
      if (!out) return FLOAM_OK;
      if (*n_bytes > cap_bytes) return FLOAM_ERR_CAPACITY;
      std::memcpy(out, tl, sizeof(tl));
      return FLOAM_OK;
    }
    case FLOAM_DBG_FEATURE_SRC_EDGE: return copy_dev(c->d_edge_src, (size_t)ne * 4);
    case FLOAM_DBG_FEATURE_SRC_SURF: return copy_dev(c->d_surf_src, (size_t)ns * 4);
    default: return FLOAM_ERR_ARG;
  }
}

int floam_launch_count(floam_ctx* c, int64_t* launches, int reset) {
  if (!c || !launches) return FLOAM_ERR_ARG;
  *launches = (int64_t)(g_launches - c->launches_base);
  if (reset) c->launches_base = g_launches;
  return FLOAM_OK;
}

int floam_last_frame_ms(floam_ctx* c, float* ms) {
  if (!c || !ms) return FLOAM_ERR_ARG;
  *ms = c->last_frame_ms;
  return FLOAM_OK;
}

int floam_replay_staged(floam_ctx* c, int first, int count, int deskew, double* poses_out, float* total_ms) {
  if (!c || !c->d_staged || first < 0 || count < 1 || first + count >= (int)c->staged_offsets.size() + 0 || c->inflight != 0) return FLOAM_ERR_ARG;
  if (count > c->odom.traj_cap) return FLOAM_ERR_CAPACITY;
  if (set_device(c)) return FLOAM_ERR_CUDA;
  int rc = sync_state(c);
  if (rc) return rc;
  const int counter0 = c->h_state[0]->frame_counter;
  FLOAM_CUDA_OK(cudaMemsetAsync(&c->odom.state->error_sticky, 0, 4, c->stream));
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_replay_begin, c->front_stream));
  int last_slot = 0;
  const auto t_host0 = std::chrono::steady_clock::now();
  for (int f = first; f < first + count; ++f) {
    last_slot = c->submit_slot;
    if (f - first >= 4) {
      // this mailbox last carried frame f - 4: once it has arrived the host is at most four frames ahead of the device (three stay
      // queued, the device never waits) and knows the map sizes of a recent frame when it picks the next frame's graph
      FLOAM_CUDA_OK(cudaEventSynchronize(c->ev_end[last_slot]));
      note_map_sizes(c, c->h_state[last_slot]);
    }
    if ((rc = enqueue_staged_frame(c, f, deskew))) return rc;
  }
  FLOAM_CUDA_OK(cudaEventRecord(c->ev_replay_end, c->stream));
  const auto t_host1 = std::chrono::steady_clock::now();
  FLOAM_CUDA_OK(cudaStreamSynchronize(c->stream));
  if (getenv("FLOAM_DBG_ENQUEUE"))   // development aid: is the replay bound by the host's enqueue rate or by the device?
    fprintf(stderr, "replay_staged: %d frames enqueued in %.3f ms of host time (%.4f ms/frame)\n", count,
            std::chrono::duration<double, std::milli>(t_host1 - t_host0).count(),
            std::chrono::duration<double, std::milli>(t_host1 - t_host0).count() / count);
  if ((rc = check_async("replay_staged"))) return rc;
  cudaEventElapsedTime(&c->last_frame_ms, c->ev_replay_begin, c->ev_replay_end);
  if (total_ms) *total_ms = c->last_frame_ms;
  if (poses_out) {
    const int cap = c->odom.traj_cap;
    for (int k = 0; k < count; ++k)
      FLOAM_CUDA_OK(cudaMemcpy(poses_out + (size_t)k * 7, c->odom.traj + (size_t)((counter0 + k) % cap) * 7, 56, cudaMemcpyDeviceToHost));
  }
  // flags are per frame; the device keeps the OR over the frames of this replay in error_sticky (mailed before the last frame's own
  // flags are folded in, which status_from_flags covers): bits 0-1 map / grid capacity, bits 4-5 feature flags
  const int sticky = c->h_state[last_slot]->error_sticky;
  if (sticky & 0x10) return FLOAM_ERR_NONFINITE;
  if (sticky & 0x23) return FLOAM_ERR_CAPACITY;
  return status_from_flags(c, last_slot);
}

int floam_set_kernel_timing(floam_ctx* c, int enabled) {
  if (!c) return FLOAM_ERR_ARG;
  if (set_device(c)) return FLOAM_ERR_CUDA;
  LaunchTimer& t = c->timer;
  if (enabled && !t.created) {
    for (int i = 0; i < 2 * LaunchTimer::kPairs; ++i) FLOAM_CUDA_OK(cudaEventCreate(&t.ev[i]));
    t.created = true;
  }
  if (!enabled && t.enabled) launch_timer_collect(&t, c->stream);
  if (enabled && !t.enabled) {
    for (int k = 0; k < K_NUM_SLOTS; ++k) { t.total_ms[k] = 0.0; t.launches[k] = 0; }
    t.used = t.persist;
  }
  t.enabled = enabled != 0;
  g_timer = t.enabled ? &t : nullptr;   // the timer follows the calling thread, like the launch counter
  return FLOAM_OK;
}

int floam_kernel_slots(void) { return K_NUM_SLOTS; }
const char* floam_kernel_name(int slot) { return kernel_slot_name(slot); }

int floam_kernel_timing(floam_ctx* c, int slot, double* total_ms, int64_t* launches) {
  if (!c || slot < 0 || slot >= K_NUM_SLOTS) return FLOAM_ERR_ARG;
  if (c->timer.enabled) launch_timer_collect(&c->timer, c->stream);
  if (total_ms) *total_ms = c->timer.total_ms[slot];
  if (launches) *launches = c->timer.launches[slot];
  return FLOAM_OK;
}

int floam_set_graphs(floam_ctx* c, int enabled) {
  if (!c) return FLOAM_ERR_ARG;
  c->use_graphs = enabled != 0;
  return FLOAM_OK;
}

int floam_set_map_merge(floam_ctx* c, int mode) {
  if (!c || c->inflight != 0 || mode < 0 || mode > 2) return FLOAM_ERR_ARG;
  if (set_device(c)) return FLOAM_ERR_CUDA;
  FLOAM_CUDA_OK(cudaDeviceSynchronize());
  c->odom.map_merge_mode = mode;
  for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second.exec);   // the captured launch sequences bake the choice in
  c->graphs.clear();
  c->timer.persist = 0; c->timer.used = 0;
  return FLOAM_OK;
}

}  // extern "C"
