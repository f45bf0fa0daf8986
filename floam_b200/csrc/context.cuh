// floam_ctx: owns the device buffers, workspaces, streams and the host mirror of the small per-frame state.
#pragma once
#include <map>
#include <vector>

#include "common.cuh"
#include "feature.cuh"
#include "voxel.cuh"
#include "odom.cuh"
#include "imu.cuh"
#include "mapping.cuh"

struct floam_graph_key {
  int kind;           // 0 = first frame (initMapWithPoints), 1 = update
  int outer;          // optimization_count of the update
  int deskew;
  int slot;           // scan buffer the graph reads
  bool operator<(const floam_graph_key& o) const {
    if (kind != o.kind) return kind < o.kind;
    if (outer != o.outer) return outer < o.outer;
    if (deskew != o.deskew) return deskew < o.deskew;
    return slot < o.slot;
  }
};

struct floam_graph_entry {
  cudaGraphExec_t exec = nullptr;
  int launches = 0;   // kernels inside the graph (bench.py's gpu_launches)
  int pair_first = 0, pair_last = 0;   // event pairs recorded inside the graph (kernel-timing mode)
};

struct floam_ctx {
  floam_params prm;
  int device = 0;
  cudaStream_t stream = nullptr;       // all kernels
  cudaStream_t copy_stream = nullptr;  // uploads of the next scan
  cudaEvent_t ev_begin[4] = {nullptr, nullptr, nullptr, nullptr}, ev_end[4] = {nullptr, nullptr, nullptr, nullptr};   // per mailbox (frame ring index)
  cudaEvent_t ev_upload[2] = {nullptr, nullptr};
  cudaEvent_t ev_replay_begin = nullptr, ev_replay_end = nullptr;
  bool consumed_valid[2] = {false, false};   // a FRONT has read d_scan[slot] before (ev_front_done[slot] is meaningful)
  std::vector<void*> allocs;           // everything cudaMalloc'ed
  std::vector<void*> host_allocs;      // everything cudaMallocHost'ed

  // scan + features (device)
  floam::PointIRT* d_scan[2] = {nullptr, nullptr};  // double-buffered upload target
  unsigned char* d_raw[2] = {nullptr, nullptr};     // raw PointCloud2 bytes (floam_process_submit_pc2), allocated on first use
  size_t raw_cap = 0;
  int* d_scan_n[2] = {nullptr, nullptr};
  // feature clouds: two buffer sets (frame parity); the unsuffixed names alias the set of the frame being enqueued / last completed
  floam::PointIRT *d_edge = nullptr, *d_surf = nullptr;
  int *d_ne = nullptr, *d_ns = nullptr, *d_edge_src = nullptr, *d_surf_src = nullptr;
  floam::PointIRT *d_edge_b[2] = {nullptr, nullptr}, *d_surf_b[2] = {nullptr, nullptr};
  int *d_ne_b[2] = {nullptr, nullptr}, *d_ns_b[2] = {nullptr, nullptr}, *d_edge_src_b[2] = {nullptr, nullptr}, *d_surf_src_b[2] = {nullptr, nullptr};
  int* d_flags = nullptr;                       // feature-extraction flags word of the frame being enqueued (alias of d_flags_b[parity])
  int* d_flags_b[2] = {nullptr, nullptr};       // one per frame parity, zeroed at the head of the frame's FRONT: flags are per frame
  floam::FeatureParams fprm;
  floam::FeatureWorkspace fws;
  floam::VoxelWorkspace vws;
  floam::VoxelWorkspace vws_aux;        // edge-side branch (sized for scans and local maps)
  cudaStream_t aux_stream = nullptr;
  // frame pipeline: the FRONT of frame k+1 (features + downsampling, independent of pose and map) runs on its own stream pair while
  // the BACK of frame k (prediction, association + solve, keyframe map update) is still running on stream / aux_stream
  floam::VoxelWorkspace vws_front, vws_front_aux;
  cudaStream_t front_stream = nullptr, front_aux = nullptr;
  cudaEvent_t ev_ffork = nullptr, ev_fjoin = nullptr;
  cudaEvent_t ev_front_done[2] = {nullptr, nullptr}, ev_back_done[2] = {nullptr, nullptr};
  bool back_valid[2] = {false, false};

  // staging for the stage entry points (voxel/crop/knn/set_map on host clouds)
  char* d_stage_in = nullptr;            // stage_cap x 32 B
  floam::P4* d_stage_p4 = nullptr;       // stage_cap
  floam::P4* d_stage_out = nullptr;      // stage_cap
  int* d_stage_n = nullptr;              // 4 ints
  float* d_stage_bounds = nullptr;       // 6 floats
  int stage_cap = 0;
  floam::LocalMap knn_map;               // lazily allocated map for floam_knn5
  bool knn_map_ready = false;
  int* d_knn_ids = nullptr;
  float* d_knn_d2 = nullptr;

  floam::OdomDevice odom;      // odometry state, maps, grids, LM (odom.cuh)
  floam::ImuDevice imu;        // dmapping::ImuHandler mirror + device samples (imu.cuh)
  floam::DeskewPlan* d_plan[2] = {nullptr, nullptr};   // per scan slot, for the fused IMU path
  floam::MappingDevice mapping;

  // pinned host mailboxes
  int* h_ints = nullptr;                 // 64 ints
  double* h_doubles = nullptr;           // 64 doubles
  floam::PoseState* h_state[4] = {nullptr, nullptr, nullptr, nullptr};   // ring of four: up to three frames in flight
  int* h_flags[4] = {nullptr, nullptr, nullptr, nullptr};

  // staged scans for device-resident replay
  floam::PointIRT* d_staged = nullptr;
  int* d_staged_n = nullptr;
  int* d_staged_counts = nullptr;
  std::vector<long long> staged_offsets;

  // submit/wait pipeline
  int inflight = 0, submit_slot = 0, wait_slot = 0;   // frame ring indices 0..3: mailbox = index, buffer parity = index & 1
  bool map_initialised = false;
  bool use_graphs = true;
  std::map<floam_graph_key, floam_graph_entry> graphs;
  float last_frame_ms = 0.f;
  floam::LaunchTimer timer;
  long long launches_base = 0;
};

namespace floam {
void* ctx_alloc(void* ctx, size_t bytes);   // cudaMalloc tracked by the context; nullptr on failure
}
