// Subsystem (2): pcl::VoxelGrid / pcl::CropBox replacements — device entry points. See voxel.cu.
#pragma once
#include "common.cuh"

namespace floam {

struct VoxelWorkspace {
  unsigned int* keys;   // [n_max]
  int* vals;            // [n_max]
  int* flags;           // [n_max + 1] segment-head flags -> exclusive scan (segment ids)
  unsigned int* bbox;   // 6 order-preserving-encoded floats: min xyz, max xyz
  int* d_nbits;         // key width of the current grid
  int* d_passthrough;   // Q13: leaf too small for the extent -> output = input
  int* d_counts;        // [0] points of the running filter's input, [1] points its crop box kept, [2] MAIN / [3] SIDE points (merge path)
  unsigned int* main_keys;   // [n_max] merge path: keys / indices of the points that are already in order
  int* main_vals;
  unsigned long long* merge_state;   // look-back states of voxel_classify_kernel, two words per 4096-point tile
  SortWorkspace sort;
  ScanWorkspace scan;
  int n_max;
  bool large_input = false;   // host-side hint for the grids of the next filter call: hundreds of thousands of points expected (four CTAs per SM instead of two)
};
size_t voxel_workspace_bytes(int n_max);
void voxel_workspace_bind(VoxelWorkspace& ws, void* mem, int n_max);
// one-time initialisation of the bounding-box accumulators (every filter re-arms them for the next one)
int voxel_workspace_arm(VoxelWorkspace& ws, cudaStream_t s);

// pcl::VoxelGrid<PointXYZI>::filter (PCL 1.8.1 semantics). Input points are read with a byte stride
// (16 = float4 xyzi, 32 = PointXYZI / PointXYZIRT with intensity at +16). Output order = ascending voxel index; inside a
// voxel the float accumulation runs in ascending input index (the stable stand-in for std::sort's unspecified order).
// d_skip (optional): when *d_skip != 0 every kernel returns immediately (device-side "not a keyframe").
// d_crop (optional, 6 floats on the device: min xyz, max xyz): pcl::CropBox folded in — points outside the inclusive box are ignored,
// exactly as if CropBox::filter had run first. d_extra / cap: the input holds *d_n + *d_extra points when that fits into cap.
// append (optional; addPointsToMap :256-268 folded into the first kernel of the filter): the *d_extra points are not in the input yet —
// the bounding-box kernel reads them from append->src, applies pointAssociateToMap (double q * p + t, float store) with the pose
// at append->pose7 and stores them at d_in[*d_n ...] on the way. When they do not fit into cap, bit 0 of *append->err_flags is set.
// out_bbox (optional): flipped-float min/max accumulators that receive the bounding box of the OUTPUT cloud (the search grid's extent).
struct VoxelAppend {
  const P4* src;
  const double* pose7;   // qx qy qz qw tx ty tz on the device
  int* err_flags;
};
void voxel_grid_device(const void* d_in, int stride_bytes, const int* d_n, int n_max, float leaf, P4* d_out, int* d_nout, VoxelWorkspace& ws,
                       const int* d_skip, cudaStream_t s, const float* d_crop = nullptr, const int* d_extra = nullptr, int cap = 0,
                       const VoxelAppend* append = nullptr, unsigned int* out_bbox = nullptr);

// The same filter for an input that is mostly in voxel order already (a local map + the new frame's points, addPointsToMap): only
// the out-of-place points are sorted and then merged into the rest (voxel.cu). Identical output for ANY input.
void voxel_grid_merge_device(const void* d_in, int stride_bytes, const int* d_n, int n_max, float leaf, P4* d_out, int* d_nout, VoxelWorkspace& ws,
                             const int* d_skip, cudaStream_t s, const float* d_crop = nullptr, const int* d_extra = nullptr, int cap = 0,
                             const VoxelAppend* append = nullptr, unsigned int* out_bbox = nullptr);

// pcl::CropBox<PointXYZI>::filter, identity transform, negative=false, inclusive float bounds read from device memory
// (d_bounds: min xyz, max xyz). Order-preserving compaction.
// d_extra (optional): the input holds *d_n + *d_extra points when that sum fits into cap (points appended since *d_n was written).
void crop_box_device(const P4* d_in, const int* d_n, int n_max, const float* d_bounds, P4* d_out, int* d_nout, VoxelWorkspace& ws,
                     const int* d_skip, cudaStream_t s, const int* d_extra = nullptr, int cap = 0);

// stride-32 PointXYZI / PointXYZIRT cloud -> float4 (x, y, z, intensity)
void repack_xyzi_device(const void* d_in32, const int* d_n, int n_max, P4* d_out, cudaStream_t s);

}  // namespace floam
