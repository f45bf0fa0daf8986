// LaserMappingClass (src/laserMappingClass.cpp:7-200) on the device: the global map is one flat cloud plus a 50 m cell key per
// point; a frame touches the 5x5x5 block of cells around the sensor exactly like the reference's 125 in-place VoxelGrid filters.
// See mapping.cu.
#pragma once
#include <set>
#include <vector>

#include "common.cuh"
#include "voxel.cuh"

namespace floam {

struct MappingDevice {
  P4* pts = nullptr;          // global map (cells of the last touched block at the tail, each in voxel order)
  unsigned int* cell = nullptr;  // packed 50 m cell id per point: (cx+512) << 20 | (cy+512) << 10 | (cz+512)
  P4* pts_alt = nullptr;      // ping-pong target of an update
  unsigned int* cell_alt = nullptr;
  P4* work = nullptr;         // points of the touched block: old in-block points followed by the new frame
  unsigned int* work_cell = nullptr;
  int* d_counts = nullptr;    // [0] map size, [1] rest, [2] old in block, [3] work size, [4] out-of-block drops, [5] voxels, [6] new count
  int* flags = nullptr;       // predicate / scan buffer, cap + 1
  unsigned int* keys = nullptr;
  int* vals = nullptr;
  int* d_nbits = nullptr;     // [0] bits of the (kx,ky) key, [1] bits of the (kz,cell) key
  VoxelWorkspace* vws = nullptr;
  std::set<unsigned int> allocated;   // cells some block has allocated so far (checkPoints), host mirror
  unsigned int* d_allocated = nullptr; // the same, sorted, on the device
  int alloc_cap = 0;
  // incremental getMap(): one byte per allocated cell (same order as d_allocated), set when an update rewrites the cell or appends to it,
  // cleared when mapping_get_dirty_device hands the cell's points out
  unsigned char* d_dirty = nullptr;
  std::vector<unsigned int> allocated_prev;   // the allocated list as the device currently holds it (sorted)
  std::set<unsigned int> dirty_host;   // block cells of the updates since the last hand-out (marked on the device when the call runs)
  int cap = 0;
  float leaf = 0.4f;
  bool enabled = false;
};

int mapping_device_init(MappingDevice& md, int cap, double map_resolution, VoxelWorkspace* vws, void* (*alloc)(void*, size_t), void* actx, cudaStream_t s);
// updateCurrentPointsToMap :148-186 ; d_in = stride-32 PointXYZI cloud, pose row-major 4x4 (host)
int mapping_update_device(MappingDevice& md, const void* d_in, int stride, const int* d_n, int n_max, const double pose16[16], cudaStream_t s);
// getMap :188-200 : cells in (x, y, z) order, each cell in voxel order. Sorts into the alt buffers; *d_out_n = size.
int mapping_get_map_device(MappingDevice& md, P4** d_out, int** d_out_n, cudaStream_t s);
// Incremental getMap(): the points of every cell changed since the previous call (all of such a cell's points, map order kept, so the
// caller can replace its copy of the cell), with their packed cell ids. *d_out / *d_out_cell / *d_out_n point into the ping-pong
// buffers and stay valid until the next update. The change marks stay set until mapping_clear_dirty_device.
int mapping_get_dirty_device(MappingDevice& md, P4** d_out, unsigned int** d_out_cell, int** d_out_n, cudaStream_t s);
int mapping_clear_dirty_device(MappingDevice& md, cudaStream_t s);

}  // namespace floam
