// Feature extraction (subsystem 1) — device entry point and workspace. See feature.cu.
#pragma once
#include "common.cuh"

namespace floam {

struct FeatureParams {
  double min_distance, max_distance;
  int num_lines;
};

struct FeatureWorkspace {
  PointIRT* ring_pts;  // gated points bucketed by ring (input order kept inside a ring)
  int* ring_src;       // index of each ring point in the input scan
  int* surf_tmp;       // per sector: ring-array positions of surf points in ascending-curvature order
  int* tile_off;       // [num_lines * ntiles + 1] counts -> exclusive offsets (ring-major)
  int* edge_tmp;       // [nsectors * 20]
  int *edge_cnt, *surf_cnt, *edge_off, *surf_off;  // [nsectors]
  int ntiles, nsectors, max_scan_points;
};

size_t feature_workspace_bytes_padded(int max_scan_points, int num_lines);
void feature_workspace_bind(FeatureWorkspace& ws, void* mem, int max_scan_points, int num_lines);

// d_flags bit 0: non-finite input; bit 1: a sector exceeded the shared-memory capacity (ring longer than ~6150 points)
void feature_extract_device(const PointIRT* d_scan, const int* d_n, const FeatureParams& prm, FeatureWorkspace& ws, PointIRT* d_edge, int* d_ne,
                            PointIRT* d_surf, int* d_ns, int* d_edge_src, int* d_surf_src, int* d_flags, cudaStream_t s);

}  // namespace floam
