// ORACLE — TEST INFRASTRUCTURE ONLY (see types.h).  Pinned: tests/test_reference_pin.py runs this restatement against the reference's own
// sources compiled unmodified (oracle/_ref, `make ref`) on identical inputs — identical selections, bytes and poses.
// Restates src/laserProcessingClass.cpp:11-22 (RingExtractionVelodyne), :72-118 (featureExtraction),
// :121-231 (featureExtractionFromSector).  Compile with -ffp-contract=off: the reference build has no FMA
// (CMakeLists.txt:4-6, no -march) so the float 11-tap sum and the double squares are plain add/mul.
#include "floam_oracle.h"
#include <algorithm>

namespace fo {
namespace {

struct Double2d {  // include/laserProcessingClass.h:22-27
  int id;
  double value;
};

struct RingCloud {
  std::vector<PointXYZIRT> points;
  std::vector<int> src;  // index in the input cloud (oracle bookkeeping only)
};

// src/laserProcessingClass.cpp:121-231
void featureExtractionFromSector(const RingCloud& pc_in, std::vector<Double2d>& cloudCurvature, CloudIRT& pc_out_edge,
                                 CloudIRT& pc_out_surf, bool total_order, FeatureStats* stats, std::vector<int>* edge_src,
                                 std::vector<int>* surf_src) {
  if (total_order) {
    std::sort(cloudCurvature.begin(), cloudCurvature.end(), [](const Double2d& a, const Double2d& b) {
      return a.value < b.value || (a.value == b.value && a.id < b.id);
    });
  } else {
    std::sort(cloudCurvature.begin(), cloudCurvature.end(), [](const Double2d& a, const Double2d& b) { return a.value < b.value; });
  }
  if (stats) {
    for (size_t i = 1; i < cloudCurvature.size(); ++i)
      if (cloudCurvature[i].value == cloudCurvature[i - 1].value) stats->curvature_ties++;
    stats->sectors++;
  }

  int largestPickedNum = 0;
  std::vector<int> picked_points;
  for (int i = (int)cloudCurvature.size() - 1; i >= 0; i--) {
    int ind = cloudCurvature[i].id;
    if (std::find(picked_points.begin(), picked_points.end(), ind) == picked_points.end()) {
      if (cloudCurvature[i].value <= 0.1) break;
      largestPickedNum++;
      picked_points.push_back(ind);
      if (largestPickedNum <= 20) {
        pc_out_edge.push_back(pc_in.points[ind]);
        if (edge_src) edge_src->push_back(pc_in.src[ind]);
      } else {
        break;
      }
      for (int k = 1; k <= 5; k++) {
        double diffX = pc_in.points[ind + k].x - pc_in.points[ind + k - 1].x;
        double diffY = pc_in.points[ind + k].y - pc_in.points[ind + k - 1].y;
        double diffZ = pc_in.points[ind + k].z - pc_in.points[ind + k - 1].z;
        if (diffX * diffX + diffY * diffY + diffZ * diffZ > 0.05) break;
        picked_points.push_back(ind + k);
      }
      for (int k = -1; k >= -5; k--) {
        double diffX = pc_in.points[ind + k].x - pc_in.points[ind + k + 1].x;
        double diffY = pc_in.points[ind + k].y - pc_in.points[ind + k + 1].y;
        double diffZ = pc_in.points[ind + k].z - pc_in.points[ind + k + 1].z;
        if (diffX * diffX + diffY * diffY + diffZ * diffZ > 0.05) break;
        picked_points.push_back(ind + k);
      }
    }
  }
  for (int i = 0; i <= (int)cloudCurvature.size() - 1; i++) {
    int ind = cloudCurvature[i].id;
    if (std::find(picked_points.begin(), picked_points.end(), ind) == picked_points.end()) {
      pc_out_surf.push_back(pc_in.points[ind]);
      if (surf_src) surf_src->push_back(pc_in.src[ind]);
    }
  }
}

}  // namespace

void LaserProcessing::featureExtraction(const CloudIRT& pc_in, CloudIRT& pc_out_edge, CloudIRT& pc_out_surf, bool total_order,
                                        FeatureStats* stats, std::vector<int>* edge_src, std::vector<int>* surf_src) const {
  // removeNaNFromPointCloud(*pc_in, indices) is the index-only overload: the cloud is untouched (Q9).
  int N_SCANS = lidar_param.num_lines;
  std::vector<RingCloud> laserCloudScans(N_SCANS);

  // RingExtractionVelodyne, :11-22
  for (int i = 0; i < (int)pc_in.size(); i++) {
    const int scanID = pc_in[i].ring;
    double distance = std::sqrt(pc_in[i].x * pc_in[i].x + pc_in[i].y * pc_in[i].y);  // float products+sum, float sqrt (math.h overload set via pcl_macros.h), widened
    if (distance < lidar_param.min_distance || distance > lidar_param.max_distance) continue;
    if (scanID < 0 || scanID >= N_SCANS) continue;  // the reference would index out of bounds; inputs must respect num_lines
    PointXYZIRT p_tmp{};
    p_tmp.x = pc_in[i].x; p_tmp.y = pc_in[i].y; p_tmp.z = pc_in[i].z;
    p_tmp._pad0 = 1.0f;
    p_tmp.intensity = pc_in[i].intensity; p_tmp.ring = pc_in[i].ring; p_tmp.time = pc_in[i].time;
    laserCloudScans[scanID].points.push_back(p_tmp);
    laserCloudScans[scanID].src.push_back(i);
  }

  for (int i = 0; i < N_SCANS; i++) {
    const std::vector<PointXYZIRT>& P = laserCloudScans[i].points;
    if (P.size() < 131) continue;
    if (stats) stats->rings_used++;
    std::vector<Double2d> cloudCurvature;
    int total_points = (int)P.size() - 10;
    for (int j = 5; j < (int)P.size() - 5; j++) {
      // float left-to-right sums (int 10 is converted to float), then widened to double (:96-98)
      double diffX = P[j - 5].x + P[j - 4].x + P[j - 3].x + P[j - 2].x + P[j - 1].x - 10 * P[j].x + P[j + 1].x + P[j + 2].x + P[j + 3].x + P[j + 4].x + P[j + 5].x;
      double diffY = P[j - 5].y + P[j - 4].y + P[j - 3].y + P[j - 2].y + P[j - 1].y - 10 * P[j].y + P[j + 1].y + P[j + 2].y + P[j + 3].y + P[j + 4].y + P[j + 5].y;
      double diffZ = P[j - 5].z + P[j - 4].z + P[j - 3].z + P[j - 2].z + P[j - 1].z - 10 * P[j].z + P[j + 1].z + P[j + 2].z + P[j + 3].z + P[j + 4].z + P[j + 5].z;
      cloudCurvature.push_back(Double2d{j, diffX * diffX + diffY * diffY + diffZ * diffZ});
    }
    for (int j = 0; j < 6; j++) {
      int sector_length = (int)(total_points / 6);
      int sector_start = sector_length * j;
      int sector_end = sector_length * (j + 1) - 1;
      if (j == 5) sector_end = total_points - 1;
      // exclusive end: the last element of every sector is dropped (Q5)
      std::vector<Double2d> subCloudCurvature(cloudCurvature.begin() + sector_start, cloudCurvature.begin() + sector_end);
      featureExtractionFromSector(laserCloudScans[i], subCloudCurvature, pc_out_edge, pc_out_surf, total_order, stats, edge_src, surf_src);
    }
  }
}

}  // namespace fo
