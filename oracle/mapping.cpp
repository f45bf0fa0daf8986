// ORACLE — TEST INFRASTRUCTURE ONLY (see types.h).  Pinned: tests/test_reference_pin.py runs this restatement against the reference's own
// sources compiled unmodified (oracle/_ref, `make ref`) on identical inputs — identical selections, bytes and poses.
// Restates src/laserMappingClass.cpp:7-32,106-200 (50 m cell grid that grows by slabs, float transform,
// z-based intensity, per-cell in-place VoxelGrid over the 5x5x5 neighbourhood).
#include "floam_oracle.h"
#include <algorithm>

namespace fo {

static const double LASER_CELL_WIDTH = 50.0, LASER_CELL_HEIGHT = 50.0, LASER_CELL_DEPTH = 50.0;  // include/laserMappingClass.h:26-28
static const int RANGE_H = 2, RANGE_V = 2;                                                        // :32-33

void LaserMapping::init(double map_resolution) {
  map.clear();
  for (int i = 0; i < RANGE_H * 2 + 1; i++) {
    std::vector<std::vector<Cell>> map_height_temp;
    for (int j = 0; j < RANGE_H * 2 + 1; j++) {
      std::vector<Cell> map_depth_temp;
      for (int k = 0; k < RANGE_V * 2 + 1; k++) map_depth_temp.push_back(std::make_shared<CloudI>());
      map_height_temp.push_back(map_depth_temp);
    }
    map.push_back(map_height_temp);
  }
  origin_in_map_x = RANGE_H; origin_in_map_y = RANGE_H; origin_in_map_z = RANGE_V;
  map_width = RANGE_H * 2 + 1; map_height = RANGE_H * 2 + 1; map_depth = RANGE_H * 2 + 1;
  leaf_ = (float)map_resolution;
}

void LaserMapping::checkPoints(int& x, int& y, int& z) {
  while (x + RANGE_H > map_width - 1) {  // addWidthCellPositive
    map.push_back(std::vector<std::vector<Cell>>(map_height, std::vector<Cell>(map_depth)));
    map_width++;
  }
  while (x - RANGE_H < 0) {  // addWidthCellNegative
    map.insert(map.begin(), std::vector<std::vector<Cell>>(map_height, std::vector<Cell>(map_depth)));
    origin_in_map_x++; map_width++;
    x++;
  }
  while (y + RANGE_H > map_height - 1) {  // addHeightCellPositive
    for (int i = 0; i < map_width; i++) map[i].push_back(std::vector<Cell>(map_depth));
    map_height++;
  }
  while (y - RANGE_H < 0) {  // addHeightCellNegative
    for (int i = 0; i < map_width; i++) map[i].insert(map[i].begin(), std::vector<Cell>(map_depth));
    origin_in_map_y++; map_height++;
    y++;
  }
  while (z + RANGE_V > map_depth - 1) {  // addDepthCellPositive
    for (int i = 0; i < map_width; i++)
      for (int j = 0; j < map_height; j++) map[i][j].push_back(Cell());
    map_depth++;
  }
  while (z - RANGE_V < 0) {  // addDepthCellNegative
    for (int i = 0; i < map_width; i++)
      for (int j = 0; j < map_height; j++) map[i][j].insert(map[i][j].begin(), Cell());
    origin_in_map_z++; map_depth++;
    z++;
  }
  for (int i = x - RANGE_H; i < x + RANGE_H + 1; i++)
    for (int j = y - RANGE_H; j < y + RANGE_H + 1; j++)
      for (int k = z - RANGE_V; k < z + RANGE_V + 1; k++)
        if (!map[i][j][k]) map[i][j][k] = std::make_shared<CloudI>();
}

void LaserMapping::updateCurrentPointsToMap(const CloudI& pc_in, const Iso3& pose) {
  int currentPosIdX = int(std::floor(pose.t.x / LASER_CELL_WIDTH + 0.5)) + origin_in_map_x;
  int currentPosIdY = int(std::floor(pose.t.y / LASER_CELL_HEIGHT + 0.5)) + origin_in_map_y;
  int currentPosIdZ = int(std::floor(pose.t.z / LASER_CELL_DEPTH + 0.5)) + origin_in_map_z;
  checkPoints(currentPosIdX, currentPosIdY, currentPosIdZ);

  // pcl::transformPointCloud(in, out, pose.cast<float>()): float R*p + t
  float R[3][3], t[3] = {(float)pose.t.x, (float)pose.t.y, (float)pose.t.z};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) R[a][b] = (float)pose.R.m[a][b];
  for (int i = 0; i < (int)pc_in.size(); i++) {
    const PointXYZI& p = pc_in[i];
    PointXYZI point_temp = p;
    point_temp.x = static_cast<float>(R[0][0] * p.x + R[0][1] * p.y + R[0][2] * p.z + t[0]);
    point_temp.y = static_cast<float>(R[1][0] * p.x + R[1][1] * p.y + R[1][2] * p.z + t[1]);
    point_temp.z = static_cast<float>(R[2][0] * p.x + R[2][1] * p.y + R[2][2] * p.z + t[2]);
    point_temp.intensity = (float)std::min(1.0, std::max(p.z + 2.0, 0.0) / 5);
    int currentPointIdX = int(std::floor(point_temp.x / LASER_CELL_WIDTH + 0.5)) + origin_in_map_x;
    int currentPointIdY = int(std::floor(point_temp.y / LASER_CELL_HEIGHT + 0.5)) + origin_in_map_y;
    int currentPointIdZ = int(std::floor(point_temp.z / LASER_CELL_DEPTH + 0.5)) + origin_in_map_z;
    // the reference dereferences whatever is there; points outside the allocated 5x5x5 block would crash it
    if (currentPointIdX < 0 || currentPointIdX >= map_width || currentPointIdY < 0 || currentPointIdY >= map_height ||
        currentPointIdZ < 0 || currentPointIdZ >= map_depth || !map[currentPointIdX][currentPointIdY][currentPointIdZ])
      continue;
    map[currentPointIdX][currentPointIdY][currentPointIdZ]->push_back(point_temp);
  }
  for (int i = currentPosIdX - RANGE_H; i < currentPosIdX + RANGE_H + 1; i++)
    for (int j = currentPosIdY - RANGE_H; j < currentPosIdY + RANGE_H + 1; j++)
      for (int k = currentPosIdZ - RANGE_V; k < currentPosIdZ + RANGE_V + 1; k++) {
        CloudI filtered;
        voxel_grid_filter(*map[i][j][k], leaf_, filtered, total_order);
        map[i][j][k]->swap(filtered);
      }
}

CloudI LaserMapping::getMap() const {
  CloudI out;
  for (int i = 0; i < map_width; i++)
    for (int j = 0; j < map_height; j++)
      for (int k = 0; k < map_depth; k++)
        if (map[i][j][k]) out.insert(out.end(), map[i][j][k]->begin(), map[i][j][k]->end());
  return out;
}

}  // namespace fo
