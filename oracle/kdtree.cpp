// ORACLE — TEST INFRASTRUCTURE ONLY (see types.h).  Restates FLANN (un-vendored): checked against the FLANN copy bundled with OpenCV (tests/test_oracle.py), otherwise unpinned.
// Restates pcl::KdTreeFLANN<PointXYZI>::{setInputCloud,nearestKSearch} (PCL 1.8.1) over FLANN 1.9.1's
// KDTreeSingleIndex<L2_Simple<float>> with KDTreeSingleIndexParams(15), KNNSimpleResultSet, eps=0, sorted=true
// (un-vendored).  Reference call sites: src/odomEstimationClass.cpp:17-18,78-79,153,206.
#include "floam_oracle.h"
#include <algorithm>
#include <limits>

namespace fo {

namespace {
const int kLeafMaxSize = 15;
inline float l2_simple(const float* a, const float* b) {  // flann::L2_Simple<float>::operator()
  float result = 0.f;
  for (int i = 0; i < 3; ++i) {
    float diff = a[i] - b[i];
    result += diff * diff;
  }
  return result;
}
inline float accum_dist(float a, float b) { return (a - b) * (a - b); }
}  // namespace

struct KdTreeFlann::ResultSet {  // flann::KNNSimpleResultSet
  int capacity, count;
  float worst;
  float dist[8];
  int index[8];
  explicit ResultSet(int k) : capacity(k), count(0), worst(std::numeric_limits<float>::max()) {
    for (int i = 0; i < 8; ++i) { dist[i] = std::numeric_limits<float>::max(); index[i] = -1; }
  }
  void addPoint(float d, int idx) {
    if (d >= worst) return;
    if (count < capacity) ++count;
    int i;
    for (i = count - 1; i > 0; --i) {
      if (dist[i - 1] > d) { dist[i] = dist[i - 1]; index[i] = index[i - 1]; }
      else break;
    }
    dist[i] = d; index[i] = idx;
    worst = dist[capacity - 1];
  }
};

void KdTreeFlann::setInputCloud(const PointXYZI* cloud, size_t n) {
  n_ = n;
  pts_.resize(n_ * 3);
  for (size_t i = 0; i < n_; ++i) { pts_[3 * i] = cloud[i].x; pts_[3 * i + 1] = cloud[i].y; pts_[3 * i + 2] = cloud[i].z; }
  vind_.resize(n_);
  for (size_t i = 0; i < n_; ++i) vind_[i] = (int)i;
  nodes_.clear();
  root_ = -1;
  if (n_ == 0) return;
  // computeBoundingBox
  for (int d = 0; d < 3; ++d) root_bbox_[d].low = root_bbox_[d].high = pts_[d];
  for (size_t k = 1; k < n_; ++k)
    for (int d = 0; d < 3; ++d) {
      float v = pts_[3 * k + d];
      if (v < root_bbox_[d].low) root_bbox_[d].low = v;
      if (v > root_bbox_[d].high) root_bbox_[d].high = v;
    }
  nodes_.reserve(2 * n_ / kLeafMaxSize + 16);
  Interval bbox[3] = {root_bbox_[0], root_bbox_[1], root_bbox_[2]};
  root_ = divideTree(0, (int)n_, bbox);
  root_bbox_[0] = bbox[0]; root_bbox_[1] = bbox[1]; root_bbox_[2] = bbox[2];
  data_.resize(n_ * 3);  // reorder_ = true
  for (size_t i = 0; i < n_; ++i)
    for (int d = 0; d < 3; ++d) data_[3 * i + d] = pts_[3 * (size_t)vind_[i] + d];
}

int KdTreeFlann::divideTree(int left, int right, Interval bbox[3]) {
  int id = (int)nodes_.size();
  nodes_.push_back(Node());
  if ((right - left) <= kLeafMaxSize) {
    Node nd; nd.child1 = nd.child2 = -1; nd.left = left; nd.right = right; nd.divfeat = 0; nd.divlow = nd.divhigh = 0;
    for (int d = 0; d < 3; ++d) bbox[d].low = bbox[d].high = pts_[3 * (size_t)vind_[left] + d];
    for (int k = left + 1; k < right; ++k)
      for (int d = 0; d < 3; ++d) {
        float v = pts_[3 * (size_t)vind_[k] + d];
        if (bbox[d].low > v) bbox[d].low = v;
        if (bbox[d].high < v) bbox[d].high = v;
      }
    nodes_[id] = nd;
  } else {
    int idx, cutfeat;
    float cutval;
    middleSplit(&vind_[0] + left, right - left, idx, cutfeat, cutval, bbox);
    Node nd; nd.left = nd.right = 0; nd.divfeat = cutfeat;
    Interval left_bbox[3] = {bbox[0], bbox[1], bbox[2]};
    left_bbox[cutfeat].high = cutval;
    nd.child1 = divideTree(left, left + idx, left_bbox);
    Interval right_bbox[3] = {bbox[0], bbox[1], bbox[2]};
    right_bbox[cutfeat].low = cutval;
    nd.child2 = divideTree(left + idx, right, right_bbox);
    nd.divlow = left_bbox[cutfeat].high;
    nd.divhigh = right_bbox[cutfeat].low;
    for (int d = 0; d < 3; ++d) {
      bbox[d].low = std::min(left_bbox[d].low, right_bbox[d].low);
      bbox[d].high = std::max(left_bbox[d].high, right_bbox[d].high);
    }
    nodes_[id] = nd;
  }
  return id;
}

void KdTreeFlann::middleSplit(int* ind, int count, int& index, int& cutfeat, float& cutval, const Interval bbox[3]) {
  float max_span = bbox[0].high - bbox[0].low;
  cutfeat = 0;
  cutval = (bbox[0].high + bbox[0].low) / 2;
  for (int i = 1; i < 3; ++i) {
    float span = bbox[i].high - bbox[i].low;
    if (span > max_span) { max_span = span; cutfeat = i; cutval = (bbox[i].high + bbox[i].low) / 2; }
  }
  auto computeMinMax = [&](int dim, float& mn, float& mx) {
    mn = mx = pts_[3 * (size_t)ind[0] + dim];
    for (int i = 1; i < count; ++i) {
      float v = pts_[3 * (size_t)ind[i] + dim];
      if (v < mn) mn = v;
      if (v > mx) mx = v;
    }
  };
  float min_elem, max_elem;
  computeMinMax(cutfeat, min_elem, max_elem);
  cutval = (min_elem + max_elem) / 2;
  max_span = max_elem - min_elem;
  int k = cutfeat;
  for (int i = 0; i < 3; ++i) {
    if (i == k) continue;
    float span = bbox[i].high - bbox[i].low;
    if (span > max_span) {
      computeMinMax(i, min_elem, max_elem);
      span = max_elem - min_elem;
      if (span > max_span) { max_span = span; cutfeat = i; cutval = (min_elem + max_elem) / 2; }
    }
  }
  int lim1, lim2;
  planeSplit(ind, count, cutfeat, cutval, lim1, lim2);
  if (lim1 > count / 2) index = lim1;
  else if (lim2 < count / 2) index = lim2;
  else index = count / 2;
}

void KdTreeFlann::planeSplit(int* ind, int count, int cutfeat, float cutval, int& lim1, int& lim2) {
  int left = 0, right = count - 1;
  for (;;) {
    while (left <= right && pts_[3 * (size_t)ind[left] + cutfeat] < cutval) ++left;
    while (left <= right && pts_[3 * (size_t)ind[right] + cutfeat] >= cutval) --right;
    if (left > right) break;
    std::swap(ind[left], ind[right]); ++left; --right;
  }
  lim1 = left;
  right = count - 1;
  for (;;) {
    while (left <= right && pts_[3 * (size_t)ind[left] + cutfeat] <= cutval) ++left;
    while (left <= right && pts_[3 * (size_t)ind[right] + cutfeat] > cutval) --right;
    if (left > right) break;
    std::swap(ind[left], ind[right]); ++left; --right;
  }
  lim2 = left;
}

void KdTreeFlann::searchLevel(ResultSet& rs, const float* vec, int node_id, float mindistsq, float dists[3]) const {
  const Node& node = nodes_[node_id];
  if (node.child1 < 0 && node.child2 < 0) {
    float worst_dist = rs.worst;
    for (int i = node.left; i < node.right; ++i) {
      float dist = l2_simple(vec, &data_[3 * (size_t)i]);
      if (dist < worst_dist) rs.addPoint(dist, vind_[i]);
    }
    return;
  }
  int idx = node.divfeat;
  float val = vec[idx];
  float diff1 = val - node.divlow;
  float diff2 = val - node.divhigh;
  int bestChild, otherChild;
  float cut_dist;
  if ((diff1 + diff2) < 0) { bestChild = node.child1; otherChild = node.child2; cut_dist = accum_dist(val, node.divhigh); }
  else { bestChild = node.child2; otherChild = node.child1; cut_dist = accum_dist(val, node.divlow); }
  searchLevel(rs, vec, bestChild, mindistsq, dists);
  float dst = dists[idx];
  mindistsq = mindistsq + cut_dist - dst;
  dists[idx] = cut_dist;
  if (mindistsq * 1.0f <= rs.worst) searchLevel(rs, vec, otherChild, mindistsq, dists);
  dists[idx] = dst;
}

int KdTreeFlann::nearestKSearch(const PointXYZI& q, int k, int* ids, float* sqdist) const {
  if (k > (int)n_) k = (int)n_;
  if (k > 8) k = 8;
  if (k <= 0) return 0;
  const float vec[3] = {q.x, q.y, q.z};
  ResultSet rs(k);
  float dists[3] = {0, 0, 0};
  float distsq = 0.f;
  for (int i = 0; i < 3; ++i) {  // computeInitialDistances
    if (vec[i] < root_bbox_[i].low) { dists[i] = accum_dist(vec[i], root_bbox_[i].low); distsq += dists[i]; }
    if (vec[i] > root_bbox_[i].high) { dists[i] = accum_dist(vec[i], root_bbox_[i].high); distsq += dists[i]; }
  }
  searchLevel(rs, vec, root_, distsq, dists);
  for (int i = 0; i < k; ++i) { ids[i] = rs.index[i]; sqdist[i] = rs.dist[i]; }
  return k;
}

int knn_bruteforce(const PointXYZI* cloud, size_t n_cloud, const PointXYZI& q, int k, int* ids, float* sqdist) {
  if (k > (int)n_cloud) k = (int)n_cloud;
  if (k > 8) k = 8;
  float bd[8];
  int bi[8];
  int cnt = 0;
  const float vec[3] = {q.x, q.y, q.z};
  for (int i = 0; i < (int)n_cloud; ++i) {
    const float p[3] = {cloud[i].x, cloud[i].y, cloud[i].z};
    float d = l2_simple(vec, p);
    if (cnt == k && !(d < bd[k - 1])) continue;  // ascending index scan: a later equal distance never displaces
    int j = (cnt < k) ? cnt++ : k - 1;
    while (j > 0 && bd[j - 1] > d) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
    bd[j] = d; bi[j] = i;
  }
  for (int i = 0; i < k; ++i) { ids[i] = bi[i]; sqdist[i] = bd[i]; }
  return k;
}

void GridKnn::setInputCloud(const PointXYZI* cloud, size_t n) {
  n_ = n;
  pts_.clear(); index_.clear(); cell_start_.assign(1, 0);
  nx_ = ny_ = nz_ = 0;
  if (n == 0) return;
  float mn[3] = {cloud[0].x, cloud[0].y, cloud[0].z}, mx[3] = {cloud[0].x, cloud[0].y, cloud[0].z};
  for (size_t i = 1; i < n; ++i) {
    const float v[3] = {cloud[i].x, cloud[i].y, cloud[i].z};
    for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], v[d]); mx[d] = std::max(mx[d], v[d]); }
  }
  ix0_ = (int)std::floor(mn[0]); iy0_ = (int)std::floor(mn[1]); iz0_ = (int)std::floor(mn[2]);
  nx_ = (int)std::floor(mx[0]) - ix0_ + 1; ny_ = (int)std::floor(mx[1]) - iy0_ + 1; nz_ = (int)std::floor(mx[2]) - iz0_ + 1;
  const size_t ncells = (size_t)nx_ * ny_ * nz_;
  std::vector<int> cell(n);
  cell_start_.assign(ncells + 1, 0);
  for (size_t i = 0; i < n; ++i) {
    const int cx = (int)std::floor(cloud[i].x) - ix0_, cy = (int)std::floor(cloud[i].y) - iy0_, cz = (int)std::floor(cloud[i].z) - iz0_;
    cell[i] = cx + nx_ * (cy + ny_ * cz);
    cell_start_[(size_t)cell[i] + 1]++;
  }
  for (size_t c = 0; c < ncells; ++c) cell_start_[c + 1] += cell_start_[c];
  std::vector<int> cursor(cell_start_.begin(), cell_start_.end() - 1);
  pts_.resize(3 * n); index_.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const int p = cursor[cell[i]]++;
    pts_[3 * (size_t)p] = cloud[i].x; pts_[3 * (size_t)p + 1] = cloud[i].y; pts_[3 * (size_t)p + 2] = cloud[i].z;
    index_[p] = (int)i;
  }
}

int GridKnn::nearestKSearch(const PointXYZI& q, int k, int* ids, float* sqdist) const {
  if (k > 8) k = 8;
  float bd[8]; int bi[8]; int cnt = 0;
  const float vec[3] = {q.x, q.y, q.z};
  if (nx_ > 0 && std::isfinite(q.x) && std::isfinite(q.y) && std::isfinite(q.z) && std::fabs(q.x) < 1e9f && std::fabs(q.y) < 1e9f && std::fabs(q.z) < 1e9f) {
    const int cx = (int)std::floor(q.x) - ix0_, cy = (int)std::floor(q.y) - iy0_, cz = (int)std::floor(q.z) - iz0_;
    for (int z = std::max(cz - 1, 0); z <= std::min(cz + 1, nz_ - 1); ++z)
      for (int y = std::max(cy - 1, 0); y <= std::min(cy + 1, ny_ - 1); ++y) {
        const int x0 = std::max(cx - 1, 0), x1 = std::min(cx + 1, nx_ - 1);
        if (x0 > x1) continue;
        const size_t row = (size_t)nx_ * ((size_t)y + (size_t)ny_ * z);
        for (int p = cell_start_[row + x0]; p < cell_start_[row + x1 + 1]; ++p) {
          const float d = l2_simple(vec, &pts_[3 * (size_t)p]);
          const int idx = index_[p];
          if (cnt == k && !(d < bd[k - 1] || (d == bd[k - 1] && idx < bi[k - 1]))) continue;
          int j = (cnt < k) ? cnt++ : k - 1;
          while (j > 0 && (bd[j - 1] > d || (bd[j - 1] == d && bi[j - 1] > idx))) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
          bd[j] = d; bi[j] = idx;
        }
      }
  }
  const bool near = cnt == k && bd[k - 1] < 1.0f;
  for (int i = 0; i < k; ++i) { ids[i] = near ? bi[i] : -1; sqdist[i] = near ? bd[i] : std::numeric_limits<float>::max(); }
  return k;
}

}  // namespace fo
