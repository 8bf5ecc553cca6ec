// oracle/shim/pcl/kdtree/kdtree_flann.h — TEST INFRASTRUCTURE: pcl::PointXYZ / PointCloud / KdTreeFLANN::radiusSearch without PCL.
// radiusSearch returns every point whose float squared distance is <= radius^2, found through a uniform cell grid.  The
// reference only asks "is ANY returned point inside the ellipsoid" (kino_astar.cpp:744-753) and its radius (robot_r + 0.1)
// exceeds the longest semi-axis, so neither the order of the result nor points at the rim of the ball can change the answer
// (SURVEY.md §9.1 Q12).
#pragma once
#include <boost/make_shared.hpp>
#include <cmath>
#include <cstdint>
#include <algorithm>
#include <vector>

namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
template <typename P>
struct PointCloud {
  typedef boost::shared_ptr<PointCloud<P>> Ptr;
  typedef boost::shared_ptr<const PointCloud<P>> ConstPtr;
  std::vector<P> points;
  uint32_t width = 0, height = 0;
};
template <typename P>
class KdTreeFLANN {
  typename PointCloud<P>::ConstPtr cloud_;
  std::vector<int> start_, idx_;     // dense cell grid (CSR): points of cell c are idx_[start_[c] .. start_[c + 1])
  float cell_ = 0.5f, lo_[3] = {0, 0, 0};
  int n_[3] = {0, 0, 0};
  int cell_of(float v, int a) const { return (int)std::floor((v - lo_[a]) / cell_); }
 public:
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) {
    cloud_ = c;
    start_.clear(); idx_.clear();
    n_[0] = n_[1] = n_[2] = 0;
    const size_t np = c->points.size();
    if (!np) return;
    float hi[3] = {c->points[0].x, c->points[0].y, c->points[0].z};
    lo_[0] = hi[0]; lo_[1] = hi[1]; lo_[2] = hi[2];
    for (const P& p : c->points) {
      const float v[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; a++) { if (v[a] < lo_[a]) lo_[a] = v[a]; if (v[a] > hi[a]) hi[a] = v[a]; }
    }
    for (int a = 0; a < 3; a++) n_[a] = cell_of(hi[a], a) + 1;
    const size_t nc = (size_t)n_[0] * n_[1] * n_[2];
    start_.assign(nc + 1, 0);
    auto id = [&](const P& p) { return ((size_t)cell_of(p.x, 0) * n_[1] + cell_of(p.y, 1)) * n_[2] + cell_of(p.z, 2); };
    for (const P& p : c->points) start_[id(p) + 1]++;
    for (size_t i = 0; i < nc; i++) start_[i + 1] += start_[i];
    idx_.resize(np);
    std::vector<int> at(start_.begin(), start_.end() - 1);
    for (size_t i = 0; i < np; i++) idx_[at[id(c->points[i])]++] = (int)i;
  }
  int radiusSearch(const P& q, double radius, std::vector<int>& idx, std::vector<float>& d2, unsigned = 0) const {
    idx.clear(); d2.clear();
    if (!cloud_ || !n_[0]) return 0;
    const float r = (float)radius, r2 = r * r;
    const float qv[3] = {q.x, q.y, q.z};
    int c0[3], c1[3];
    for (int a = 0; a < 3; a++) {
      c0[a] = std::max(cell_of(qv[a] - r, a), 0);
      c1[a] = std::min(cell_of(qv[a] + r, a), n_[a] - 1);
      if (c0[a] > c1[a]) return 0;
    }
    for (int x = c0[0]; x <= c1[0]; x++) for (int y = c0[1]; y <= c1[1]; y++) {
      const size_t row = ((size_t)x * n_[1] + y) * n_[2];
      for (int k = start_[row + c0[2]]; k < start_[row + c1[2] + 1]; k++) {
        const int i = idx_[k];
        const P& p = cloud_->points[i];
        const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        const float d = dx * dx + dy * dy + dz * dz;
        if (d <= r2) { idx.push_back(i); d2.push_back(d); }
      }
    }
    return (int)idx.size();
  }
};
}  // namespace pcl

// sensor_msgs::PointCloud2 + pcl::fromROSMsg (pcl_conversions): the message simply carries the points
namespace sensor_msgs {
struct PointCloud2 {
  typedef boost::shared_ptr<const PointCloud2> ConstPtr;
  std::vector<pcl::PointXYZ> pts;
};
typedef boost::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
}  // namespace sensor_msgs
namespace pcl {
inline void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointXYZ>& cloud) {
  cloud.points = msg.pts; cloud.width = (uint32_t)msg.pts.size(); cloud.height = 1;
}
}  // namespace pcl
