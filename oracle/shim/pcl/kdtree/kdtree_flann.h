// oracle/shim/pcl/kdtree/kdtree_flann.h — TEST INFRASTRUCTURE: pcl::PointXYZ / PointCloud / KdTreeFLANN::radiusSearch without PCL.
// radiusSearch returns every point whose float squared distance is <= radius^2, found through a uniform cell grid.  The
// reference only asks "is ANY returned point inside the ellipsoid" (kino_astar.cpp:744-753) and its radius (robot_r + 0.1)
// exceeds the longest semi-axis, so neither the order of the result nor points at the rim of the ball can change the answer
// (SURVEY.md §9.1 Q12).
#pragma once
#include <boost/make_shared.hpp>
#include <cmath>
#include <cstdint>
#include <unordered_map>
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
  std::unordered_map<long long, std::vector<int>> cells_;
  float cell_ = 0.5f;
  static long long key(int x, int y, int z) { return ((long long)(x + (1 << 20)) << 42) | ((long long)(y + (1 << 20)) << 21) | (long long)(z + (1 << 20)); }
 public:
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) {
    cloud_ = c;
    cells_.clear();
    for (size_t i = 0; i < c->points.size(); i++) {
      const P& p = c->points[i];
      cells_[key((int)std::floor(p.x / cell_), (int)std::floor(p.y / cell_), (int)std::floor(p.z / cell_))].push_back((int)i);
    }
  }
  int radiusSearch(const P& q, double radius, std::vector<int>& idx, std::vector<float>& d2, unsigned = 0) const {
    idx.clear(); d2.clear();
    if (!cloud_) return 0;
    const float r = (float)radius, r2 = r * r;
    const int x0 = (int)std::floor((q.x - r) / cell_), x1 = (int)std::floor((q.x + r) / cell_);
    const int y0 = (int)std::floor((q.y - r) / cell_), y1 = (int)std::floor((q.y + r) / cell_);
    const int z0 = (int)std::floor((q.z - r) / cell_), z1 = (int)std::floor((q.z + r) / cell_);
    for (int x = x0; x <= x1; x++) for (int y = y0; y <= y1; y++) for (int z = z0; z <= z1; z++) {
      auto it = cells_.find(key(x, y, z));
      if (it == cells_.end()) continue;
      for (int i : it->second) {
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
