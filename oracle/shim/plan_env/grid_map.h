// oracle/shim/plan_env/grid_map.h — TEST INFRASTRUCTURE: the four GridMap members KinoAstar calls, over a caller-owned
// occupancy_buffer_inflate_.  Restated from /root/reference/src/planner/plan_env/include/plan_env/grid_map.h:
// isInMap :370-385 (strict +-1e-4 margins), posToIndex :400-404, toAddress :257-260, getInflateOccupancy :350-359,
// getResolution :443-446; boundaries as grid_map.cpp:72-73 (min = origin, max = origin + size).
// The real header cannot be used: it pulls in ROS, PCL, OpenCV and message_filters.
// Every position passed to isInMap is folded into `lookup_digest`: the sequence of map lookups of a search is a fingerprint
// of its ordered expansion sequence (each expansion starts by testing the popped node's own position, t = 0).
#pragma once
// <math.h> / <stdlib.h> (the C++ wrappers): in the reference build they arrive through ROS / PCL / OpenCV, and with them the
// global-namespace overloads of abs() that a_star.cpp:77-79 relies on (`abs(double)` without std::)
#include <math.h>
#include <stdlib.h>
#include <Eigen/Eigen>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>

class GridMap {
 public:
  typedef std::shared_ptr<GridMap> Ptr;
  const int8_t* occupancy_buffer_inflate_ = nullptr;
  Eigen::Vector3i map_voxel_num_;
  Eigen::Vector3d map_origin_, map_size_, map_min_boundary_, map_max_boundary_;
  double resolution_ = 0, resolution_inv_ = 0;
  unsigned long long lookup_digest = 0xcbf29ce484222325ull, n_in_map_calls = 0, n_occ_lookup = 0;

  void getRegion(Eigen::Vector3d& ori, Eigen::Vector3d& size) { ori = map_origin_; size = map_size_; }
  double getResolution() { return resolution_; }
  bool isInMap(const Eigen::Vector3d& pos) {
    n_in_map_calls++;
    for (int i = 0; i < 3; i++) {
      unsigned long long b;
      const double d = pos(i);
      std::memcpy(&b, &d, 8);
      lookup_digest ^= b; lookup_digest *= 0x100000001b3ull; lookup_digest ^= lookup_digest >> 29;
    }
    if (pos(0) < map_min_boundary_(0) + 1e-4 || pos(1) < map_min_boundary_(1) + 1e-4 || pos(2) < map_min_boundary_(2) + 1e-4) return false;
    if (pos(0) > map_max_boundary_(0) - 1e-4 || pos(1) > map_max_boundary_(1) - 1e-4 || pos(2) > map_max_boundary_(2) - 1e-4) return false;
    return true;
  }
  void posToIndex(const Eigen::Vector3d& pos, Eigen::Vector3i& id) {
    for (int i = 0; i < 3; ++i) id(i) = floor((pos(i) - map_origin_(i)) * resolution_inv_);
  }
  int toAddress(const Eigen::Vector3i& id) { return id(0) * map_voxel_num_(1) * map_voxel_num_(2) + id(1) * map_voxel_num_(2) + id(2); }
  int getInflateOccupancy(Eigen::Vector3d pos) {
    const unsigned long long keep_d = lookup_digest, keep_n = n_in_map_calls;
    const bool in = isInMap(pos);
    lookup_digest = keep_d; n_in_map_calls = keep_n;  // only KinoAstar's own isInMap calls are fingerprinted
    if (!in) return -1;
    Eigen::Vector3i id;
    posToIndex(pos, id);
    n_occ_lookup++;
    return int(occupancy_buffer_inflate_[toAddress(id)]);
  }
};
