// include/uavmp/a_star.hpp — C++ shim with the reference's class interface over the C-ABI (include/uavmp.h).
//
// Drop-in for path_searching::Astar (reference: src/planner/path_searching/include/path_searching/a_star.h:106-152): same
// method names, argument meaning and return codes (REACH_END = 1, NO_PATH = 2, a_star.h:122-126).  search() fills `path` the
// way Astar::search does through retrievePath (a_star.cpp:80-86,180-190): start point first.  Compiled only where Eigen
// exists; nothing in this repository includes it except the compile check in tests/host/shim_check.cpp.
#pragma once
#include <Eigen/Eigen>
#include <memory>
#include <stdexcept>
#include <vector>

#include "../uavmp.h"

namespace uavmp {

class Astar {
 public:
  typedef std::shared_ptr<Astar> Ptr;
  enum { REACH_END = 1, NO_PATH = 2 };

  // shares a context (and therefore the map) with a KinoAstar when one is passed in
  explicit Astar(uavmp_ctx* shared = nullptr, int device = 0) : ctx_(shared), owned_(shared == nullptr) {
    if (owned_ && uavmp_ctx_create(&ctx_, device) != UAVMP_OK) throw std::runtime_error("uavmp: no CUDA device (no CPU fallback)");
  }
  ~Astar() { if (owned_) uavmp_ctx_destroy(ctx_); }
  Astar(const Astar&) = delete;
  Astar& operator=(const Astar&) = delete;

  // void setParam(ros::NodeHandle& nh): astar/lambda_heu, astar/allocated_node_num (a_star.cpp:3-11; astar/resolution is
  // overwritten by the map's in init, :17)
  void setParam(double lambda_heu = 1.0, int allocated_node_num = 100000, int path_cap_nodes = 4096) {
    check(uavmp_astar_set_params(ctx_, lambda_heu, allocated_node_num, path_cap_nodes));
  }
  // void setGridMap(GridMap::Ptr&): the inflated buffer and geometry GridMap exposes (grid_map.h:350-385)
  void setGridMap(const std::vector<char>& occupancy_buffer_inflate, const Eigen::Vector3i& map_voxel_num,
                  const Eigen::Vector3d& map_origin, const Eigen::Vector3d& map_size, double resolution) {
    check(uavmp_map_set(ctx_, reinterpret_cast<const int8_t*>(occupancy_buffer_inflate.data()), map_voxel_num(0),
                        map_voxel_num(1), map_voxel_num(2), map_origin.data(), map_size.data(), resolution, nullptr, 0));
  }
  void init() {}   // node pools are per-warp arenas on the device (a_star.cpp:13-38)
  void reset() {}  // per-query state is reset by the kernel (a_star.cpp:196-213)

  // int search(Eigen::Vector3d start_pt, Eigen::Vector3d end_pt, std::vector<Eigen::Vector3d>& path)   (a_star.h:147)
  int search(Eigen::Vector3d start_pt, Eigen::Vector3d end_pt, std::vector<Eigen::Vector3d>& path) {
    int status = 0;
    long long off[2] = {0, 0};
    long long n = uavmp_astar_search_batch(ctx_, 1, start_pt.data(), end_pt.data(), &status, nullptr, off, nullptr, nullptr);
    if (n < 0) throw std::runtime_error(uavmp_last_error(ctx_));
    std::vector<double> xyz(3 * (size_t)(n > 0 ? n : 1));
    check(uavmp_astar_get_paths(ctx_, xyz.data(), n > 0 ? n : 1));
    for (long long i = 0; i < n; i++) path.push_back(Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    return status;
  }
  long long searchBatch(int B, const double* start_pt, const double* end_pt, int* status, long long* path_offsets) {
    return uavmp_astar_search_batch(ctx_, B, start_pt, end_pt, status, nullptr, path_offsets, nullptr, nullptr);
  }
  uavmp_ctx* context() { return ctx_; }

 private:
  void check(int rc) { if (rc < 0) throw std::runtime_error(uavmp_last_error(ctx_)); }
  uavmp_ctx* ctx_ = nullptr;
  bool owned_;
};

}  // namespace uavmp
