// include/uavmp/kino_astar.hpp — C++ shim with the reference's class interface over the C-ABI (include/uavmp.h).
//
// Drop-in for path_searching::KinoAstar (reference: src/planner/path_searching/include/path_searching/kino_astar.h:117-206):
// same method names, argument meaning, return codes and ownership rules (search() only push_back's into `path`, the caller
// clears it — kino_astar.cpp:509,533; test_kino_astar_searching.cpp:69).  Compiled only where Eigen exists (the caller's
// tree); nothing in this repository includes it.  Instead of ros::NodeHandle / GridMap::Ptr the shim takes the values those
// objects would provide: the 12 parameters by the same names, and the GridMap's inflated buffer + geometry + the cloud.
#pragma once
#include <Eigen/Eigen>
#include <memory>
#include <stdexcept>
#include <vector>

#include "../uavmp.h"

namespace uavmp {

class KinoAstar {
 public:
  typedef std::shared_ptr<KinoAstar> Ptr;
  enum { REACH_END = UAVMP_REACH_END, NO_PATH_FOUND = UAVMP_NO_PATH_FOUND };  // kino_astar.h:155-159

  explicit KinoAstar(int device = 0) {
    if (uavmp_ctx_create(&ctx_, device) != UAVMP_OK) throw std::runtime_error("uavmp: no CUDA device (no CPU fallback)");
    uavmp_kino_params_default(&params_);  // the C++ defaults of KinoAstar::setParam (kino_astar.cpp:8-19)
  }
  ~KinoAstar() { uavmp_ctx_destroy(ctx_); }
  KinoAstar(const KinoAstar&) = delete;
  KinoAstar& operator=(const KinoAstar&) = delete;

  // void setParam(ros::NodeHandle& nh): pass the values nh.param("kino_astar/...") would yield
  void setParam(const uavmp_kino_params& p) { params_ = p; check(uavmp_kino_set_params(ctx_, &params_)); }
  uavmp_kino_params& params() { return params_; }

  // void setGridMap(GridMap::Ptr&) + localCloudCallback: md_.occupancy_buffer_inflate_, mp_.map_origin_, mp_.map_size_,
  // mp_.resolution_, mp_.map_voxel_num_ (grid_map.h) and the "local_cloud" PointCloud2 as n x 3 float32
  void setGridMap(const std::vector<char>& occupancy_buffer_inflate, const Eigen::Vector3i& map_voxel_num,
                  const Eigen::Vector3d& map_origin, const Eigen::Vector3d& map_size, double resolution,
                  const float* cloud_xyz, int n_cloud) {
    check(uavmp_kino_set_params(ctx_, &params_));
    check(uavmp_map_set(ctx_, reinterpret_cast<const int8_t*>(occupancy_buffer_inflate.data()), map_voxel_num(0),
                        map_voxel_num(1), map_voxel_num(2), map_origin.data(), map_size.data(), resolution, cloud_xyz, n_cloud));
  }
  void init() {}   // pools are per-CTA arenas on the device (kino_astar.cpp:57-74)
  void reset() {}  // per-query state is reset by the kernel (kino_astar.cpp:274-300)

  // int search(Vector3d start_pt, Vector3d start_vel, Vector3d end_pt, Vector3d end_vel, std::vector<Vector3d>& path)
  int search(Eigen::Vector3d start_pt, Eigen::Vector3d start_vel, Eigen::Vector3d end_pt, Eigen::Vector3d end_vel,
             std::vector<Eigen::Vector3d>& path) {
    int status = 0;
    long long off[2] = {0, 0};
    long long n = uavmp_kino_search_batch(ctx_, 1, start_pt.data(), start_vel.data(), end_pt.data(), end_vel.data(), &status,
                                          nullptr, off, nullptr, nullptr);
    if (n < 0) throw std::runtime_error(uavmp_last_error(ctx_));
    std::vector<double> xyz(3 * (size_t)(n > 0 ? n : 1));
    check(uavmp_kino_get_paths(ctx_, xyz.data(), n > 0 ? n : 1));
    for (long long i = 0; i < n; i++) path.push_back(Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    return status;
  }

  // batched form (what the B200 path is for): B queries, row-major B x 3
  long long searchBatch(int B, const double* start_pt, const double* start_vel, const double* end_pt, const double* end_vel,
                        int* status, long long* path_offsets) {
    return uavmp_kino_search_batch(ctx_, B, start_pt, start_vel, end_pt, end_vel, status, nullptr, path_offsets, nullptr, nullptr);
  }
  uavmp_ctx* context() { return ctx_; }

 private:
  void check(int rc) { if (rc < 0) throw std::runtime_error(uavmp_last_error(ctx_)); }
  uavmp_ctx* ctx_ = nullptr;
  uavmp_kino_params params_;
};

}  // namespace uavmp
