// oracle/astar_ref_driver.cpp — TEST INFRASTRUCTURE.  C entry point around the reference's OWN path_searching::Astar, compiled
// UNMODIFIED from /root/reference/src/planner/path_searching/src/a_star.cpp against the header shims in oracle/shim/ into
// oracle/_ref/libastar_ref.so (recipe: oracle/Makefile).  It pins the restatement oracle/astar_ref.cpp:
// tests/test_astar_reference_build.py compares status, use_node_num_, every path point (bit for bit) and a digest of all
// GridMap::isInMap arguments (= the ordered sequence of neighbour evaluations) between the two.  A fresh Astar serves every query.
#include <cstdint>
#include <iostream>
#include <sstream>
#include <vector>

#define private public
#include <path_searching/a_star.h>
#undef private

extern "C" {

typedef struct {
  int status, use_node_num, n_path, pad;
  unsigned long long lookup_digest;
  long long n_in_map_calls;
} refastar_result;

int refastar_search(double lambda_heu, int allocated_node_num, const int8_t* occ_inflate, int nx, int ny, int nz, const double origin[3],
                    const double map_size[3], double resolution, const double sp[3], const double ep[3], refastar_result* res,
                    double* path_xyz, int path_cap) {
  static std::ios_base::Init iostreams_ready;
  std::ostringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  std::streambuf* olde = std::cerr.rdbuf(sink.rdbuf());
  int status;
  {
    GridMap::Ptr map = std::make_shared<GridMap>();
    GridMap& g = *map;
    g.occupancy_buffer_inflate_ = occ_inflate;
    g.map_voxel_num_ = Eigen::Vector3i(nx, ny, nz);
    g.map_origin_ = Eigen::Vector3d(origin[0], origin[1], origin[2]);
    g.map_size_ = Eigen::Vector3d(map_size[0], map_size[1], map_size[2]);
    g.map_min_boundary_ = g.map_origin_;
    g.map_max_boundary_ = g.map_origin_ + g.map_size_;
    g.resolution_ = resolution;
    g.resolution_inv_ = 1.0 / resolution;
    ros::NodeHandle nh;
    nh.values["astar/resolution"] = resolution;
    nh.values["astar/lambda_heu"] = lambda_heu;
    nh.values["astar/allocated_node_num"] = allocated_node_num;
    path_searching::Astar as;
    as.setParam(nh);
    as.setGridMap(map);
    as.init();
    g.lookup_digest = 0xcbf29ce484222325ull; g.n_in_map_calls = 0; g.n_occ_lookup = 0;
    std::vector<Eigen::Vector3d> path;
    status = as.search(Eigen::Vector3d(sp[0], sp[1], sp[2]), Eigen::Vector3d(ep[0], ep[1], ep[2]), path);
    res->status = status;
    res->use_node_num = as.use_node_num_;
    res->n_path = (int)path.size();
    res->lookup_digest = g.lookup_digest;
    res->n_in_map_calls = (long long)g.n_in_map_calls;
    for (int i = 0; i < (int)path.size() && i < path_cap; i++) { path_xyz[3 * i] = path[i](0); path_xyz[3 * i + 1] = path[i](1); path_xyz[3 * i + 2] = path[i](2); }
  }
  std::cout.rdbuf(old);
  std::cerr.rdbuf(olde);
  return status;
}

}  // extern "C"
