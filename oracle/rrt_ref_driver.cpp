// oracle/rrt_ref_driver.cpp — TEST INFRASTRUCTURE.  C entry point around the reference's OWN path_searching::RRTStar, compiled
// UNMODIFIED from /root/reference/src/planner/path_searching/src/rrt_star.cpp and src/kdtree/kdtree.cpp against the header shims in
// oracle/shim/ into oracle/_ref/librrt_ref.so (recipe: oracle/Makefile).  The reference search is not reproducible as written: every
// sample comes from a fresh std::random_device and the loop ends on wall-clock time.  Both are pinned here WITHOUT touching the sources:
// shim/rrt_seeded_random.h redirects `random_device` to a counter-based seed stream, and ros::Time::now() (shim) returns the number of
// samples drawn so far, so `rrt_star/max_tolerance_time` becomes a sample budget.  tests/test_rrt_star_reference_build.py compares
// status, use_node_num_, the number of samples, getOptimalPath() and a digest of the whole tree (position, parent, g_cost of every
// node) with the restatement oracle/rrt_star_ref.cpp.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <unordered_map>
#include <vector>

#define private public
#include <path_searching/rrt_star.h>
#undef private

extern "C" {
unsigned long long rrt_shim_query_seed = 0;
long long rrt_shim_samples = 0;
}
static double rrt_clock() { return (double)rrt_shim_samples; }

extern "C" {

typedef struct {
  int status, use_node_num, n_opt_path, reach_goal;
  long long n_samples;
  unsigned long long tree_digest;
  double goal_g_cost;
} refrrt_result;

static inline void fold(unsigned long long& h, unsigned long long v) { h ^= v; h *= 0x100000001b3ull; h ^= h >> 29; }
static inline unsigned long long bits(double d) { unsigned long long b; std::memcpy(&b, &d, 8); return b; }

int refrrt_search(int max_tree_node_num, double step_length, double search_radius, double collision_check_resolution,
                  double sample_budget, unsigned long long query_seed, const int8_t* occ_inflate, int nx, int ny, int nz,
                  const double origin[3], const double map_size[3], double resolution, const double sp[3], const double ep[3],
                  refrrt_result* res, double* opt_path_xyz, int path_cap) {
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
    nh.values["rrt_star/max_tree_node_num"] = max_tree_node_num;
    nh.values["rrt_star/step_length"] = step_length;
    nh.values["rrt_star/search_radius"] = search_radius;
    nh.values["rrt_star/collision_check_resolution"] = collision_check_resolution;
    nh.values["rrt_star/max_tolerance_time"] = sample_budget;
    rrt_shim_query_seed = query_seed;
    rrt_shim_samples = 0;
    ros::time_hook() = rrt_clock;
    path_searching::RRTStar rrt;
    rrt.setParam(nh);
    rrt.setGridMap(map);
    rrt.init();
    std::vector<Eigen::Vector3d> path;
    status = rrt.search(Eigen::Vector3d(sp[0], sp[1], sp[2]), Eigen::Vector3d(ep[0], ep[1], ep[2]), path);
    ros::time_hook() = nullptr;
    std::vector<Eigen::Vector3d> opt = rrt.getOptimalPath();
    res->status = status;
    res->use_node_num = rrt.use_node_num_;
    res->n_opt_path = (int)opt.size();
    res->reach_goal = rrt.reach_goal_ ? 1 : 0;
    res->n_samples = rrt_shim_samples;
    res->goal_g_cost = rrt.path_node_pool_[1]->g_cost;
    std::unordered_map<const path_searching::RRTStarNode*, int> index;
    for (int i = 0; i < rrt.use_node_num_; i++) index[rrt.path_node_pool_[i]] = i;
    unsigned long long h = 0;
    for (int i = 0; i < rrt.use_node_num_; i++) {
      const path_searching::RRTStarNode* n = rrt.path_node_pool_[i];
      unsigned long long hn = 0xcbf29ce484222325ull ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1));
      for (int k = 0; k < 3; k++) fold(hn, bits(n->position(k)));
      fold(hn, bits(n->g_cost));
      fold(hn, n->parent ? (unsigned long long)index.at(n->parent) : 0xffffffffull);
      h += hn;
    }
    res->tree_digest = h;
    for (int i = 0; i < (int)opt.size() && i < path_cap; i++) { opt_path_xyz[3 * i] = opt[i](0); opt_path_xyz[3 * i + 1] = opt[i](1); opt_path_xyz[3 * i + 2] = opt[i](2); }
  }
  std::cout.rdbuf(old);
  std::cerr.rdbuf(olde);
  return status;
}

}  // extern "C"
