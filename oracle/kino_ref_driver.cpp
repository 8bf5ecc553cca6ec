// oracle/kino_ref_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's OWN path_searching::KinoAstar,
// compiled UNMODIFIED from /root/reference/src/planner/path_searching/src/kino_astar.cpp against the header shims in
// oracle/shim/ (no ROS / Eigen / PCL in this image) into oracle/_ref/libkino_ref.so (recipe: oracle/Makefile).
// It exists to pin the restatement oracle/kino_ref.cpp to the reference's control flow: tests/test_kino_reference_build.py
// compares status, use_node_num_, every sampled path point (bit for bit) and a digest of all GridMap::isInMap arguments
// (= the ordered expansion sequence) between the two on the golden queries and on random ones.
//
// A fresh KinoAstar object serves every query (setParam -> setGridMap -> init -> cloud -> search), so the reference's
// reset() (whose path_node_pool_.clear() makes later indexing undefined) is never needed.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

// access to use_node_num_ for the comparison: only this translation unit sees the members as public; kino_astar.cpp itself is
// compiled as it stands (access specifiers do not change the object layout)
#define private public
#include <path_searching/kino_astar.h>
#undef private

extern "C" {

typedef struct {
  int allocated_node_num, collision_check_type;
  double rou_time, lambda_heu, goal_tolerance, time_step_size, max_velocity, max_acceleration, acc_resolution, sample_tau;
  double robot_r, robot_h;
} refkino_params;

typedef struct {
  int status, use_node_num, n_path, pad;
  unsigned long long lookup_digest;
  long long n_in_map_calls, n_occ_lookup;
} refkino_result;

struct refkino {
  refkino_params p;
  GridMap::Ptr map;
  boost::shared_ptr<sensor_msgs::PointCloud2> cloud;
};

refkino* refkino_create(const refkino_params* p, const int8_t* occ_inflate, int nx, int ny, int nz, const double origin[3],
                        const double map_size[3], double resolution, const float* cloud_xyz, int n_cloud) {
  refkino* k = new refkino();
  k->p = *p;
  k->map = std::make_shared<GridMap>();
  GridMap& g = *k->map;
  g.occupancy_buffer_inflate_ = occ_inflate;  // caller keeps it alive
  g.map_voxel_num_ = Eigen::Vector3i(nx, ny, nz);
  g.map_origin_ = Eigen::Vector3d(origin[0], origin[1], origin[2]);
  g.map_size_ = Eigen::Vector3d(map_size[0], map_size[1], map_size[2]);
  g.map_min_boundary_ = g.map_origin_;                 // grid_map.cpp:72
  g.map_max_boundary_ = g.map_origin_ + g.map_size_;   // grid_map.cpp:73
  g.resolution_ = resolution;
  g.resolution_inv_ = 1.0 / resolution;
  k->cloud = boost::make_shared<sensor_msgs::PointCloud2>();
  k->cloud->pts.resize(n_cloud);
  for (int i = 0; i < n_cloud; i++) { k->cloud->pts[i].x = cloud_xyz[3 * i]; k->cloud->pts[i].y = cloud_xyz[3 * i + 1]; k->cloud->pts[i].z = cloud_xyz[3 * i + 2]; }
  return k;
}

void refkino_destroy(refkino* k) { delete k; }

int refkino_search(refkino* k, const double sp[3], const double sv[3], const double ep[3], const double ev[3], refkino_result* res,
                   double* path_xyz, int path_cap) {
  // This library carries its own (statically linked) libstdc++ when built by the image's g++ wrapper; loaded with RTLD_LOCAL
  // from Python nobody has constructed that copy's std::cout yet, and the reference prints to it.
  static std::ios_base::Init iostreams_ready;
  std::ostringstream sink;                       // the reference prints to std::cout; keep the test output clean
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  std::streambuf* olde = std::cerr.rdbuf(sink.rdbuf());
  int status;
  {
    ros::NodeHandle nh;
    const refkino_params& p = k->p;
    nh.values["kino_astar/allocated_node_num"] = p.allocated_node_num;
    nh.values["kino_astar/collision_check_type"] = p.collision_check_type;
    nh.values["kino_astar/rou_time"] = p.rou_time;
    nh.values["kino_astar/lambda_heu"] = p.lambda_heu;
    nh.values["kino_astar/goal_tolerance"] = p.goal_tolerance;
    nh.values["kino_astar/time_step_size"] = p.time_step_size;
    nh.values["kino_astar/max_velocity"] = p.max_velocity;
    nh.values["kino_astar/max_accelration"] = p.max_acceleration;
    nh.values["kino_astar/acc_resolution"] = p.acc_resolution;
    nh.values["kino_astar/sample_tau"] = p.sample_tau;
    nh.values["kino_se3/robot_r"] = p.robot_r;
    nh.values["kino_se3/robot_h"] = p.robot_h;
    path_searching::KinoAstar ka;
    ka.setParam(nh);
    ka.setGridMap(k->map);
    ka.init();
    nh.deliver<sensor_msgs::PointCloud2>("local_cloud", k->cloud);   // -> KinoAstar::localCloudCallback
    GridMap& g = *k->map;
    g.lookup_digest = 0xcbf29ce484222325ull; g.n_in_map_calls = 0; g.n_occ_lookup = 0;
    std::vector<Eigen::Vector3d> path;
    status = ka.search(Eigen::Vector3d(sp[0], sp[1], sp[2]), Eigen::Vector3d(sv[0], sv[1], sv[2]), Eigen::Vector3d(ep[0], ep[1], ep[2]),
                       Eigen::Vector3d(ev[0], ev[1], ev[2]), path);
    res->status = status;
    res->use_node_num = ka.use_node_num_;
    res->n_path = (int)path.size();
    res->lookup_digest = g.lookup_digest;
    res->n_in_map_calls = (long long)g.n_in_map_calls;
    res->n_occ_lookup = (long long)g.n_occ_lookup;
    for (int i = 0; i < (int)path.size() && i < path_cap; i++) { path_xyz[3 * i] = path[i](0); path_xyz[3 * i + 1] = path[i](1); path_xyz[3 * i + 2] = path[i](2); }
  }
  std::cout.rdbuf(old);
  std::cerr.rdbuf(olde);
  return status;
}

}  // extern "C"
