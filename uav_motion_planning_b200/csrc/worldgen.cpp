// worldgen.cpp — synthetic world generator used by tests and bench.py (host utility, not the hot path).
//
// It DEFINES the inputs both the oracle and the CUDA path consume:
//   1. global cloud  : the obstacle generators of the reference simulator
//                      (src/simulator/map_generator/src/random_forest.cpp:55-155 random pillars + rings,
//                       :286-306,:347-351 the two-slab wall map), float32 points;
//   2. sensed cloud  : one float32 centroid per occupied 0.1 m leaf, sorted by leaf
//                      (stands in for pcl::VoxelGrid at pointcloud_render_node.cpp:84-86);
//   3. inflated grid : GridMap::cloudCallback's 3x3x3 stamping rule
//                      (src/planner/plan_env/src/grid_map.cpp:733-785) into an x-major / z-fastest int8 grid
//                      (grid_map.h:257-260).
// The RNG is minstd_rand0 + the 2-draw generate_canonical<double,53> that libstdc++'s
// std::default_random_engine / uniform_real_distribution use, written out so the maps do not depend on the
// standard library in use.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "uavmp_worldgen.h"

namespace {

struct Rng {
  uint32_t s;
  explicit Rng(uint32_t seed) { s = seed % 2147483647u; if (s == 0) s = 1; }
  uint32_t next() { s = (uint32_t)(((uint64_t)s * 16807u) % 2147483647u); return s; }
  double canonical() {
    const double r = 2147483646.0;
    double sum = (double)(next() - 1u);
    sum += (double)(next() - 1u) * r;
    double ret = sum / (r * r);
    if (ret >= 1.0) ret = std::nextafter(1.0, 0.0);
    return ret;
  }
  double uniform(double a, double b) { return canonical() * (b - a) + a; }
};

struct P3f { float x, y, z; };

void gen_random_forest(const uavmp_mapgen_params& mp, std::vector<P3f>& cloud) {
  Rng eng(mp.seed);
  const double res = mp.resolution;
  const double xl = -mp.x_size / 2.0, xh = mp.x_size / 2.0, yl = -mp.y_size / 2.0, yh = mp.y_size / 2.0;
  auto clear_ok = [&](double x, double y) {
    double dx = x - mp.init_x, dy = y - mp.init_y;
    return std::sqrt(dx * dx + dy * dy) >= mp.init_radius;
  };
  // pillars (random_forest.cpp:68-100)
  for (int i = 0; i < mp.polar_num; i++) {
    double x = eng.uniform(xl, xh), y = eng.uniform(yl, yh), w = eng.uniform(mp.w_l, mp.w_h);
    if (!clear_ok(x, y)) { i--; continue; }
    x = std::floor(x / res) * res + res / 2.0;
    y = std::floor(y / res) * res + res / 2.0;
    int num_w = (int)std::ceil(w / res);
    for (int r = (int)(-num_w / 2.0); r < num_w / 2.0; r++)
      for (int s = (int)(-num_w / 2.0); s < num_w / 2.0; s++) {
        double h = eng.uniform(mp.h_l, mp.h_h);
        int num_h = (int)std::ceil(h / res);
        for (int t = -20; t < num_h; t++) {
          P3f p;
          p.x = (float)(x + (r + 0.5) * res + 1e-2);
          p.y = (float)(y + (s + 0.5) * res + 1e-2);
          p.z = (float)((t + 0.5) * res + 1e-2);
          cloud.push_back(p);
        }
      }
  }
  // rings (random_forest.cpp:103-154)
  for (int i = 0; i < mp.circle_num; i++) {
    double x = eng.uniform(xl, xh), y = eng.uniform(yl, yh), z = eng.uniform(mp.z_l, mp.z_h);
    double r1 = eng.uniform(mp.radius_l, mp.radius_h), r2 = eng.uniform(mp.radius_l, mp.radius_h);
    double th = eng.uniform(-mp.theta, mp.theta);
    if (!clear_ok(x, y)) { i--; continue; }
    x = std::floor(x / res) * res + res / 2.0;
    y = std::floor(y / res) * res + res / 2.0;
    z = std::floor(z / res) * res + res / 2.0;
    double c = std::cos(th), s = std::sin(th);
    for (double ang = 0.0; ang < M_PI * 2; ang += res / 2) {
      double cy = r1 * std::cos(ang), cz = r2 * std::sin(ang);
      P3f p;
      p.x = (float)((c * 0.0 + (-s) * cy) + x);
      p.y = (float)((s * 0.0 + c * cy) + y);
      p.z = (float)(cz + z);
      cloud.push_back(p);
    }
  }
}

void gen_wall(double res, double x_min, double x_max, double y_min, double y_max, double z_min, double z_max,
              std::vector<P3f>& cloud) {
  int nx = (int)std::ceil((x_max - x_min) / res), ny = (int)std::ceil((y_max - y_min) / res),
      nz = (int)std::ceil((z_max - z_min) / res);
  for (int i = 0; i < nx; i++)
    for (int j = 0; j < ny; j++)
      for (int k = 0; k < nz; k++) {
        P3f p;
        p.x = (float)(x_min + i * res);
        p.y = (float)(y_min + j * res);
        p.z = (float)(z_min + k * res);
        cloud.push_back(p);
      }
}

// one centroid per occupied leaf, leaves ordered (z, y, x)-major like pcl::VoxelGrid's index sort
void voxel_downsample(const std::vector<P3f>& in, float leaf, std::vector<P3f>& out) {
  struct Acc { float sx, sy, sz; int n; };
  std::map<std::array<int, 3>, Acc> cells;  // key = (iz, iy, ix)
  const float inv = 1.0f / leaf;
  for (const P3f& p : in) {
    std::array<int, 3> k = {(int)std::floor(p.z * inv), (int)std::floor(p.y * inv), (int)std::floor(p.x * inv)};
    Acc& a = cells[k];
    a.sx += p.x; a.sy += p.y; a.sz += p.z; a.n += 1;
  }
  out.clear();
  out.reserve(cells.size());
  for (auto& kv : cells) {
    const Acc& a = kv.second;
    out.push_back(P3f{a.sx / (float)a.n, a.sy / (float)a.n, a.sz / (float)a.n});
  }
}

}  // namespace

extern "C" {

// simulator.xml:16-41 with the obstacle counts the code actually reads (random_forest.cpp:519 reads
// map/polar_num, default 30; circle_num 50), scaled by area relative to the 20 x 20 m case
void uavmp_mapgen_params_default(uavmp_mapgen_params* p, double x_size, double y_size, uint32_t seed) {
  memset(p, 0, sizeof(*p));
  p->map_type = 0;
  p->seed = seed;
  p->x_size = x_size; p->y_size = y_size; p->resolution = 0.1;
  p->init_x = 0.0; p->init_y = 0.0; p->init_radius = 1.0;
  double scale = (x_size * y_size) / 400.0;
  p->polar_num = (int)std::lround(30 * scale);
  p->circle_num = (int)std::lround(50 * scale);
  p->w_l = 0.5; p->w_h = 0.7; p->h_l = 0.0; p->h_h = 3.0;
  p->radius_l = 0.8; p->radius_h = 0.5; p->z_l = 0.7; p->z_h = 0.8; p->theta = 0.6;
  p->wall_x = 0.0; p->wall_y = 0.0; p->wall_w = 0.5;
}

int uavmp_mapgen_cloud(const uavmp_mapgen_params* mp, float* cloud_xyz, int cap) {
  if (!mp) return -1;
  std::vector<P3f> raw, ds;
  if (mp->map_type == 0) {
    gen_random_forest(*mp, raw);
  } else if (mp->map_type == 2) {
    // random_forest.cpp:347-351 with wall_x = wall_y = 0, wall_w as given
    gen_wall(mp->resolution, mp->wall_x - 0.25, mp->wall_x + 0.25, mp->wall_y + mp->wall_w / 2.0, mp->wall_y + 20.0,
             -0.5, 4.0, raw);
    gen_wall(mp->resolution, mp->wall_x - 0.25, mp->wall_x + 0.25, mp->wall_y - 20.0, mp->wall_y - mp->wall_w / 2.0,
             -0.5, 4.0, raw);
  } else {
    return -2;
  }
  voxel_downsample(raw, 0.1f, ds);
  int n = (int)ds.size();
  if (cloud_xyz) {
    int m = std::min(n, cap);
    for (int i = 0; i < m; i++) {
      cloud_xyz[3 * i] = ds[i].x; cloud_xyz[3 * i + 1] = ds[i].y; cloud_xyz[3 * i + 2] = ds[i].z;
    }
  }
  return n;
}

// GridMap::cloudCallback inflation (grid_map.cpp:733-785), whole map in range
// (test_kino_astar_searching.launch:9-11 sets local_update_range = map size).
int uavmp_grid_inflate_host(const float* cloud_xyz, int n, const double origin[3], const double map_size[3],
                            double res, double obstacles_inflation, int8_t* occ, int nx, int ny, int nz) {
  if (!cloud_xyz || !occ) return -1;
  memset(occ, 0, (size_t)nx * ny * nz);
  const double inv = 1.0 / res;
  const int inf_step = (int)std::ceil(obstacles_inflation / res);
  const int inf_step_z = 1;
  (void)map_size;
  for (int i = 0; i < n; i++) {
    float px = cloud_xyz[3 * i], py = cloud_xyz[3 * i + 1], pz = cloud_xyz[3 * i + 2];
    for (int x = -inf_step; x <= inf_step; ++x)
      for (int y = -inf_step; y <= inf_step; ++y)
        for (int z = -inf_step_z; z <= inf_step_z; ++z) {
          double qx = px + x * res, qy = py + y * res, qz = pz + z * res;
          int ix = (int)std::floor((qx - origin[0]) * inv);
          int iy = (int)std::floor((qy - origin[1]) * inv);
          int iz = (int)std::floor((qz - origin[2]) * inv);
          if (ix < 0 || iy < 0 || iz < 0 || ix > nx - 1 || iy > ny - 1 || iz > nz - 1) continue;
          occ[(size_t)ix * ny * nz + (size_t)iy * nz + iz] = 1;
        }
  }
  return 0;
}

}  // extern "C"
