// oracle/astar_ref.cpp — TEST INFRASTRUCTURE ONLY (tests/, smoke, bench CPU legs; see oracle/README.md).
//
// CPU restatement of path_searching::Astar (reference: src/planner/path_searching/src/a_star.cpp:48-154 search, :161-170
// getDiagonalHeu, :180-190 retrievePath; include/path_searching/a_star.h:19-104), bug for bug, with std::priority_queue /
// std::unordered_map exactly as the reference uses them but without Eigen / ROS:
//   * nodes are keyed by their exact position (a_star.h:66-69: unordered_map<Vector3d, ...>, equality = component-wise ==);
//   * 27 neighbour offsets from `for (x = -res; x <= res; x += res)`, the centre included (:91-93);
//   * a smaller g on an expanded, not yet closed node overwrites g / parent / f in place with no re-heapify (:141-146);
//   * pool check in the middle of an expansion (:134-138); per-axis goal test on pop (:77-89); end point outside the map
//     returns at once (:52-56).
// Assumption (cannot be read off the reference alone): the unqualified `abs` of the goal test (:77-79) resolves to the double
// overload — <stdlib.h> / <math.h> reach a_star.cpp through ROS / PCL / OpenCV headers (libstdc++'s C++ wrappers put
// std::abs's overloads into the global namespace); oracle/shim/plan_env/grid_map.h includes them for the same reason.
// PARITY STATUS: pinned by the reference's own a_star.cpp compiled unmodified against the header shims
// (oracle/_ref/libastar_ref.so, tests/test_astar_reference_build.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <queue>
#include <unordered_map>
#include <vector>

namespace {

struct P3 {
  double x, y, z;
  bool operator==(const P3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct P3Hash {  // a_star.h:46-60 (any hash gives the same find / insert results)
  size_t operator()(const P3& p) const {
    size_t seed = 0;
    const double v[3] = {p.x, p.y, p.z};
    for (int i = 0; i < 3; i++) seed ^= std::hash<double>()(v[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
  }
};
struct ANode {
  P3 position;
  double g_cost = (double)(1 << 30), f_cost = (double)(1 << 30);
  char node_state = 'c';
  ANode* parent = nullptr;
};
struct ACmp { bool operator()(const ANode* a, const ANode* b) const { return a->f_cost > b->f_cost; } };

uint64_t bits(double x) { if (x == 0.0) x = 0.0; uint64_t u; std::memcpy(&u, &x, 8); return u; }
uint64_t mix(uint64_t h, uint64_t v) { h ^= v; h *= 0x100000001b3ull; h ^= h >> 29; return h; }

}  // namespace

extern "C" {

typedef struct {
  int status, use_node_num, n_pop, n_path;
  unsigned long long pop_hash;
  unsigned long long lookup_digest;  // digest of every position the search itself passes to GridMap::isInMap, in call order (the
  long long n_in_map_calls;          // fingerprint oracle/shim/plan_env/grid_map.h takes of the reference build)
} oracle_astar_result;

// one fresh Astar per query (setParam -> setGridMap -> init -> search)
int oracle_astar_search(double lambda_heu, int allocated_node_num, const int8_t* occ_inflate, int nx, int ny, int nz,
                        const double origin[3], const double map_size[3], double resolution, const double start_pt[3],
                        const double end_pt[3], oracle_astar_result* res, double* path_xyz, int path_cap) {
  const double tie_breaker = 1.0 + 1e-4;  // a_star.cpp:16
  const double bmin[3] = {origin[0], origin[1], origin[2]};
  const double bmax[3] = {origin[0] + map_size[0], origin[1] + map_size[1], origin[2] + map_size[2]};  // grid_map.cpp:72-73
  const double inv_res = 1.0 / resolution;
  unsigned long long digest = 0xcbf29ce484222325ull;
  long long n_calls = 0;
  auto in_map_raw = [&](const P3& p) {  // grid_map.h:370-385
    if (p.x < bmin[0] + 1e-4 || p.y < bmin[1] + 1e-4 || p.z < bmin[2] + 1e-4) return false;
    if (p.x > bmax[0] - 1e-4 || p.y > bmax[1] - 1e-4 || p.z > bmax[2] - 1e-4) return false;
    return true;
  };
  auto isInMap = [&](const P3& p) {  // the search's own calls are fingerprinted
    const double c[3] = {p.x, p.y, p.z};
    for (int i = 0; i < 3; i++) { uint64_t u; std::memcpy(&u, &c[i], 8); digest ^= u; digest *= 0x100000001b3ull; digest ^= digest >> 29; }
    n_calls++;
    return in_map_raw(p);
  };
  auto occupancy = [&](const P3& p) {  // getInflateOccupancy, grid_map.h:350-359
    if (!in_map_raw(p)) return -1;
    const int ix = (int)std::floor((p.x - origin[0]) * inv_res), iy = (int)std::floor((p.y - origin[1]) * inv_res),
              iz = (int)std::floor((p.z - origin[2]) * inv_res);
    return (int)occ_inflate[(size_t)ix * ny * nz + (size_t)iy * nz + iz];
  };
  auto heu = [&](const P3& a, const P3& b) {  // getDiagonalHeu :161-170
    const double dx = std::abs(a.x - b.x), dy = std::abs(a.y - b.y), dz = std::abs(a.z - b.z);
    const double mn = std::min({dx, dy, dz});
    const double h = dx + dy + dz + (std::sqrt(3) - 3) * mn;
    return tie_breaker * h;
  };
  res->status = 2; res->use_node_num = 0; res->n_pop = 0; res->n_path = 0; res->pop_hash = 0xcbf29ce484222325ull;
  res->lookup_digest = digest; res->n_in_map_calls = 0;
  const P3 start{start_pt[0], start_pt[1], start_pt[2]}, end{end_pt[0], end_pt[1], end_pt[2]};
  if (!isInMap(end)) { res->lookup_digest = digest; res->n_in_map_calls = n_calls; return 2; }  // :52-56
  std::vector<ANode> pool(allocated_node_num);
  std::priority_queue<ANode*, std::vector<ANode*>, ACmp> open_list;
  std::unordered_map<P3, ANode*, P3Hash> close_list, expanded;
  int use_node_num = 0;
  ANode* s = &pool[use_node_num];
  s->g_cost = 0.0; s->position = start; s->parent = nullptr;
  s->f_cost = lambda_heu * heu(start, end);
  s->node_state = 'a';
  open_list.push(s);
  expanded.insert({s->position, s});
  use_node_num += 1;
  uint64_t ph = res->pop_hash;
  int n_pop = 0;
  auto finish = [&](int st) {
    res->status = st; res->use_node_num = use_node_num; res->n_pop = n_pop; res->pop_hash = ph;
    res->lookup_digest = digest; res->n_in_map_calls = n_calls;
    return st;
  };
  while (!open_list.empty()) {
    ANode* cur = open_list.top();
    open_list.pop();
    cur->node_state = 'b';
    close_list.insert({cur->position, cur});
    n_pop++;
    ph = mix(ph, bits(cur->position.x)); ph = mix(ph, bits(cur->position.y)); ph = mix(ph, bits(cur->position.z));
    { uint64_t u; std::memcpy(&u, &cur->g_cost, 8); ph = mix(ph, u); }
    if (std::abs(cur->position.x - end.x) < resolution && std::abs(cur->position.y - end.y) < resolution &&
        std::abs(cur->position.z - end.z) < resolution) {
      std::vector<P3> path;  // retrievePath :180-190
      ANode* c = cur;
      while (c->parent != nullptr) { path.push_back(c->position); c = c->parent; }
      path.push_back(c->position);
      std::reverse(path.begin(), path.end());
      res->n_path = (int)path.size();
      for (int i = 0; i < (int)path.size() && i < path_cap; i++) { path_xyz[3 * i] = path[i].x; path_xyz[3 * i + 1] = path[i].y; path_xyz[3 * i + 2] = path[i].z; }
      return finish(1);
    }
    for (double x = -resolution; x <= resolution; x += resolution)
      for (double y = -resolution; y <= resolution; y += resolution)
        for (double z = -resolution; z <= resolution; z += resolution) {
          const P3 np{cur->position.x + x, cur->position.y + y, cur->position.z + z};
          if (!isInMap(np)) continue;
          if (occupancy(np) == 1) continue;  // == true (:106)
          if (close_list.find(np) != close_list.end()) continue;
          const double delta_pos = std::sqrt((x * x + y * y) + z * z);  // Vector3d(x, y, z).norm(), Eigen's 3-vector reduction
          const double tmp_g = cur->g_cost + delta_pos;
          auto it = expanded.find(np);
          if (it == expanded.end()) {
            ANode* nb = &pool[use_node_num];
            use_node_num += 1;
            nb->g_cost = tmp_g; nb->position = np; nb->parent = cur;
            nb->f_cost = nb->g_cost + lambda_heu * heu(np, end);
            nb->node_state = 'a';
            open_list.push(nb);
            expanded.insert({nb->position, nb});
            if (use_node_num >= allocated_node_num) return finish(2);  // :134-138
          } else if (tmp_g < it->second->g_cost) {
            ANode* t = it->second;
            t->g_cost = tmp_g; t->parent = cur;
            t->f_cost = t->g_cost + lambda_heu * heu(t->position, end);
          }
        }
  }
  return finish(2);
}

}  // extern "C"
