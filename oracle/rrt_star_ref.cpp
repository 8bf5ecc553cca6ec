// oracle/rrt_star_ref.cpp — TEST INFRASTRUCTURE (CPU oracle; the product never links or calls it).
//
// Restatement of path_searching::RRTStar::search and the kd-tree library it uses, for SURVEY.md §8(f) row 4 (second half):
//   reference: src/planner/path_searching/src/rrt_star.cpp:104-116 getRandomNode, :118-123 Step, :125-137 isCollisionFree,
//              :139-172 ChooseParent, :174-226 ReWireTree, :228-238 retrievePath, :304-429 search;
//              src/planner/path_searching/src/kdtree/kdtree.cpp:117-149 kd_insert, :152-182 find_nearest (range), :231-344 kd_nearest,
//              :474-496 rlist_insert (unordered results are PREPENDED: iteration runs in reverse visiting order).
// The reference search is not reproducible as written (a fresh std::random_device per sample, wall-clock termination).  The
// deterministic form pinned here and in oracle/_ref/librrt_ref.so (rrt_ref_driver.cpp): sample i is seeded by
// seed32(query_seed, i) (the one 32-bit value the reference takes from random_device per sample), and `max_tolerance_time` is a
// budget of drawn samples.  Quirks kept (each one changes the tree):
//   * isCollisionFree normalises the direction and then walks t in [0, |dir|) = [0, ~1) m in steps of the check resolution,
//     whatever the real distance between the two points (:127-135);
//   * `getInflateOccupancy(x_new) != true` accepts a point OUTSIDE the map (-1 != 1) (:341); inside isCollisionFree -1 is a hit;
//   * the goal node is pushed into the children of every node that improves it and never removed from the old ones (:362,:384),
//     so the cost propagation of a later rewire can overwrite its g_cost through a stale link, in queue order (:207-219);
//   * getOptimalPath() is only written when a LATER sample improves on the first feasible cost (:396-404): empty otherwise;
//   * `inf` is the int (1 << 30).
// Parity: tests/test_rrt_star_reference_build.py (this file vs the reference build), tests/test_rrt_star_parity.py (CUDA vs this file).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <random>
#include <vector>

extern "C" {
typedef struct {
  int status, use_node_num, n_opt_path, reach_goal;
  long long n_samples;
  unsigned long long tree_digest;
  double goal_g_cost;
  long long n_kd_visits;
} oracle_rrt_result;
}

namespace {

struct V3 { double x[3]; };
inline V3 sub(const V3& a, const V3& b) { return V3{{a.x[0] - b.x[0], a.x[1] - b.x[1], a.x[2] - b.x[2]}}; }
inline double sqn(const V3& a) { return (a.x[0] * a.x[0] + a.x[1] * a.x[1]) + a.x[2] * a.x[2]; }  // Eigen 3-vector reduction order
inline double norm(const V3& a) { return std::sqrt(sqn(a)); }
inline V3 normalized(const V3& a) {
  const double n = sqn(a);
  if (n > 0.0) { const double s = std::sqrt(n); return V3{{a.x[0] / s, a.x[1] / s, a.x[2] / s}}; }
  return a;
}

struct Map {
  const int8_t* occ; int nx, ny, nz; double o[3], mx[3], res, inv;
  int lookup(const V3& p) const {  // GridMap::getInflateOccupancy (grid_map.h:350-359) incl. isInMap (:370-385)
    for (int k = 0; k < 3; k++) if (p.x[k] < o[k] + 1e-4) return -1;
    for (int k = 0; k < 3; k++) if (p.x[k] > mx[k] - 1e-4) return -1;
    int id[3];
    for (int k = 0; k < 3; k++) id[k] = (int)std::floor((p.x[k] - o[k]) * inv);
    return occ[id[0] * ny * nz + id[1] * nz + id[2]];
  }
};

struct TreeNode {
  V3 pos; double g; int parent; std::vector<int> children;
  int dir, left, right;  // kd-tree links (the goal node is never inserted)
};

struct Rrt {
  Map m; std::vector<TreeNode> n; int use = 0; int root = -1; bool have_rect = false; double rmin[3], rmax[3];
  double step, radius, ccres;
  long long kd_visits = 0;

  void kd_insert(int id) {
    int* slot = &root; int dir = 0;
    while (*slot >= 0) {
      TreeNode& t = n[*slot];
      dir = (t.dir + 1) % 3;
      slot = (n[id].pos.x[t.dir] < t.pos.x[t.dir]) ? &t.left : &t.right;
    }
    n[id].dir = dir; n[id].left = n[id].right = -1; *slot = id;
    if (!have_rect) { for (int k = 0; k < 3; k++) rmin[k] = rmax[k] = n[id].pos.x[k]; have_rect = true; }
    else for (int k = 0; k < 3; k++) { if (n[id].pos.x[k] < rmin[k]) rmin[k] = n[id].pos.x[k]; if (n[id].pos.x[k] > rmax[k]) rmax[k] = n[id].pos.x[k]; }
  }
  void range_rec(int id, const V3& q, std::vector<int>& visit_order) {
    if (id < 0) return;
    kd_visits++;
    const TreeNode& t = n[id];
    double d2 = 0;
    for (int k = 0; k < 3; k++) d2 += (t.pos.x[k] - q.x[k]) * (t.pos.x[k] - q.x[k]);
    if (d2 <= radius * radius) visit_order.push_back(id);
    const double dx = q.x[t.dir] - t.pos.x[t.dir];
    range_rec(dx <= 0.0 ? t.left : t.right, q, visit_order);
    if (std::fabs(dx) < radius) range_rec(dx <= 0.0 ? t.right : t.left, q, visit_order);
  }
  std::vector<int> range(const V3& q) {  // in the order kd_res_next walks them
    std::vector<int> v; range_rec(root, q, v);
    return std::vector<int>(v.rbegin(), v.rend());
  }
  static double rect_d2(const double* lo, const double* hi, const V3& q) {
    double r = 0;
    for (int k = 0; k < 3; k++) {
      if (q.x[k] < lo[k]) r += (lo[k] - q.x[k]) * (lo[k] - q.x[k]);
      else if (q.x[k] > hi[k]) r += (hi[k] - q.x[k]) * (hi[k] - q.x[k]);
    }
    return r;
  }
  void nearest_rec(int id, const V3& q, int& best, double& best_d2, double* lo, double* hi) {
    kd_visits++;
    const TreeNode& t = n[id];
    const int dir = t.dir;
    const bool neg = (q.x[dir] - t.pos.x[dir]) <= 0;
    const int nearer = neg ? t.left : t.right, farther = neg ? t.right : t.left;
    double* ncoord = neg ? hi + dir : lo + dir;
    double* fcoord = neg ? lo + dir : hi + dir;
    if (nearer >= 0) { const double keep = *ncoord; *ncoord = t.pos.x[dir]; nearest_rec(nearer, q, best, best_d2, lo, hi); *ncoord = keep; }
    double d2 = 0;
    for (int k = 0; k < 3; k++) d2 += (t.pos.x[k] - q.x[k]) * (t.pos.x[k] - q.x[k]);
    if (d2 < best_d2) { best = id; best_d2 = d2; }
    if (farther >= 0) {
      const double keep = *fcoord; *fcoord = t.pos.x[dir];
      if (rect_d2(lo, hi, q) < best_d2) nearest_rec(farther, q, best, best_d2, lo, hi);
      *fcoord = keep;
    }
  }
  int nearest(const V3& q) {
    double lo[3], hi[3];
    for (int k = 0; k < 3; k++) { lo[k] = rmin[k]; hi[k] = rmax[k]; }
    int best = root; double d2 = 0;
    for (int k = 0; k < 3; k++) d2 += (n[root].pos.x[k] - q.x[k]) * (n[root].pos.x[k] - q.x[k]);
    nearest_rec(root, q, best, d2, lo, hi);
    return best;
  }
  bool collision_free(const V3& from, const V3& to) {
    const V3 d = normalized(sub(to, from));
    const double len = norm(d);
    for (double t = 0; t < len; t += ccres) {
      const V3 p{{from.x[0] + t * d.x[0], from.x[1] + t * d.x[1], from.x[2] + t * d.x[2]}};
      if (m.lookup(p)) return false;
    }
    return true;
  }
  int choose_parent(const V3& x_new) {
    double compare = (double)(1 << 30); int parent = -1;
    for (int nb : range(x_new)) {
      const double g_new = n[nb].g + norm(sub(n[nb].pos, x_new));
      if (g_new < compare && collision_free(n[nb].pos, x_new)) { compare = g_new; parent = nb; }
    }
    if (compare == (double)(1 << 30)) return -1;
    const int id = use++;
    n[id].pos = x_new; n[id].parent = parent; n[id].g = compare;
    n[parent].children.push_back(id);
    return id;
  }
  void rewire(int nw) {
    for (int nb : range(n[nw].pos)) {
      const double g_new = n[nw].g + norm(sub(n[nb].pos, n[nw].pos));
      if (g_new < n[nb].g && collision_free(n[nw].pos, n[nb].pos)) {
        std::vector<int>& oc = n[n[nb].parent].children;
        std::vector<int> kept;
        for (int c : oc) if (c != nb) kept.push_back(c);
        oc.swap(kept);
        n[nb].parent = nw; n[nb].g = g_new;
        n[nw].children.push_back(nb);
        std::deque<int> q{nb};
        while (!q.empty()) {
          const int cur = q.front(); q.pop_front();
          for (int c : n[cur].children) { n[c].g = n[cur].g + norm(sub(n[c].pos, n[cur].pos)); q.push_back(c); }
        }
      }
    }
  }
};

inline unsigned int seed32(unsigned long long query_seed, long long i) {
  unsigned long long z = query_seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned int)(z >> 32);
}
inline void fold(unsigned long long& h, unsigned long long v) { h ^= v; h *= 0x100000001b3ull; h ^= h >> 29; }
inline unsigned long long bits(double d) { unsigned long long b; std::memcpy(&b, &d, 8); return b; }

}  // namespace

extern "C" int oracle_rrt_search(int max_tree_node_num, double step_length, double search_radius, double collision_check_resolution,
                                 double sample_budget, unsigned long long query_seed, const int8_t* occ_inflate, int nx, int ny, int nz,
                                 const double origin[3], const double map_size[3], double resolution, const double sp[3],
                                 const double ep[3], oracle_rrt_result* res, double* opt_path_xyz, int path_cap) {
  Rrt r;
  r.m.occ = occ_inflate; r.m.nx = nx; r.m.ny = ny; r.m.nz = nz; r.m.res = resolution; r.m.inv = 1.0 / resolution;
  for (int k = 0; k < 3; k++) { r.m.o[k] = origin[k]; r.m.mx[k] = origin[k] + map_size[k]; }
  r.step = step_length; r.radius = search_radius; r.ccres = collision_check_resolution;
  r.n.resize((size_t)max_tree_node_num + 2);
  for (auto& t : r.n) { t.g = (double)(1 << 30); t.parent = -1; t.left = t.right = -1; t.dir = 0; }
  const V3 start{{sp[0], sp[1], sp[2]}}, end{{ep[0], ep[1], ep[2]}};
  r.n[0].pos = start; r.n[0].g = 0.0; r.n[0].parent = -1;
  r.n[1].pos = end;
  r.use = 2;
  r.kd_insert(0);
  const int goal = 1;
  bool reach = false;
  double feasible = (double)(1 << 30);
  std::vector<V3> optimal;
  long long samples = 0;
  int status = 0;
  for (int i = 0; i < max_tree_node_num && !status; i++) {
    std::mt19937_64 gen(seed32(query_seed, samples++));             // rrt_star.cpp:107-108: mt19937_64 gen(rd())
    std::uniform_real_distribution<> dis(0.0, 1.0);
    V3 x_rand;
    for (int k = 0; k < 3; k++) x_rand.x[k] = dis(gen) * map_size[k] + origin[k];
    const int nn = r.nearest(x_rand);
    const V3 dir = normalized(sub(x_rand, r.n[nn].pos));
    const V3 x_new{{r.n[nn].pos.x[0] + dir.x[0] * step_length, r.n[nn].pos.x[1] + dir.x[1] * step_length, r.n[nn].pos.x[2] + dir.x[2] * step_length}};
    if (r.m.lookup(x_new) == 1) continue;
    const int nw = r.choose_parent(x_new);
    if (nw < 0) continue;
    r.kd_insert(nw);
    r.rewire(nw);
    if (norm(sub(r.n[nw].pos, end)) <= search_radius) {
      if (!r.collision_free(r.n[nw].pos, end)) continue;
      const double via = r.n[nw].g + norm(sub(end, r.n[nw].pos));
      if (!reach) {
        reach = true;
        r.n[nw].children.push_back(goal); r.n[goal].parent = nw; r.n[goal].g = via; feasible = via;
      } else if (via < feasible) {
        r.n[goal].parent = nw; r.n[nw].children.push_back(goal); r.n[goal].g = via;
      }
    }
    if (reach) {
      if (r.n[goal].g < feasible) {
        feasible = r.n[goal].g;
        optimal.clear();
        for (int t = goal; t >= 0; t = r.n[t].parent) optimal.push_back(r.n[t].pos);
        optimal = std::vector<V3>(optimal.rbegin(), optimal.rend());
      }
      if ((double)samples - 0.0 >= sample_budget) status = 1;
    }
  }
  if (!status) status = reach ? 1 : 2;
  res->status = status; res->use_node_num = r.use; res->n_opt_path = (int)optimal.size(); res->reach_goal = reach ? 1 : 0;
  res->n_samples = samples; res->goal_g_cost = r.n[goal].g; res->n_kd_visits = r.kd_visits;
  unsigned long long h = 0;  // order-free sum of per-node digests (the CUDA kernel computes it lane-parallel)
  for (int i = 0; i < r.use; i++) {
    unsigned long long hn = 0xcbf29ce484222325ull ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1));
    for (int k = 0; k < 3; k++) fold(hn, bits(r.n[i].pos.x[k]));
    fold(hn, bits(r.n[i].g));
    fold(hn, r.n[i].parent >= 0 ? (unsigned long long)r.n[i].parent : 0xffffffffull);
    h += hn;
  }
  res->tree_digest = h;
  for (int i = 0; i < (int)optimal.size() && i < path_cap; i++) for (int k = 0; k < 3; k++) opt_path_xyz[3 * i + k] = optimal[i].x[k];
  return status;
}
