// oracle/kino_ref.cpp — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).
//
// CPU restatement of path_searching::KinoAstar (reference:
//   src/planner/path_searching/src/kino_astar.cpp, include/path_searching/kino_astar.h) and of the four
//   GridMap lookups it calls (src/planner/plan_env/include/plan_env/grid_map.h:257-260,350-359,370-385,400-404).
// It follows the reference statement by statement, bug for bug, with
// std::priority_queue / std::unordered_map exactly as the reference uses them, but without Eigen / ROS / PCL:
//   * Eigen expressions are written out in scalar f64 with the association given in SURVEY.md §9.1
//     ("floating-point contract"); Eigen itself is absent from this container, so that association is
//     "believed", not verified;
//   * pcl::KdTreeFLANN::radiusSearch is replaced by a uniform cell list followed by the same float32
//     squared-distance filter; because the filter radius (r + 0.1) exceeds the largest ellipsoid semi-axis the
//     accept/reject result of isCollisionFree does not depend on which superset is returned;
//   * cbrt/acos/cos/pow come from csrc/fpmath.h (shared with the device) unless libm_mode == 1, in which case
//     glibc is called like the reference does; tests compare both modes.
// PARITY STATUS: the reference has no expected outputs for this path.  This restatement is pinned to the
// reference ITSELF: /root/reference's kino_astar.cpp compiles unmodified against the header shims in oracle/shim into
// oracle/_ref/libkino_ref.so (oracle/Makefile), and tests/test_kino_reference_build.py requires bit-identical return
// codes, use_node_num_, path points and map-lookup sequences from the two on the golden queries and on > 150 random ones.
// What remains "believed" is Eigen's floating-point association (restated in oracle/shim/Eigen/Eigen, same contract).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <queue>
#include <unordered_map>
#include <vector>

#include "../uav_motion_planning_b200/csrc/fpmath.h"
#include "oracle.h"

namespace {

struct V3 {
  double x, y, z;
  double operator()(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
// Eigen 3.3 fixed-size-3 reduction: packet of two, then the remainder
inline double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { return a / norm(a); }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double linf(V3 a) { return std::max(std::max(std::fabs(a.x), std::fabs(a.y)), std::fabs(a.z)); }

struct I3 {
  int x, y, z;
  bool operator==(const I3& o) const { return x == o.x && y == o.y && z == o.z; }
};

// kino_astar.h:63-77 (vector3i_hash; std::hash<int> is the identity in libstdc++)
struct I3Hash {
  size_t operator()(const I3& v) const {
    size_t seed = 0;
    const int e[3] = {v.x, v.y, v.z};
    for (int i = 0; i < 3; ++i) seed ^= std::hash<int>()(e[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
  }
};

const double kInf = (double)(1 << 30);  // kino_astar.h:22

struct Node {  // kino_astar.h:27-51
  V3 position, velocity, input;
  I3 index;
  double g_cost, f_cost, duration;
  Node* parent;
  char node_state;
  int lattice_id;  // oracle-only bookkeeping (which primitive produced `input`)
  Node() { reset(); }
  void reset() {
    g_cost = kInf; f_cost = kInf; parent = nullptr; duration = kInf; input = {0, 0, 0}; node_state = 'c';
    lattice_id = -1;
  }
};
struct NodeCmp {  // kino_astar.h:54-61
  bool operator()(Node* a, Node* b) const { return a->f_cost > b->f_cost; }
};

struct Mat3 { double m[3][3]; };

}  // namespace

struct oracle_kino {
  oracle_kino_params p;
  // GridMap state (grid_map.cpp:52-54,69-73)
  int nx, ny, nz;
  V3 origin, map_size, bmin, bmax;
  double resolution, inv_resolution;
  std::vector<int8_t> occ;  // occupancy_buffer_inflate_, x-major / z-fastest
  // cloud (kino_astar.cpp:38-55): obs_ as f64 copies of the float32 message
  std::vector<V3> obs;
  std::vector<float> obs_f;
  // stand-in for the KD-tree: uniform cell list
  double cell = 0.5;
  int cnx = 0, cny = 0, cnz = 0;
  V3 cmin{0, 0, 0};
  std::vector<int> cell_start, cell_pts;

  // search state (kino_astar.h:124-128)
  std::vector<Node*> pool;
  int use_node_num = 0;
  double tie_breaker = 1.0 + (3 / 1e4);
  double shot_coef[3][4], vel_coef[3][4], acc_coef[3][4];
  oracle_kino_counters cnt;

  // ---- libm selection ------------------------------------------------------------------
  double m_cbrt(double x) const { return p.libm_mode ? std::cbrt(x) : fpm::cbrt(x); }
  double m_acos(double x) const { return p.libm_mode ? std::acos(x) : fpm::acos(x); }
  double m_cos(double x) const { return p.libm_mode ? std::cos(x) : fpm::cos(x); }
  double m_powi(double x, int n) const { return p.libm_mode ? std::pow(x, n) : fpm::powi(x, n); }

  // ---- GridMap lookups -----------------------------------------------------------------
  bool isInMap(V3 pos) const {  // grid_map.h:370-385
    if (pos.x < bmin.x + 1e-4 || pos.y < bmin.y + 1e-4 || pos.z < bmin.z + 1e-4) return false;
    if (pos.x > bmax.x - 1e-4 || pos.y > bmax.y - 1e-4 || pos.z > bmax.z - 1e-4) return false;
    return true;
  }
  I3 posToIndex(V3 pos) const {  // grid_map.h:400-404 == kino_astar.cpp:302-310
    I3 id;
    id.x = (int)std::floor((pos.x - origin.x) * inv_resolution);
    id.y = (int)std::floor((pos.y - origin.y) * inv_resolution);
    id.z = (int)std::floor((pos.z - origin.z) * inv_resolution);
    return id;
  }
  int getInflateOccupancy(V3 pos) {  // grid_map.h:350-359
    if (!isInMap(pos)) return -1;
    I3 id = posToIndex(pos);
    cnt.n_occ_lookup++;
    return (int)occ[(size_t)id.x * ny * nz + (size_t)id.y * nz + id.z];
  }

  void buildCells() {
    if (obs.empty()) return;
    V3 lo = obs[0], hi = obs[0];
    for (const V3& q : obs) {
      lo = {std::min(lo.x, q.x), std::min(lo.y, q.y), std::min(lo.z, q.z)};
      hi = {std::max(hi.x, q.x), std::max(hi.y, q.y), std::max(hi.z, q.z)};
    }
    cmin = lo;
    cnx = (int)std::floor((hi.x - lo.x) / cell) + 1;
    cny = (int)std::floor((hi.y - lo.y) / cell) + 1;
    cnz = (int)std::floor((hi.z - lo.z) / cell) + 1;
    std::vector<int> count((size_t)cnx * cny * cnz + 1, 0);
    auto cid = [&](const V3& q) {
      int ix = (int)std::floor((q.x - lo.x) / cell), iy = (int)std::floor((q.y - lo.y) / cell),
          iz = (int)std::floor((q.z - lo.z) / cell);
      return ((size_t)ix * cny + iy) * cnz + iz;
    };
    for (const V3& q : obs) count[cid(q) + 1]++;
    for (size_t i = 1; i < count.size(); i++) count[i] += count[i - 1];
    cell_start = count;
    cell_pts.resize(obs.size());
    std::vector<int> fill(count.begin(), count.end() - 1);
    for (size_t i = 0; i < obs.size(); i++) cell_pts[fill[cid(obs[i])]++] = (int)i;
  }

  // kino_astar.cpp:721-758
  bool isCollisionFree(V3 pt, V3 acc) {
    V3 b3 = normalized(acc + 9.81 * V3{0, 0, 1});
    V3 c1{std::cos(0.0), std::sin(0.0), 0};
    V3 b2 = normalized(cross(b3, c1));
    V3 b1 = normalized(cross(b2, b3));
    double Rot[3][3] = {{b1.x, b2.x, b3.x}, {b1.y, b2.y, b3.y}, {b1.z, b2.z, b3.z}};
    double Pd[3] = {p.robot_r, p.robot_r, p.robot_h};
    // E = Rot * P * Rot^T, (Rot*P) first, each 3-term sum packet-of-two then remainder
    double RP[3][3], E[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) RP[i][j] = Rot[i][j] * Pd[j];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) E[i][j] = (RP[i][0] * Rot[j][0] + RP[i][1] * Rot[j][1]) + RP[i][2] * Rot[j][2];
    if (obs.empty()) return true;
    float sx = (float)pt.x, sy = (float)pt.y, sz = (float)pt.z;
    float radius = (float)(p.robot_r + 1e-1);
    float r2 = radius * radius;
    // candidate cells
    int x0 = (int)std::floor((pt.x - radius - cmin.x) / cell), x1 = (int)std::floor((pt.x + radius - cmin.x) / cell);
    int y0 = (int)std::floor((pt.y - radius - cmin.y) / cell), y1 = (int)std::floor((pt.y + radius - cmin.y) / cell);
    int z0 = (int)std::floor((pt.z - radius - cmin.z) / cell), z1 = (int)std::floor((pt.z + radius - cmin.z) / cell);
    x0 = std::max(x0, 0); y0 = std::max(y0, 0); z0 = std::max(z0, 0);
    x1 = std::min(x1, cnx - 1); y1 = std::min(y1, cny - 1); z1 = std::min(z1, cnz - 1);
    bool have_inv = false;
    double Ei[3][3];
    for (int ix = x0; ix <= x1; ix++)
      for (int iy = y0; iy <= y1; iy++)
        for (int iz = z0; iz <= z1; iz++) {
          size_t c = ((size_t)ix * cny + iy) * cnz + iz;
          for (int k = cell_start[c]; k < cell_start[c + 1]; k++) {
            int id = cell_pts[k];
            float dx = obs_f[3 * id] - sx, dy = obs_f[3 * id + 1] - sy, dz = obs_f[3 * id + 2] - sz;
            float d2 = dx * dx + dy * dy + dz * dz;
            if (!(d2 <= r2)) continue;  // the KD-tree radius filter (float32)
            cnt.n_cloud_pts_tested++;
            if (!have_inv) {  // E.inverse() (kino_astar.cpp:752): Eigen's 3x3 path = cofactors of column 0,
                              // det = cofactors_col0 . E.col(0), then every cofactor times 1/det
              auto cof = [&](int i, int j) {
                int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                return E[i1][j1] * E[i2][j2] - E[i1][j2] * E[i2][j1];
              };
              double cc0 = cof(0, 0), cc1 = cof(1, 0), cc2 = cof(2, 0);
              double det = (cc0 * E[0][0] + cc1 * E[1][0]) + cc2 * E[2][0];
              double invdet = 1.0 / det;
              Ei[0][0] = cc0 * invdet; Ei[0][1] = cc1 * invdet; Ei[0][2] = cc2 * invdet;
              Ei[1][0] = cof(0, 1) * invdet; Ei[1][1] = cof(1, 1) * invdet; Ei[1][2] = cof(2, 1) * invdet;
              Ei[2][0] = cof(0, 2) * invdet; Ei[2][1] = cof(1, 2) * invdet; Ei[2][2] = cof(2, 2) * invdet;
              have_inv = true;
            }
            V3 d = obs[id] - pt;
            V3 t{(Ei[0][0] * d.x + Ei[0][1] * d.y) + Ei[0][2] * d.z, (Ei[1][0] * d.x + Ei[1][1] * d.y) + Ei[1][2] * d.z,
                 (Ei[2][0] * d.x + Ei[2][1] * d.y) + Ei[2][2] * d.z};
            if (norm(t) <= 1.0) return false;
          }
        }
    return true;
  }

  // kino_astar.cpp:651-670 — the zero entries of e_At / Integral contribute exact 0.0 terms
  void StateTransit(const double x0[6], double xt[6], V3 ut, double t) const {
    double h = 0.5 * t * t;
    xt[0] = (x0[0] + t * x0[3]) + h * ut.x;
    xt[1] = (x0[1] + t * x0[4]) + h * ut.y;
    xt[2] = (x0[2] + t * x0[5]) + h * ut.z;
    xt[3] = x0[3] + t * ut.x;
    xt[4] = x0[4] + t * ut.y;
    xt[5] = x0[5] + t * ut.z;
  }

  // kino_astar.cpp:339-372
  int cubic(double a, double b, double c, double d, double* dts) const {
    int n = 0;
    double a2 = b / a, a1 = c / a, a0 = d / a;
    double Q = (3 * a1 - a2 * a2) / 9;
    double R = (9 * a1 * a2 - 27 * a0 - 2 * a2 * a2 * a2) / 54;
    double D = Q * Q * Q + R * R;
    if (D > 0) {
      double S = m_cbrt(R + std::sqrt(D));
      double T = m_cbrt(R - std::sqrt(D));
      dts[n++] = -a2 / 3 + (S + T);
    } else if (D == 0) {
      double S = m_cbrt(R);
      dts[n++] = -a2 / 3 + S + S;
      dts[n++] = -a2 / 3 - S;
    } else {
      double theta = m_acos(R / std::sqrt(-Q * Q * Q));
      dts[n++] = 2 * std::sqrt(-Q) * m_cos(theta / 3) - a2 / 3;
      dts[n++] = 2 * std::sqrt(-Q) * m_cos((theta + 2 * M_PI) / 3) - a2 / 3;
      dts[n++] = 2 * std::sqrt(-Q) * m_cos((theta + 4 * M_PI) / 3) - a2 / 3;
    }
    return n;
  }

  // kino_astar.cpp:374-414
  int quartic(double a, double b, double c, double d, double e, double* dts) const {
    int n = 0;
    double a3 = b / a, a2 = c / a, a1 = d / a, a0 = e / a;
    double ys[3];
    cubic(1, -a2, a1 * a3 - 4 * a0, 4 * a2 * a0 - a1 * a1 - a3 * a3 * a0, ys);
    double y1 = ys[0];
    double r = a3 * a3 / 4 - a2 + y1;
    if (r < 0) return 0;
    double R = std::sqrt(r);
    double D, E;
    if (R != 0) {
      D = std::sqrt(0.75 * a3 * a3 - R * R - 2 * a2 + 0.25 * (4 * a3 * a2 - 8 * a1 - a3 * a3 * a3) / R);
      E = std::sqrt(0.75 * a3 * a3 - R * R - 2 * a2 - 0.25 * (4 * a3 * a2 - 8 * a1 - a3 * a3 * a3) / R);
    } else {
      D = std::sqrt(0.75 * a3 * a3 - 2 * a2 + 2 * std::sqrt(y1 * y1 - 4 * a0));
      E = std::sqrt(0.75 * a3 * a3 - 2 * a2 - 2 * std::sqrt(y1 * y1 - 4 * a0));
    }
    if (!std::isnan(D)) {
      dts[n++] = -a3 / 4 + R / 2 + D / 2;
      dts[n++] = -a3 / 4 + R / 2 - D / 2;
    }
    if (!std::isnan(E)) {
      dts[n++] = -a3 / 4 - R / 2 + E / 2;
      dts[n++] = -a3 / 4 - R / 2 - E / 2;
    }
    return n;
  }

  // kino_astar.cpp:312-337
  double getHeuristicCost(V3 x1, V3 v1, V3 x2, V3 v2, double& optimal_time) {
    cnt.n_heuristic++;
    V3 dp = x2 - x1;
    double optimal_cost = kInf;
    double a = -36 * dot(dp, dp);
    double b = 24 * dot(dp, v1 + v2);
    double c = -4 * (dot(v1, v1) + dot(v1, v2) + dot(v2, v2));
    double d = 0;
    double e = p.rou_time;
    double dts[4];
    int n = quartic(e, d, c, b, a, dts);
    double T_bar = linf(x1 - x2) / p.max_velocity;
    for (int i = 0; i < n; i++) {
      double t = dts[i];
      double tmp_cost = a / (-3 * t * t * t) + b / (-2 * t * t) + c / (-1 * t) + e * t;
      if (tmp_cost < optimal_cost && t > T_bar && tmp_cost > 0) {
        optimal_cost = tmp_cost;
        optimal_time = t;
      }
    }
    return tie_breaker * optimal_cost;
  }

  // kino_astar.cpp:416-471
  bool computeShotTraj(V3 x1, V3 v1, V3 x2, V3 v2, double optimal_time) {
    double td = optimal_time;
    V3 dp = x2 - x1, dv = v2 - v1;
    V3 c2 = 0.5 * ((6 / (td * td)) * (dp - v1 * td) - (2 * dv) / td);
    V3 c3 = (1.0 / 6.0) * ((-12 / (td * td * td)) * (dp - v1 * td) + (6 * dv) / (td * td));
    for (int i = 0; i < 3; i++) {
      shot_coef[i][0] = x1(i); shot_coef[i][1] = v1(i); shot_coef[i][2] = c2(i); shot_coef[i][3] = c3(i);
      // vel_coef = shot_coef * Transit_v, acc_coef = vel_coef * Transit_a (zeros are exact no-ops)
      vel_coef[i][0] = shot_coef[i][1]; vel_coef[i][1] = shot_coef[i][2] * 2; vel_coef[i][2] = shot_coef[i][3] * 3;
      vel_coef[i][3] = 0;
      acc_coef[i][0] = vel_coef[i][1]; acc_coef[i][1] = vel_coef[i][2] * 2; acc_coef[i][2] = 0; acc_coef[i][3] = 0;
    }
    int segment_num = (int)std::floor(td / p.time_step_size);
    for (int j = 0; j <= segment_num; j++) {
      double curr_t = j * p.time_step_size;
      V3 shot_pos = shotPos(curr_t);
      if (getInflateOccupancy(shot_pos)) return false;
    }
    return true;
  }
  V3 shotPos(double t) const {
    double tv[4];
    for (int i = 0; i < 4; i++) tv[i] = m_powi(t, i);
    double r[3];
    for (int i = 0; i < 3; i++)
      r[i] = ((shot_coef[i][0] * tv[0] + shot_coef[i][1] * tv[1]) + shot_coef[i][2] * tv[2]) + shot_coef[i][3] * tv[3];
    return {r[0], r[1], r[2]};
  }

  // kino_astar.cpp:492-557 (samplePath) and :559-649 (sampleEllipsoid: same positions, plus rot_list)
  void samplePath(const std::vector<Node*>& path_pool, std::vector<V3>& path) {
    if (path_pool.size() != 1) {
      for (size_t i = 0; i + 1 < path_pool.size(); i++) {
        Node* curr = path_pool[i];
        Node* next = path_pool[i + 1];
        double x0[6] = {curr->position.x, curr->position.y, curr->position.z,
                        curr->velocity.x, curr->velocity.y, curr->velocity.z};
        double xt[6];
        int segment_num = (int)std::floor(curr->duration / p.time_step_size);
        for (int j = 0; j < segment_num; j++) {
          double curr_t = j * p.time_step_size;
          StateTransit(x0, xt, next->input, curr_t);
          path.push_back({xt[0], xt[1], xt[2]});
        }
      }
    }
    Node* last = path_pool.back();
    double td = last->duration;
    int segment_num = (int)std::floor(td / p.time_step_size);
    for (int j = 0; j <= segment_num; j++) path.push_back(shotPos(j * p.time_step_size));
  }

  static uint64_t mix(uint64_t h, uint64_t v) {  // pop-sequence digest shared with the device (FNV-1a style)
    h ^= v;
    h *= 0x100000001b3ull;
    h ^= h >> 29;
    return h;
  }

  // kino_astar.cpp:81-272
  int search(V3 start_pt, V3 start_vel, V3 end_pt, V3 end_vel, std::vector<V3>& path, oracle_kino_result* res,
             int32_t* pop_trace, int pop_cap) {
    std::priority_queue<Node*, std::vector<Node*>, NodeCmp> open_list;
    std::unordered_map<I3, Node*, I3Hash> close_list, expanded_list;
    memset(&cnt, 0, sizeof(cnt));
    uint64_t pop_hash = 0xcbf29ce484222325ull;
    // digest of every position the search itself passes to isInMap (kino_astar.cpp:176), in call order: the same fingerprint
    // oracle/shim/plan_env/grid_map.h takes of the UNMODIFIED reference search compiled in oracle/_ref/libkino_ref.so
    uint64_t lookup_digest = 0xcbf29ce484222325ull, n_in_map_calls = 0;
    auto fold_lookup = [&](const V3& q) {
      const double c[3] = {q.x, q.y, q.z};
      for (int i = 0; i < 3; i++) { lookup_digest ^= fpm::to_bits(c[i]); lookup_digest *= 0x100000001b3ull; lookup_digest ^= lookup_digest >> 29; }
      n_in_map_calls++;
    };
    int n_pop = 0;
    auto finish = [&](int status) {
      res->status = status;
      res->use_node_num = use_node_num;
      res->n_pop = n_pop;
      res->pop_hash = pop_hash;
      res->counters = cnt;
      res->lookup_digest = lookup_digest;
      res->n_in_map_calls = (long long)n_in_map_calls;
      return status;
    };

    double inv_acc_res = 1.0 / p.acc_resolution;
    double optimal_time = kInf;
    Node* start_node = pool[use_node_num];
    start_node->position = start_pt;
    start_node->velocity = start_vel;
    start_node->index = posToIndex(start_pt);
    start_node->g_cost = 0.0;
    start_node->f_cost = p.lambda_heu * getHeuristicCost(start_pt, start_vel, end_pt, end_vel, optimal_time);
    use_node_num++;
    open_list.push(start_node);
    expanded_list.insert({start_node->index, start_node});
    cnt.n_insert++;
    start_node->node_state = 'a';

    while (!open_list.empty()) {
      Node* cur = open_list.top();
      open_list.pop();
      close_list.insert({cur->index, cur});
      cur->node_state = 'b';
      cur->duration = p.sample_tau;
      cnt.n_pop++;
      cnt.heap_len_sum += (long long)open_list.size();
      // digest of the expansion order: voxel index + exact state bits of the popped node
      pop_hash = mix(pop_hash, (uint64_t)(uint32_t)cur->index.x);
      pop_hash = mix(pop_hash, (uint64_t)(uint32_t)cur->index.y);
      pop_hash = mix(pop_hash, (uint64_t)(uint32_t)cur->index.z);
      pop_hash = mix(pop_hash, fpm::to_bits(cur->position.x));
      pop_hash = mix(pop_hash, fpm::to_bits(cur->velocity.x));
      pop_hash = mix(pop_hash, fpm::to_bits(cur->g_cost));
      if (pop_trace && n_pop < pop_cap) {
        pop_trace[3 * n_pop] = cur->index.x; pop_trace[3 * n_pop + 1] = cur->index.y; pop_trace[3 * n_pop + 2] = cur->index.z;
      }
      n_pop++;

      if (norm(cur->position - end_pt) < p.goal_tolerance) {
        // :114 is a comma expression, not a call: tmp_cost = lambda_heu * optimal_time (only printed)
        bool shot = computeShotTraj(cur->position, cur->velocity, end_pt, end_vel, optimal_time);
        cnt.n_shot++;
        if (shot) {
          cur->duration = optimal_time;
          std::vector<Node*> path_pool;  // retrievePath :473-490
          Node* c = cur;
          while (c->parent != nullptr) { path_pool.push_back(c); c = c->parent; }
          path_pool.push_back(c);
          std::reverse(path_pool.begin(), path_pool.end());
          samplePath(path_pool, path);  // same positions for collision_check_type 1 and 2
          res->n_path_nodes = (int)path_pool.size();
          res->shot_duration = optimal_time;
          return finish(1);
        } else if (cur->parent != nullptr) {
        } else {
          return finish(2);
        }
      }

      for (double ax = -p.max_acceleration; ax <= p.max_acceleration + 1e-3; ax += inv_acc_res * p.max_acceleration)
        for (double ay = -p.max_acceleration; ay <= p.max_acceleration + 1e-3; ay += inv_acc_res * p.max_acceleration)
          for (double az = -p.max_acceleration; az <= p.max_acceleration + 1e-3; az += inv_acc_res * p.max_acceleration) {
            V3 ut{ax, ay, az};
            double x0[6] = {cur->position.x, cur->position.y, cur->position.z,
                            cur->velocity.x, cur->velocity.y, cur->velocity.z};
            double xt[6];
            int segment_num = (int)std::floor(p.sample_tau / p.time_step_size);
            bool flag = false, collision_flag = false;
            for (int i = 0; i <= segment_num; i++) {
              double t = i * p.time_step_size;
              StateTransit(x0, xt, ut, t);
              V3 tmp_pos{xt[0], xt[1], xt[2]};
              fold_lookup(tmp_pos);
              if (isInMap(tmp_pos) == false) { flag = true; break; }
              switch (p.collision_check_type) {
                case 1:
                  if (getInflateOccupancy(tmp_pos) == 1) { collision_flag = true; break; }
                  // no break: falls through into case 2
                  [[fallthrough]];
                case 2:
                  if (isCollisionFree(tmp_pos, ut) == false) { collision_flag = true; break; }
              }
              if (collision_flag) { flag = true; break; }
              if (xt[3] < -p.max_velocity || xt[3] > p.max_velocity || xt[4] < -p.max_velocity ||
                  xt[4] > p.max_velocity || xt[5] < -p.max_velocity || xt[5] > p.max_velocity) {
                flag = true;
                break;
              }
            }
            if (flag) continue;
            StateTransit(x0, xt, ut, p.sample_tau);
            V3 xpos{xt[0], xt[1], xt[2]}, xvel{xt[3], xt[4], xt[5]};
            I3 idx = posToIndex(xpos);
            cnt.n_hash_probe++;
            if (close_list.find(idx) != close_list.end()) continue;
            auto it = expanded_list.find(idx);
            if (it == expanded_list.end()) {
              Node* pro = pool[use_node_num];
              pro->position = xpos;
              pro->velocity = xvel;
              pro->index = idx;
              pro->g_cost = cur->g_cost + (dot(ut, ut) + p.rou_time) * p.sample_tau;
              pro->f_cost = pro->g_cost +
                            p.lambda_heu * getHeuristicCost(pro->position, pro->velocity, end_pt, end_vel, optimal_time);
              pro->parent = cur;
              pro->input = ut;
              pro->duration = p.sample_tau;
              pro->node_state = 'a';
              use_node_num++;
              open_list.push(pro);
              expanded_list.insert({pro->index, pro});
              cnt.n_insert++;
              if (use_node_num >= p.allocated_node_num) return finish(2);
            } else {
              double tmp_g = cur->g_cost + (dot(ut, ut) + p.rou_time) * p.sample_tau;
              Node* old = it->second;
              if (tmp_g < old->g_cost) {
                old->position = xpos;
                old->velocity = xvel;
                old->g_cost = tmp_g;
                old->f_cost = old->g_cost + p.lambda_heu * getHeuristicCost(old->position, old->velocity, end_pt,
                                                                           end_vel, optimal_time);
                old->parent = cur;
                old->input = ut;
                cnt.n_update++;
              }
            }
          }
    }
    return finish(2);
  }

  void reset() {  // kino_astar.cpp:274-300 (semantics per SURVEY.md §9.1 Q6)
    for (int i = 0; i < use_node_num; i++) pool[i]->reset();
    use_node_num = 0;
    memset(shot_coef, 0, sizeof(shot_coef));
    memset(vel_coef, 0, sizeof(vel_coef));
    memset(acc_coef, 0, sizeof(acc_coef));
  }
};

extern "C" {

oracle_kino* oracle_kino_create(const oracle_kino_params* p, const int8_t* occ, int nx, int ny, int nz,
                                const double origin[3], const double map_size[3], double resolution,
                                const float* cloud_xyz, int n_cloud) {
  oracle_kino* k = new oracle_kino();
  k->p = *p;
  k->nx = nx; k->ny = ny; k->nz = nz;
  k->origin = {origin[0], origin[1], origin[2]};
  k->map_size = {map_size[0], map_size[1], map_size[2]};
  k->bmin = k->origin;
  k->bmax = k->origin + k->map_size;  // grid_map.cpp:72-73
  k->resolution = resolution;
  k->inv_resolution = 1.0 / resolution;  // kino_astar.cpp:68
  k->occ.assign(occ, occ + (size_t)nx * ny * nz);
  k->obs.resize(n_cloud);
  k->obs_f.assign(cloud_xyz, cloud_xyz + (size_t)3 * n_cloud);
  for (int i = 0; i < n_cloud; i++)
    k->obs[i] = {(double)cloud_xyz[3 * i], (double)cloud_xyz[3 * i + 1], (double)cloud_xyz[3 * i + 2]};
  k->buildCells();
  k->pool.resize(p->allocated_node_num);  // kino_astar.cpp:57-63
  for (int i = 0; i < p->allocated_node_num; i++) k->pool[i] = new Node();
  return k;
}

void oracle_kino_destroy(oracle_kino* k) {
  if (!k) return;
  for (Node* n : k->pool) delete n;
  delete k;
}

int oracle_kino_search(oracle_kino* k, const double start_pt[3], const double start_vel[3], const double end_pt[3],
                       const double end_vel[3], oracle_kino_result* res, double* path_xyz, int path_cap,
                       int32_t* pop_trace, int pop_cap) {
  std::vector<V3> path;
  memset(res, 0, sizeof(*res));
  int st = k->search({start_pt[0], start_pt[1], start_pt[2]}, {start_vel[0], start_vel[1], start_vel[2]},
                     {end_pt[0], end_pt[1], end_pt[2]}, {end_vel[0], end_vel[1], end_vel[2]}, path, res, pop_trace,
                     pop_cap);
  res->n_path = (int)path.size();
  if (path_xyz) {
    int m = std::min((int)path.size(), path_cap);
    for (int i = 0; i < m; i++) { path_xyz[3 * i] = path[i].x; path_xyz[3 * i + 1] = path[i].y; path_xyz[3 * i + 2] = path[i].z; }
  }
  k->reset();  // the caller pattern of test_kino_astar_searching.cpp:69-70
  return st;
}

// scalar entry points for tests/test_fpmath.py
double oracle_fp_eval(int op, double x, int n) {
  switch (op) {
    case 0: return fpm::cbrt(x);
    case 1: return fpm::acos(x);
    case 2: return fpm::cos(x);
    case 3: return fpm::powi(x, n);
    case 10: return std::cbrt(x);
    case 11: return std::acos(x);
    case 12: return std::cos(x);
    case 13: return std::pow(x, n);
  }
  return 0.0;
}
void oracle_fp_eval_n(int op, const double* x, int n_pow, double* y, long long n) {
  for (long long i = 0; i < n; i++) y[i] = oracle_fp_eval(op, x[i], n_pow);
}

}  // extern "C"
