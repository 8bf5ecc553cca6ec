// kino_kernel.cu — K1: batched kinodynamic A* (one CTA per query) for sm_100a.
//
// Replaces path_searching::KinoAstar::search and its callees
// (reference: src/planner/path_searching/src/kino_astar.cpp:81-272 search, :651-670 StateTransit,
//  :312-337 getHeuristicCost, :339-414 cubic/quartic, :416-471 computeShotTraj, :721-758 isCollisionFree,
//  :473-557 retrievePath/samplePath; plan_env/grid_map.h:350-404 the grid lookups).
//
// Execution model.  The search is sequential per query (pop -> expand -> commit) and the open list is a binary
// heap whose keys the reference mutates in place, so the pop ORDER is only defined by
// libstdc++'s push_heap/pop_heap.  Each CTA owns one query and an arena in HBM (node records, heap array,
// open-addressing hash table).  Per expansion:
//   A. all 256 threads evaluate the (2r+1)^3 motion primitives in parallel: checkpoint integration, in-map,
//      inflated-grid byte, SE(3) ellipsoid-vs-cloud test (only where a dilated "cloud nearby" bit is set),
//      velocity limits, end state -> voxel key;
//   B. keys are de-duplicated inside the expansion (shared-memory table), group leaders probe the global hash
//      table, the OBVP quartic heuristic is evaluated for every candidate in parallel;
//   C. new node ids are assigned by an ordered block scan (== the reference's use_node_num_++ order), node
//      records and hash slots are written in parallel;
//   D. warp 0 replays the heap pushes / in-place key mutations in lattice order.  The ancestors of all new
//      leaves are staged in shared memory first ("closure"), so one push costs a ballot, not a chain of
//      dependent HBM loads.
// All f64 arithmetic is written with the reference's association and compiled with -fmad=false; cbrt/acos/cos/pow
// come from fpmath.h, which the CPU oracle shares, so popped-node sequences are bit-identical.
#include <cub/cub.cuh>
#include <cuda.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>

#include "fpmath.h"
#include "qp_body_warp.h"
#include "qp_plan.h"
#include "uavmp_internal.h"

#ifndef KT
#define KT 256
#endif
#define TAB_SIZE 2048
#define PUSH_BATCH 256
#define HC_CAP 640
#define TB 32          // x / y edge of the flags box staged per expansion (voxels)
#define TBZ 48         // z edge: the TMA inner coordinate must be 16 B aligned, so the box starts at z & ~15
#define PTS_CAP 2048   // cloud points staged per expansion
#define CS_CAP 1280    // cell_start entries staged per expansion
#define COL_CAP 128    // cell columns staged per expansion
#define HTOP 256       // heap slots 1 .. HTOP-1 (the top 8 levels) live in shared memory: a pop's sift-down starts on chip
#define FULL 0xffffffffu
// phase timers: thread 0 charges the cycles since the last mark to phase `k`
#define PH_MARK(k) do { if (prof && tid == 0) { long long t_ = clock64(); s.ph[k] += (unsigned long long)(t_ - s.ph_t); s.ph_t = t_; } } while (0)
#define KEY_BIAS (1 << 17)
#define EPOCH_BITS 10
#define EPOCH_MASK ((1u << EPOCH_BITS) - 1u)

namespace {

enum : uint8_t {
  ST_REJECT = 0, ST_FEASIBLE = 1, ST_CLOSED = 2, ST_NEW = 3, ST_OPEN_CAND = 4, ST_OPEN_NOCAND = 5,
  ST_FOLLOW_CAND = 6, ST_FOLLOW_NOCAND = 7
};

struct PhaseA {  // staging buffers of the primitive evaluation
  uint8_t tile[TB * TB * TBZ];   // flags box, [x][y][z] z fastest (the TMA destination: keep first, 128 B aligned)
  float4 pts[PTS_CAP];          // cloud points of the cells the expansion can touch, column after column
  int cstart[CS_CAP];           // their cell_start entries, [column][z]
  int col_delta[COL_CAP];       // shared-memory index = global point index + col_delta[column]
  uint16_t need[UAVMP_MAXPRIM]; // per primitive: checkpoints whose voxel has the "cloud nearby" bit
};
struct PhaseB {  // successor classification / commit
  double f[UAVMP_MAXPRIM];
  double topt[UAVMP_MAXPRIM];
  double gcur[UAVMP_MAXPRIM];
  uint32_t hs[UAVMP_MAXPRIM];    // hash slot of the group's node
  uint32_t hpos[UAVMP_MAXPRIM];  // its open-list position at probe time
  uint32_t tab[TAB_SIZE];
  HeapSlot hc[HC_CAP];
  uint32_t hidx[HC_CAP];         // 1-based heap index of every staged ancestor slot
  int winm[UAVMP_MAXPRIM];       // position in list2 of the last improving candidate of a group (-1: none)
  uint8_t imp[UAVMP_MAXPRIM];    // per candidate: 0 not improving, 1 improving, 2 improving and ordered with the pushes
  uint8_t inun[UAVMP_MAXPRIM];   // that node is an ancestor of a new leaf: its key was already written in order
  double gpc[UAVMP_MAXPRIM];     // g of every commit candidate
  HeapSlot ev[UAVMP_MAXPRIM];    // per commit candidate (list2 order): what the ordered commit replays — a push {f, id, hash slot},
                                 // a key mutation of an ancestor of a new leaf {f, id, hash slot | EV_SETKEY}, or nothing (id NONE)
  HeapSlot evc[UAVMP_MAXPRIM];   // the same without the "nothing" records (compacted by warp 0 at the start of the replay)
};
#define EV_SETKEY 0x80000000u

struct SearchSmem {
  union __align__(128) { PhaseA a; PhaseB b; };
  union {
    struct { unsigned long long key[UAVMP_MAXPRIM]; uint32_t id[UAVMP_MAXPRIM]; };
    uint32_t units[UAVMP_MAXK * UAVMP_MAXNA * UAVMP_MAXNA];  // cloud-test work units; dead before the keys are written
  };
  uint16_t list1[UAVMP_MAXPRIM];
  uint16_t list2[UAVMP_MAXPRIM];
  uint8_t state[UAVMP_MAXPRIM];
  // separable tables of one expansion
  double X[UAVMP_MAXK][3][UAVMP_MAXNA];
  int XI[UAVMP_MAXK][3][UAVMP_MAXNA];
  double EX[3][UAVMP_MAXNA], EV[3][UAVMP_MAXNA];
  int EI[3][UAVMP_MAXNA];
  uint8_t axok[3][UAVMP_MAXNA];
  double xlo[3], xhi[3];
  int to[3], rc0[3], rc1[3];
  int any_ok, tile_ok, nT, npts, n_upd, last_ev, nq;
  HeapSlot htop[HTOP];
  double cp[3], cv[3], cg;
  double gp[3], gv[3];
  double sp[3], sv[3];
  double opt_time;
  double shot[12];
  unsigned long long pop_hash;
  unsigned long long cnt[8];
  uint32_t cur_id, cur_parent, epoch;
  int heap_len, sift_len, use_num, n_pop, n1, n2, n_new, status, flag, q, trunc;
  int wsum[KT / 32];
  // ---- everything above is per-query scratch: between two queries the in-kernel QP overlays it with its workspaces.
  // ---- everything below survives a QP round.
  unsigned long long mbar;
  KinoParamsDev P;
  MapDev M;
  unsigned long long ph[16];  // per-phase SM cycles [0..7] + diagnostics [8..15] of this CTA (thread 0's clock), only when bt.phase_cycles != nullptr
  long long ph_t;
  unsigned long long phq[16];
  float nzf[UAVMP_MAXNA], nz2[UAVMP_MAXNA];  // (float)(ua[c] + 9.81) and its square: the c-dependent part of the body axis
  int arena;                  // index of the arena this CTA took from the pool
  int npend;                  // finished queries' QP problems (3 q + axis) waiting for a round
  int pend[16];
};
#define QP_OVERLAY_BYTES (offsetof(SearchSmem, mbar))

__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz) {
  return (ax * bx + ay * by) + az * bz;  // Eigen fixed-size-3 reduction order
}

// (ut.dot(ut) + rou) * sample_tau (:231) of lattice point p, same association as the host table lat.ginc
__device__ __forceinline__ double ginc_of(const KinoParamsDev& P, int p) {
  const int na = P.na;
  const double ux = P.ua[p / (na * na)], uy = P.ua[(p / na) % na], uz = P.ua[p % na];
  return ((ux * ux + uy * uy) + uz * uz + P.rou) * P.tau;
}
__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v) {
  h ^= v;
  h *= 0x100000001b3ull;
  h ^= h >> 29;
  return h;
}

// ---- OBVP heuristic (kino_astar.cpp:312-414) ----------------------------------------------------------
__device__ int d_cubic(double a, double b, double c, double d, double* dts) {
  int n = 0;
  double a2 = b / a;
  double a1 = c / a;
  double a0 = d / a;
  double Q = (3 * a1 - a2 * a2) / 9;
  double R = (9 * a1 * a2 - 27 * a0 - 2 * a2 * a2 * a2) / 54;
  double D = Q * Q * Q + R * R;
  if (D > 0) {
    double S = fpm::cbrt(R + sqrt(D));
    double T = fpm::cbrt(R - sqrt(D));
    dts[n++] = -a2 / 3 + (S + T);
  } else if (D == 0) {
    double S = fpm::cbrt(R);
    dts[n++] = -a2 / 3 + S + S;
    dts[n++] = -a2 / 3 - S;
  } else {
    double theta = fpm::acos(R / sqrt(-Q * Q * Q));
    dts[n++] = 2 * sqrt(-Q) * fpm::cos(theta / 3) - a2 / 3;
    dts[n++] = 2 * sqrt(-Q) * fpm::cos((theta + 2 * M_PI) / 3) - a2 / 3;
    dts[n++] = 2 * sqrt(-Q) * fpm::cos((theta + 4 * M_PI) / 3) - a2 / 3;
  }
  return n;
}

__device__ int d_quartic(double a, double b, double c, double d, double e, double* dts) {
  int n = 0;
  double a3 = b / a;
  double a2 = c / a;
  double a1 = d / a;
  double a0 = e / a;
  double ys[3];
  d_cubic(1, -a2, a1 * a3 - 4 * a0, 4 * a2 * a0 - a1 * a1 - a3 * a3 * a0, ys);
  double y1 = ys[0];
  double r = a3 * a3 / 4 - a2 + y1;
  if (r < 0) return 0;
  double R = sqrt(r);
  double D, E;
  if (R != 0) {
    D = sqrt(0.75 * a3 * a3 - R * R - 2 * a2 + 0.25 * (4 * a3 * a2 - 8 * a1 - a3 * a3 * a3) / R);
    E = sqrt(0.75 * a3 * a3 - R * R - 2 * a2 - 0.25 * (4 * a3 * a2 - 8 * a1 - a3 * a3 * a3) / R);
  } else {
    D = sqrt(0.75 * a3 * a3 - 2 * a2 + 2 * sqrt(y1 * y1 - 4 * a0));
    E = sqrt(0.75 * a3 * a3 - 2 * a2 - 2 * sqrt(y1 * y1 - 4 * a0));
  }
  if (!(D != D)) {
    dts[n++] = -a3 / 4 + R / 2 + D / 2;
    dts[n++] = -a3 / 4 + R / 2 - D / 2;
  }
  if (!(E != E)) {
    dts[n++] = -a3 / 4 - R / 2 + E / 2;
    dts[n++] = -a3 / 4 - R / 2 - E / 2;
  }
  return n;
}

// returns tie_breaker * optimal_cost; topt < 0 when no root qualified (the reference then leaves
// optimal_time untouched)
__device__ double d_heuristic(const KinoParamsDev& P, double x1x, double x1y, double x1z, double v1x, double v1y,
                              double v1z, double x2x, double x2y, double x2z, double v2x, double v2y, double v2z,
                              double& topt) {
  double dpx = x2x - x1x, dpy = x2y - x1y, dpz = x2z - x1z;
  double optimal_cost = (double)(1 << 30);
  double a = -36 * dot3(dpx, dpy, dpz, dpx, dpy, dpz);
  double b = 24 * dot3(dpx, dpy, dpz, v1x + v2x, v1y + v2y, v1z + v2z);
  double c = -4 * (dot3(v1x, v1y, v1z, v1x, v1y, v1z) + dot3(v1x, v1y, v1z, v2x, v2y, v2z) +
                   dot3(v2x, v2y, v2z, v2x, v2y, v2z));
  double d = 0;
  double e = P.rou;
  double dts[4];
  int n = d_quartic(e, d, c, b, a, dts);
  double T_bar = fmax(fmax(fabs(x1x - x2x), fabs(x1y - x2y)), fabs(x1z - x2z)) / P.vmax;
  topt = -1.0;
  for (int i = 0; i < n; i++) {
    double t = dts[i];
    double tmp_cost = a / (-3 * t * t * t) + b / (-2 * t * t) + c / (-1 * t) + e * t;
    if (tmp_cost < optimal_cost && t > T_bar && tmp_cost > 0) {
      optimal_cost = tmp_cost;
      topt = t;
    }
  }
  return P.tie * optimal_cost;
}

// ---- grid lookups (grid_map.h:350-404) ------------------------------------------------------------------
__device__ __forceinline__ bool in_map(const MapDev& M, double x, double y, double z) {
  if (x < M.lox || y < M.loy || z < M.loz) return false;
  if (x > M.hix || y > M.hiy || z > M.hiz) return false;
  return true;
}
__device__ __forceinline__ void pos_to_index(const MapDev& M, double x, double y, double z, int& ix, int& iy, int& iz) {
  ix = (int)floor((x - M.ox) * M.inv_res);
  iy = (int)floor((y - M.oy) * M.inv_res);
  iz = (int)floor((z - M.oz) * M.inv_res);
}
__device__ __forceinline__ unsigned map_flags(const MapDev& M, int ix, int iy, int iz) {
  return __ldg(M.flags + ((((size_t)ix * (M.nzp >> 4) + (iz >> 4)) * M.ny + iy) << 4) + (iz & 15));
}

__device__ __forceinline__ unsigned long long pack_key(int ix, int iy, int iz, bool& ok) {
  unsigned ux = (unsigned)(ix + KEY_BIAS), uy = (unsigned)(iy + KEY_BIAS), uz = (unsigned)(iz + KEY_BIAS);
  ok = (ux < (1u << 18)) && (uy < (1u << 18)) && (uz < (1u << 18));
  return ((unsigned long long)ux << 36) | ((unsigned long long)uy << 18) | (unsigned long long)uz;
}
__device__ __forceinline__ uint32_t hash_key(unsigned long long k, int bits) {
  return (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> (64 - bits));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- open list: libstdc++ heap semantics on cached keys -----------------------------------------------
// slot i (0-based) is stored at heap[i + 1]; every move also records the new position in the node's hash slot
#define HS_DIRTY 0x80000000u
#define HS_MASK 0x7fffffffu

__device__ __forceinline__ HeapSlot hload(const SearchSmem& s, const HeapSlot* H, int i) { return i < HTOP ? s.htop[i] : H[i]; }
__device__ __forceinline__ void hstore(SearchSmem& s, HeapSlot* H, int i, const HeapSlot& v) { if (i < HTOP) s.htop[i] = v; else H[i] = v; }
__device__ __forceinline__ void hstore_key(SearchSmem& s, HeapSlot* H, int i, double f) { if (i < HTOP) s.htop[i].f = f; else __stcg(&H[i].f, f); }

// std::pop_heap + pop_back (bits/stl_heap.h __pop_heap -> __adjust_heap -> __push_heap), comp(a,b) = f[a] > f[b], executed
// by warp 0.  Indices below are 1-based (slot j lives at H[j]): the hole starts at 1 and moves to the child with the
// smaller key (the right one on ties) while both children exist (2h+1 <= n), a lone left child (2h == n) moves up
// unconditionally, and the former last element is then pushed up from the final hole.
// The sift-down is a chain of dependent loads, one L2 round trip per level once it leaves the shared-memory top of the
// heap.  The warp therefore fetches the whole cone of the next three levels below the hole (2 + 4 + 8 slots, lanes
// 0..13) in ONE round trip and resolves the three comparisons with shuffles; the lane that holds a chosen slot writes
// it to its parent position.  No slot written during the sift-down is read again by it (writes go to ancestors of
// the hole, reads to its descendants).
__device__ void heap_pop_warp(SearchSmem& s, HeapSlot* H, HashSlot* table, int& len, HeapSlot& top, int lane) {
  top = s.htop[1];
  const int old_len = len;
  len = old_len - 1;
  if (old_len <= 1) return;
  const int n = len;
  const HeapSlot value = hload(s, H, old_len);  // the last element (same address in every lane: one transaction)
  __syncwarp();  // every lane has read the old top before lane 0 overwrites it
  int hole = 1;
  double moved_f = 0.0;   // key of the slot moved in the last step == the parent of the current hole
  bool moved = false;
  // levels whose children both live in shared memory: every lane walks the same path, lane 0 stores
  while (2 * hole + 1 <= n && 2 * hole + 1 < HTOP) {
    const HeapSlot l = s.htop[2 * hole], r = s.htop[2 * hole + 1];
    const bool right = !(r.f > l.f);
    const HeapSlot pick = right ? r : l;
    if (lane == 0) { s.htop[hole] = pick; table[pick.hs].heap_pos = (uint32_t)(hole - 1); }
    hole = 2 * hole + (right ? 1 : 0);
    moved_f = pick.f; moved = true;
  }
  // deeper levels, three per round trip
  const int lvl = lane < 2 ? 1 : (lane < 6 ? 2 : 3);
  const int off = lane < 2 ? lane : (lane < 6 ? lane - 2 : lane - 6);
  while (2 * hole + 1 <= n) {
    const long long idx = ((long long)hole << lvl) + off;
    HeapSlot e;
    e.f = 0.0; e.id = 0; e.hs = 0;
    if (lane < 14 && idx <= (long long)n) e = hload(s, H, (int)idx);
    int cur = hole, rel = 0, lbase = 0;  // lbase: first lane of the level below `cur` (0, 2, 6)
#pragma unroll
    for (int step = 0; step < 3; step++) {
      if (2 * cur + 1 > n) break;
      const int lane_l = lbase + 2 * rel;
      const double fl = __shfl_sync(FULL, e.f, lane_l), fr = __shfl_sync(FULL, e.f, lane_l + 1);
      const bool right = !(fr > fl);
      if (lane == lane_l + (right ? 1 : 0)) { hstore(s, H, cur, e); table[e.hs].heap_pos = (uint32_t)(cur - 1); }
      moved_f = right ? fr : fl; moved = true;
      cur = 2 * cur + (right ? 1 : 0);
      rel = 2 * rel + (right ? 1 : 0);
      lbase = 2 * lbase + 2;
    }
    hole = cur;
  }
  __syncwarp();
  if (lane == 0) {
    if (2 * hole == n) {  // a lone left child
      const HeapSlot l = hload(s, H, n);
      hstore(s, H, hole, l);
      table[l.hs].heap_pos = (uint32_t)(hole - 1);
      hole = n;
      moved_f = l.f; moved = true;
    }
    // __push_heap of the former last element from the hole; its first parent is the slot that was just moved
    if (hole > 1 && (!moved || moved_f > value.f)) {
      int parent = hole / 2;
      while (hole > 1) {
        const HeapSlot pe = hload(s, H, parent);
        if (!(pe.f > value.f)) break;
        hstore(s, H, hole, pe);
        table[pe.hs].heap_pos = (uint32_t)(hole - 1);
        hole = parent;
        parent = hole / 2;
      }
    }
    hstore(s, H, hole, value);
    table[value.hs].heap_pos = (uint32_t)(hole - 1);
  }
  __syncwarp();
}

// closure = the ancestors of leaves len0+1 .. len0+m (1-based heap indices), staged in shared memory by warp 0.
// The caller keeps every batch inside ONE level of the tree (all new leaves at the same depth), so the ancestors
// `d` levels up form the contiguous range [(len0+1) >> d, (len0+m) >> d] and ranges of different d never share a node.
// Lane d owns level d: its range start `lo` and its offset `off` in the staging array stay in registers.
struct Closure {
  int lo, off, lo_m1, off_m1, total;
  bool active;
};

__device__ __forceinline__ void closure_load(SearchSmem& s, const HeapSlot* H, int len0, int m, int lane, Closure& c) {
  const int lo = (len0 + 1) >> lane, hi = (len0 + m) >> lane;
  const int cnt = (lo >= 1) ? hi - lo + 1 : 0;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
  c.lo = lo; c.off = incl - cnt;
  c.total = __shfl_sync(FULL, incl, 31);
  c.lo_m1 = __shfl_up_sync(FULL, c.lo, 1);
  c.off_m1 = __shfl_up_sync(FULL, c.off, 1);
  c.active = true;
  {  // level 0 = the new leaves themselves (offset 0, m entries): initialised by the whole warp
    for (int j = lane; j < m; j += 32) { HeapSlot e; e.f = 0; e.id = UAVMP_NONE; e.hs = 0; s.b.hc[j] = e; s.b.hidx[j] = (uint32_t)(len0 + 1 + j); }
  }
  // ancestors, level by level, one entry per lane (a level of a 256-leaf batch has up to 128 entries: copying it from its
  // owner lane alone was a 128-trip serial loop)
  const int depth = 32 - __clz(len0 + m);  // levels 1 .. depth-1 hold ancestors
  for (int d = 1; d < depth; d++) {
    const int lo_d = __shfl_sync(FULL, lo, d), cnt_d = __shfl_sync(FULL, cnt, d), off_d = __shfl_sync(FULL, c.off, d);
    for (int j = lane; j < cnt_d; j += 32) {
      // asynchronous 16 B global -> shared copies (L2 only): all of a level's loads are in flight at once
      if (lo_d + j < HTOP) s.b.hc[off_d + j] = s.htop[lo_d + j];
      else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&s.b.hc[off_d + j])), "l"(H + lo_d + j) : "memory");
      s.b.hidx[off_d + j] = (uint32_t)(lo_d + j);
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
}

__device__ __forceinline__ void closure_flush(SearchSmem& s, HeapSlot* H, HashSlot* table, int lane, Closure& c) {
  if (!c.active) return;
  __syncwarp();
  for (int e = lane; e < c.total; e += 32) {
    HeapSlot ent = s.b.hc[e];
    if (ent.hs & HS_DIRTY) {
      const uint32_t idx = s.b.hidx[e];
      ent.hs &= HS_MASK;
      hstore(s, H, (int)idx, ent);
      table[ent.hs].heap_pos = idx - 1;
    }
  }
  __threadfence_block();
  __syncwarp();
  c.active = false;
}

// number of leaves that can be pushed from length `len` without leaving the current tree level
__device__ __forceinline__ int level_room(int len) {
  const int x = len + 1;                // 1-based index of the next leaf
  const int depth = 31 - __clz(x);
  return (2 << depth) - x;              // leaves x .. 2^(depth+1) - 1
}

// ---- the push through a register-resident ancestor chain -------------------------------------------------------------------
// Round 1's push read and wrote the staged ancestors in shared memory for every push (~100 dependent instructions,
// measured ~800 cycles per push: the ordered commit was 41 % of all CTA time).  The chain of the CURRENT leaf n1 is kept in
// registers instead: lane d caches the staged entry of ancestor n1 >> d (lane 0: the leaf itself).  A push that moves L
// ancestors down is then one ballot plus a one-lane shuffle of the chain (lane j - 1 takes lane j's entry, lane L the new
// node); moving on to leaf n1 + 1 changes the cached index only on the lanes d <= ctz(n1 + 1), which write their entry back to
// the staging array (if modified) and fetch their next one.  Indices of one level only grow during a batch, so an entry that
// left the cache is never read again before closure_flush.
struct ChainCache {
  HeapSlot e;  // the cached staging entry (lanes above the root: f = -inf, never moves)
  int ix;      // its index in s.b.hc
};
__device__ __forceinline__ void chain_writeback(SearchSmem& s, ChainCache& cc) {
  if (cc.e.hs & HS_DIRTY) s.b.hc[cc.ix] = cc.e;
  cc.e.hs = 0;
  __syncwarp();
}
__device__ __forceinline__ void chain_fill(const SearchSmem& s, const Closure& c, ChainCache& cc, int n1, int lane) {
  const int a = n1 >> lane;
  cc.e.f = -INFINITY; cc.e.id = 0; cc.e.hs = 0; cc.ix = 0;
  if (lane == 0) cc.ix = c.off + (a - c.lo);  // the leaf's own slot, written when the chain moves on
  else if (a >= 1) { cc.ix = c.off + (a - c.lo); cc.e = s.b.hc[cc.ix]; }
}
// std::push_heap of (f, id, hs) as the leaf the chain is filled for.  lm = bits 0 .. lane, a per-lane constant: with
// cont = (ballot of "ancestor key > f") >> 1, the number L of consecutive ancestors that move down satisfies
// lane < L <=> (cont & lm) == lm and lane == L <=> (cont & lm) == lm >> 1 — no find-first-set on the critical path.
__device__ __forceinline__ void chain_push(ChainCache& cc, double f, uint32_t id, uint32_t hs, unsigned lm) {
  const unsigned cont = __ballot_sync(FULL, cc.e.f > f) >> 1;
  HeapSlot up;
  up.f = __shfl_down_sync(FULL, cc.e.f, 1);
  up.id = __shfl_down_sync(FULL, cc.e.id, 1);
  up.hs = __shfl_down_sync(FULL, cc.e.hs, 1) | HS_DIRTY;
  const unsigned cl = cont & lm;
  if (cl == lm) cc.e = up;
  else if (cl == (lm >> 1)) { cc.e.f = f; cc.e.id = id; cc.e.hs = hs | HS_DIRTY; }
}
// the chain of leaf n1 becomes the chain of leaf n1 + 1 (same tree level, same closure): ancestor n1 >> d changes on the
// lanes d <= ctz(n1 + 1); the entries of a level are contiguous in the staging array, so the next one is the next slot
__device__ __forceinline__ void chain_advance(SearchSmem& s, ChainCache& cc, int n1_next, int lane) {
  if (lane < __ffs(n1_next)) {
    s.b.hc[cc.ix] = cc.e;
    cc.ix++;
    cc.e = s.b.hc[cc.ix];
  }
}

// in-place key mutation (kino_astar.cpp:251-265) of a node that may sit among the staged ancestors: look for it in the
// closure; if it is not there, write through its live heap position
__device__ void heap_set_key_slow(SearchSmem& s, HeapSlot* H, const HashSlot* table, uint32_t id, uint32_t hs, double f,
                                  int lane, const Closure c) {
  __syncwarp();
  int found = -1;
  if (c.active) {
    for (int base = 0; base < c.total; base += 32) {
      const int j = base + lane;
      const bool hit = (j < c.total) && (s.b.hc[j].id == id);
      const unsigned m = __ballot_sync(FULL, hit);
      if (m) { found = base + __ffs(m) - 1; break; }
    }
  }
  if (lane == 0) {
    if (found >= 0) {
      s.b.hc[found].f = f;
      s.b.hc[found].hs |= HS_DIRTY;
    } else {
      __threadfence_block();
      const uint32_t pos = __ldcg(&table[hs].heap_pos);
      hstore_key(s, H, (int)pos + 1, f);
    }
  }
  __syncwarp();
}

__device__ void shot_pos(const double* c, double t, double& x, double& y, double& z) {
  // shot_coef_ * [1 t t^2 t^3]^T with pow(t, i) (kino_astar.cpp:441-446)
  double t0 = fpm::powi(t, 0), t1 = fpm::powi(t, 1), t2 = fpm::powi(t, 2), t3 = fpm::powi(t, 3);
  x = ((c[0] * t0 + c[1] * t1) + c[2] * t2) + c[3] * t3;
  y = ((c[4] * t0 + c[5] * t1) + c[6] * t2) + c[7] * t3;
  z = ((c[8] * t0 + c[9] * t1) + c[10] * t2) + c[11] * t3;
}

// ---- TMA / mbarrier primitives (sm_90+ PTX) -------------------------------------------------------------
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, void* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// ---- ellipsoid test of one (primitive, checkpoint) against the cloud points staged in shared memory ---------
// warp-cooperative: lanes take the points of one cell column at a time; kino_astar.cpp:721-758 semantics
// ("any cloud point p with || E^-1 (p - pt) || <= 1"; the KD-tree radius r + 0.1 is a pure superset filter, §9.1 Q12)
// Exact decision of kino_astar.cpp:749-753 for one cloud point, reached through two conservative filters:
//   1. float distance cull: a point farther than the longest semi-axis (+2 mm) cannot be inside;
//   2. because the ellipsoid is diag(r, r, h) in the body frame, || E^-1 d ||^2 == (|d|^2 - w^2)/r^2 + w^2/h^2 with
//      w = d . b3 in exact arithmetic; its f64 evaluation decides every point that is not within 1e-9 of the surface;
//   3. only those borderline points take the reference's own expression: sqrt(|E^-1 d|^2) <= 1 with Eigen's association.
struct EllipsoidTest {
  float fx, fy, fz, cullf;
  double px, py, pz, b3x, b3y, b3z, inv_r2, inv_h2, cull2;
  const double* e;  // E^-1, row-major, global
};
__device__ __forceinline__ bool point_hits(const EllipsoidTest& t, float4 q) {
  const float dx = q.x - t.fx, dy = q.y - t.fy, dz = q.z - t.fz;
  if (dx * dx + dy * dy + dz * dz > t.cullf) return false;
  const double ddx = (double)q.x - t.px, ddy = (double)q.y - t.py, ddz = (double)q.z - t.pz;
  const double dd = (ddx * ddx + ddy * ddy) + ddz * ddz;
  if (dd > t.cull2) return false;
  const double w = (ddx * t.b3x + ddy * t.b3y) + ddz * t.b3z;
  const double w2 = w * w;
  const double sv = (dd - w2) * t.inv_r2 + w2 * t.inv_h2;
  if (sv < 1.0 - 1e-9) return true;
  if (sv > 1.0 + 1e-9) return false;
  const double* e = t.e;
  const double tx = (__ldg(e + 0) * ddx + __ldg(e + 1) * ddy) + __ldg(e + 2) * ddz;
  const double ty = (__ldg(e + 3) * ddx + __ldg(e + 4) * ddy) + __ldg(e + 5) * ddz;
  const double tz = (__ldg(e + 6) * ddx + __ldg(e + 7) * ddy) + __ldg(e + 8) * ddz;
  return sqrt(dot3(tx, ty, tz, tx, ty, tz)) <= 1.0;
}
__device__ __forceinline__ void make_test(EllipsoidTest& t, const KinoParamsDev& P, const LatticeDev& lat, int p, double px,
                                          double py, double pz) {
  t.px = px; t.py = py; t.pz = pz;
  t.fx = (float)px; t.fy = (float)py; t.fz = (float)pz;
  t.cullf = P.cullf; t.cull2 = P.cull2; t.inv_r2 = P.inv_r2; t.inv_h2 = P.inv_h2;
  t.b3x = __ldg(lat.b3 + 3 * p); t.b3y = __ldg(lat.b3 + 3 * p + 1); t.b3z = __ldg(lat.b3 + 3 * p + 2);
  t.e = lat.Einv + 9 * p;
}

// fallback when the region's points do not fit in shared memory: the lane-serial cell-list walk over global memory
__device__ bool ellipsoid_hit_global(const MapDev& M, const KinoParamsDev& P, const EllipsoidTest& t, unsigned& n_tested) {
  const double px = t.px, py = t.py, pz = t.pz;
  int x0 = max((int)floor((px - P.box_r - M.cox) * M.inv_cell), 0);
  int x1 = min((int)floor((px + P.box_r - M.cox) * M.inv_cell), M.cnx - 1);
  int y0 = max((int)floor((py - P.box_r - M.coy) * M.inv_cell), 0);
  int y1 = min((int)floor((py + P.box_r - M.coy) * M.inv_cell), M.cny - 1);
  int z0 = max((int)floor((pz - P.box_r - M.coz) * M.inv_cell), 0);
  int z1 = min((int)floor((pz + P.box_r - M.coz) * M.inv_cell), M.cnz - 1);
  for (int ix = x0; ix <= x1; ix++)
    for (int iy = y0; iy <= y1; iy++) {
      int cbase = (ix * M.cny + iy) * M.cnz;
      int k0 = __ldg(M.cell_start + cbase + z0), k1 = __ldg(M.cell_start + cbase + z1 + 1);
      for (int k = k0; k < k1; k++) {
        n_tested++;
        if (point_hits(t, __ldg(M.pts + k))) return true;
      }
    }
  return false;
}

// =====================================================================================================
// One in-kernel QP round: warps 0 .. cnt-1 each take one pending 1-D problem (id = 3 q + axis) of a query this CTA has
// finished, build its waypoints from the staged path (the rule of uav_motion_planning_b200/planner.py: waypoint k =
// path[floor(k (n - 1) / S)], T_i = seg_time, boundary velocity = start / end velocity, acceleration and jerk 0) and run
// qp_warp_solve_one on a workspace that overlays the search's per-query shared memory.
__device__ __forceinline__ void qp_round(SearchSmem& s, unsigned char* smem_raw, const QpPlanDev& pl, const KinoQpDev& qp,
                                         const uavmp_osqp_settings& S, const KinoBatchDev& bt, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  const int cnt = min(s.npend, qp.warps);
  __syncthreads();  // every thread has left the scratch area and has read npend
  // the index block of the triangular solves (uint16), shared by the round's warps, behind their workspaces
  unsigned short* sx = reinterpret_cast<unsigned short*>(reinterpret_cast<double*>(smem_raw) + (size_t)qp.warps * pl.ws_warp);
  for (int i = tid; i < pl.n_sidx / 2; i += KT) reinterpret_cast<uint32_t*>(sx)[i] = reinterpret_cast<const uint32_t*>(pl.Sidx)[i];
  __syncthreads();
  if (warp < cnt) {
    const int b = s.pend[s.npend - 1 - warp];
    const int q = b / 3, ax = b - 3 * q;
    const int Sg = qp.Sg;
    const int np = bt.n_path[q];
    const double* path = bt.path_stage + (size_t)q * bt.path_cap * 3;
    for (int k = lane; k <= Sg; k += 32) qp.pos[(size_t)b * (Sg + 1) + k] = path[3 * (((long long)k * (np - 1)) / Sg) + ax];
    for (int sgm = lane; sgm < Sg; sgm += 32) {
      const long long i0 = ((long long)sgm * (np - 1)) / Sg, i1 = ((long long)(sgm + 1) * (np - 1)) / Sg;
      // time allocation (uavmp_plan_options): the searched trajectory's own timing, one path sample every time_step_size
      qp.T[(size_t)b * Sg + sgm] = qp.time_alloc ? (double)(i1 > i0 ? i1 - i0 : 1) * qp.step : qp.seg_time;
      if (qp.Kc > 0) {  // corridor box of the segment on this axis: its path points' extent +- margin (min / max: order-free)
        double mn = path[3 * i0 + ax], mx = mn;
        for (long long i = i0 + 1; i <= i1; i++) { const double v = path[3 * i + ax]; mn = fmin(mn, v); mx = fmax(mx, v); }
        qp.lo[(size_t)b * Sg + sgm] = mn - qp.margin;
        qp.hi[(size_t)b * Sg + sgm] = mx + qp.margin;
      }
    }
    if (lane == 0) {
      qp.bv[(size_t)b * 2] = bt.start_vel[3 * q + ax]; qp.bv[(size_t)b * 2 + 1] = bt.end_vel[3 * q + ax];
      qp.ba[(size_t)b * 2] = 0.0; qp.ba[(size_t)b * 2 + 1] = 0.0;
      qp.bj[(size_t)b * 2] = 0.0; qp.bj[(size_t)b * 2 + 1] = 0.0;
    }
    __syncwarp();
    QpIo io;
    io.pos = qp.pos; io.bv = qp.bv; io.ba = qp.ba; io.bj = qp.bj; io.T = qp.T; io.lo = qp.lo; io.hi = qp.hi;
    io.coef = qp.coef; io.solved = qp.solved3; io.status = qp.status3; io.iters = qp.iters3; io.B = 3 * bt.B; io.stride = 0;
    qp_warp_solve_one(pl, io, S, reinterpret_cast<double*>(smem_raw) + (size_t)warp * pl.ws_warp, b, sx);
    if (lane == 0 && !qp.solved3[b]) atomicAnd(qp.qp_solved + q, 0);
  }
  __syncthreads();
  if (tid == 0) s.npend -= cnt;
  __syncthreads();
}

// =====================================================================================================
__global__ void __launch_bounds__(KT, (KT <= 256 ? 2 : 1)) kino_search_kernel(const KinoParamsDev* __restrict__ Pp, LatticeDev lat,
                                                            const MapDev* __restrict__ Mp, KinoArena* arenas, int* arena_busy,
                                                            int n_arenas, const __grid_constant__ KinoBatchDev bt, int table_bits,
                                                            const __grid_constant__ CUtensorMap tmap, int use_tma, float slab_margin, float cullf_arg,
                                                            const __grid_constant__ KinoQpDev qp, const __grid_constant__ QpPlanDev pl,
                                                            const __grid_constant__ uavmp_osqp_settings qps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SearchSmem& s = *reinterpret_cast<SearchSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // parameter blocks live in shared memory: every hot loop reads them
  for (int i = tid; i < (int)(sizeof(KinoParamsDev) / 4); i += KT) reinterpret_cast<uint32_t*>(&s.P)[i] = reinterpret_cast<const uint32_t*>(Pp)[i];
  for (int i = tid; i < (int)(sizeof(MapDev) / 4); i += KT) reinterpret_cast<uint32_t*>(&s.M)[i] = reinterpret_cast<const uint32_t*>(Mp)[i];
  if (tid == 0) {
    // Take a free arena from the pool.  Search kernels of several batches run concurrently (cross-batch pipelining: the
    // CTAs of batch k + 1 become resident as the CTAs of batch k retire), so an arena belongs to a CTA, not to a block index.
    // At most n_arenas search CTAs are ever resident (occupancy), so a free one exists unless the pool was cut for memory.
    int a = (int)(blockIdx.x % (unsigned)n_arenas);
    for (;;) {
      if (atomicCAS(arena_busy + a, 0, 1) == 0) break;
      a = (a + 1 == n_arenas) ? 0 : a + 1;
    }
    // acquire: the previous owner's writes (epoch word, hash table, possibly made on another SM) are visible to the loads
    // below, including the L1-cached ones
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    s.arena = a;
  }
  __syncthreads();
  if (tid == 0) s.P.cullf = cullf_arg;  // the float pre-cull's margin depends on the map's coordinate magnitude (host: float_filter_margins)
  __syncthreads();
  const KinoParamsDev& P = s.P;
  const MapDev& M = s.M;
  KinoArena ar = arenas[s.arena];
  KinoNode* nodes = ar.nodes;
  HeapSlot* H = ar.heap;
  HashSlot* table = ar.table;
  const uint32_t tmask = (1u << table_bits) - 1u;
  const bool prof = bt.phase_cycles != nullptr;
  const int na = P.na, K = P.K, nprim = P.nprim;
  if (tid < 16) { s.ph[tid] = 0; s.phq[tid] = 0; }
  if (tid < UAVMP_MAXNA) {
    const float nz = (tid < na) ? (float)(P.ua[tid] + 9.81) : 1.f;  // a3 = u + 9.81 e_z (kino_astar.cpp:724)
    s.nzf[tid] = nz; s.nz2[tid] = nz * nz;
  }
  if (tid == 0) {
    s.ph_t = clock64(); s.npend = 0;
    mbar_init(&s.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t tma_parity = 0;
  __syncthreads();

  for (;;) {
    // ---- fetch the next query --------------------------------------------------------------------
    __syncthreads();
    if (tid == 0) {
      int w = atomicAdd(bt.next_query, 1);
      s.q = (w < bt.B) ? (bt.order ? bt.order[w] : w) : -1;
    }
    __syncthreads();
    const int q = s.q;
    if (q < 0) {
      while (qp.enabled && s.npend > 0) qp_round(s, smem_raw, pl, qp, qps, bt, tid);  // (uniform: npend is re-read after barriers)
      break;
    }

    if (tid < 8) s.cnt[tid] = 0;
    if (tid == 0) {
      uint32_t ep = *ar.epoch + 1;
      if (ep > EPOCH_MASK) ep = 0;  // wrapped: table must be cleared
      s.epoch = ep;
    }
    __syncthreads();
    if (s.epoch == 0) {
      for (size_t i = tid; i <= tmask; i += KT) table[i].key = 0ull;
      __syncthreads();
      if (tid == 0) s.epoch = 1;
      __syncthreads();
    }
    const uint32_t epoch = s.epoch;
    if (tid == 0) {
      *ar.epoch = epoch;
      for (int i = 0; i < 3; i++) {
        s.sp[i] = bt.start_pt[3 * q + i]; s.sv[i] = bt.start_vel[3 * q + i];
        s.gp[i] = bt.end_pt[3 * q + i];   s.gv[i] = bt.end_vel[3 * q + i];
      }
      // start node (kino_astar.cpp:85-97)
      double topt;
      double h = d_heuristic(P, s.sp[0], s.sp[1], s.sp[2], s.sv[0], s.sv[1], s.sv[2], s.gp[0], s.gp[1], s.gp[2],
                             s.gv[0], s.gv[1], s.gv[2], topt);
      s.opt_time = (topt >= 0.0) ? topt : (double)(1 << 30);
      int ix, iy, iz; bool ok;
      pos_to_index(M, s.sp[0], s.sp[1], s.sp[2], ix, iy, iz);
      unsigned long long k = pack_key(ix, iy, iz, ok);
      if (!ok) atomicOr(bt.error_flag, 1);
      uint32_t hh = hash_key(k, table_bits);
      while ((table[hh].key & EPOCH_MASK) == epoch) hh = (hh + 1) & tmask;  // cannot happen on a fresh epoch
      HashSlot hsl;
      hsl.key = (k << EPOCH_BITS) | epoch; hsl.id = 0; hsl.heap_pos = 0; hsl.g = 0.0; hsl.closed = 0; hsl.pad = 0;
      table[hh] = hsl;
      KinoNode nd;
      nd.px = s.sp[0]; nd.py = s.sp[1]; nd.pz = s.sp[2]; nd.vx = s.sv[0]; nd.vy = s.sv[1]; nd.vz = s.sv[2];
      nd.g = 0.0; nd.parent = UAVMP_NONE; nd.hslot = hh; nd.input = 0; nd.closed = 0; nd.pad0 = 0; nd.pad1 = 0;
      nodes[0] = nd;
      HeapSlot hs; hs.f = P.lambda * h; hs.id = 0; hs.hs = hh;
      hstore(s, H, 1, hs);
      s.heap_len = 1; s.use_num = 1; s.n_pop = 0; s.status = 0; s.trunc = 0;
      s.pop_hash = 0xcbf29ce484222325ull;
      s.cnt[4] += 1; s.cnt[6] += 1;  // insert and heuristic of the start node (its expanded_list_.insert is not a lookup, :96)
    }
    unsigned my_occ = 0, my_cloud = 0;
    __syncthreads();
    PH_MARK(7);
    long long q_t0 = prof ? clock64() : 0;

    // ================================ main loop (kino_astar.cpp:101) ===============================
    for (;;) {
      // ---- pop --------------------------------------------------------------------------------------
      if (warp == 0) {
        const int len0 = s.heap_len;
        __syncwarp();
        if (len0 == 0) {
          if (lane == 0) s.status = UAVMP_NO_PATH_FOUND;  // open list exhausted (:270-271)
        } else {
          // Only the IDENTITY of the popped node is needed to start the expansion: the top of the open list and its record
          // (prefetched at the end of the previous commit).  The sift-down that restores the heap — a chain of dependent L2
          // round trips — is deferred to the start of phase A0, where it runs on this warp under the TMA flight of the flags
          // box and the other warps' table computation; nobody reads the open list before the classification (phase B).
          const HeapSlot top = s.htop[1];
          const int len = len0 - 1;
          unsigned long long nw = 0;
          if (lane >= 16 && lane < 16 + (int)(sizeof(KinoNode) / 8))
            nw = reinterpret_cast<const unsigned long long*>(nodes + top.id)[lane - 16];
          KinoNode nd;
          nd.px = __longlong_as_double((long long)__shfl_sync(FULL, nw, 16)); nd.py = __longlong_as_double((long long)__shfl_sync(FULL, nw, 17));
          nd.pz = __longlong_as_double((long long)__shfl_sync(FULL, nw, 18)); nd.vx = __longlong_as_double((long long)__shfl_sync(FULL, nw, 19));
          nd.vy = __longlong_as_double((long long)__shfl_sync(FULL, nw, 20)); nd.vz = __longlong_as_double((long long)__shfl_sync(FULL, nw, 21));
          nd.g = __longlong_as_double((long long)__shfl_sync(FULL, nw, 22));
          nd.parent = (uint32_t)(__shfl_sync(FULL, nw, 23) & 0xffffffffull);
          if (lane == 0) {
          s.heap_len = len;
          s.sift_len = len0;
          nodes[top.id].closed = 1;
          table[top.hs].closed = 1;
          s.cur_id = top.id; s.cur_parent = nd.parent;
          s.cp[0] = nd.px; s.cp[1] = nd.py; s.cp[2] = nd.pz; s.cv[0] = nd.vx; s.cv[1] = nd.vy; s.cv[2] = nd.vz;
          s.cg = nd.g;
          int ix, iy, iz;
          pos_to_index(M, nd.px, nd.py, nd.pz, ix, iy, iz);
          // the reference never updates node->index on an in-place mutation (:256 is commented out), but the mutated
          // position lies in the same voxel by construction (it was found through that key), so this is identical
          unsigned long long ph = s.pop_hash;
          ph = mix64(ph, (unsigned long long)(uint32_t)ix);
          ph = mix64(ph, (unsigned long long)(uint32_t)iy);
          ph = mix64(ph, (unsigned long long)(uint32_t)iz);
          ph = mix64(ph, fpm::to_bits(nd.px));
          ph = mix64(ph, fpm::to_bits(nd.vx));
          ph = mix64(ph, fpm::to_bits(nd.g));
          s.pop_hash = ph;
          if (bt.pop_trace && s.n_pop < bt.pop_cap) {
            int* tr = bt.pop_trace + ((size_t)q * bt.pop_cap + s.n_pop) * 3;
            tr[0] = ix; tr[1] = iy; tr[2] = iz;
          }
          s.n_pop++;
          double dx = nd.px - s.gp[0], dy = nd.py - s.gp[1], dz = nd.pz - s.gp[2];
          s.flag = (sqrt(dot3(dx, dy, dz, dx, dy, dz)) < P.goal_tol) ? 1 : 0;
          s.n1 = 0; s.n2 = 0; s.nT = 0; s.n_upd = 0; s.last_ev = -1;
          }
        }
      } else if (tid >= 32 && tid < 32 + 3 * UAVMP_MAXNA) {
        (&s.axok[0][0])[tid - 32] = 1;  // reset the per-axis feasibility flags for this expansion's tables
      }
      __syncthreads();
      PH_MARK(0);
      if (s.status) break;

      // ---- near goal: one-shot trajectory (:112-154, :416-471) ---------------------------------------
      if (s.flag) {
        const double td = s.opt_time;  // stale by design (§9.1 Q2)
        if (tid < 3) {
          double x1 = s.cp[tid], v1 = s.cv[tid], x2 = s.gp[tid], v2 = s.gv[tid];
          double dp = x2 - x1, dv = v2 - v1;
          s.shot[4 * tid + 0] = x1;
          s.shot[4 * tid + 1] = v1;
          s.shot[4 * tid + 2] = 0.5 * ((6 / (td * td)) * (dp - v1 * td) - (2 * dv) / td);
          s.shot[4 * tid + 3] = (1.0 / 6.0) * ((-12 / (td * td * td)) * (dp - v1 * td) + (6 * dv) / (td * td));
        }
        __syncthreads();
        const double segf = floor(td / P.step);
        const long long seg = (segf < 9.0e18) ? (long long)segf : (long long)9.0e18;
        bool blocked = false;
        for (long long base = 0; base <= seg; base += KT) {
          long long j = base + tid;
          bool coll = false;
          if (j <= seg) {
            double t = (double)j * P.step, x, y, z;
            shot_pos(s.shot, t, x, y, z);
            if (!in_map(M, x, y, z)) {
              coll = true;
            } else {
              int ix, iy, iz;
              pos_to_index(M, x, y, z, ix, iy, iz);
              my_occ++;
              coll = (map_flags(M, ix, iy, iz) & 2u) != 0;
            }
          }
          if (__syncthreads_or(coll ? 1 : 0)) { blocked = true; break; }
        }
        if (tid == 0) s.cnt[7] += 1;
        if (!blocked) {
          // ---- REACH_END: retrievePath + samplePath (:473-557) --------------------------------------
          if (tid == 0) {
            int n = 0;
            uint32_t c = s.cur_id;
            while (c != UAVMP_NONE && n < UAVMP_MAXPRIM) { s.id[n++] = c; c = nodes[c].parent; }
            if (c != UAVMP_NONE) { atomicOr(bt.error_flag, 2); s.trunc = 1; }
            s.n1 = n;
          }
          __syncthreads();
          const int nn = s.n1;
          const int segk = P.K - 1;  // floor(duration / step) samples per primitive, duration == sample_tau
          const long long shot_n = seg + 1;
          const long long total = (long long)(nn - 1) * segk + shot_n;
          if (total > bt.path_cap) { if (tid == 0) { atomicOr(bt.error_flag, 4); s.trunc = 1; } }
          const long long lim = total < bt.path_cap ? total : bt.path_cap;
          double* out = bt.path_stage + (size_t)q * bt.path_cap * 3;
          for (long long i = tid; i < lim; i += KT) {
            double x, y, z;
            if (i < (long long)(nn - 1) * segk) {
              int sgi = (int)(i / segk), j = (int)(i % segk);
              const KinoNode& cn = nodes[s.id[nn - 1 - sgi]];
              int inp = nodes[s.id[nn - 2 - sgi]].input;
              double ux = lat.ux[inp], uy = lat.uy[inp], uz = lat.uz[inp];
              double t = P.tk[j], h = P.hk[j];
              x = (cn.px + t * cn.vx) + h * ux;
              y = (cn.py + t * cn.vy) + h * uy;
              z = (cn.pz + t * cn.vz) + h * uz;
            } else {
              long long j = i - (long long)(nn - 1) * segk;
              shot_pos(s.shot, (double)j * P.step, x, y, z);
            }
            out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
          }
          if (tid == 0) { s.status = UAVMP_REACH_END; bt.n_path[q] = (int)lim; }
        } else if (s.cur_parent == UAVMP_NONE) {
          if (tid == 0) s.status = UAVMP_NO_PATH_FOUND;  // :148-152
        }
        __syncthreads();
        PH_MARK(1);
        if (s.status) break;
      }

      // ---- the deferred half of the pop: std::pop_heap's sift-down (see the pop phase), concurrent with A0 ----------------
      if (warp == 0) {
        int len = s.sift_len;
        HeapSlot top_unused;
        heap_pop_warp(s, H, table, len, top_unused, lane);
      }
      // ---- A0. separable tables: every checkpoint / end-state coordinate is (c + t v) + h u per axis, and u comes
      // from a tensor lattice, so there are only K * 3 * na distinct coordinates (StateTransit, :651-670) -----------
      if (warp == KT / 32 - 1) {
        // the flags box is requested first, from a conservative extent (all checkpoints, feasible or not: x is monotone in
        // u, so the extremes sit at the two end values of the lattice), so that the TMA runs under the table computation.
        // One lane per (axis, checkpoint), then a warp reduction per axis: the request leaves a few hundred cycles earlier
        // than from a single thread.
        int lo = INT_MAX, hi = INT_MIN;
        for (int j = lane; j < 3 * K; j += 32) {
          const int ax = j / K, i = j % K;
          const double org = ax == 0 ? M.ox : (ax == 1 ? M.oy : M.oz);
          const double b = s.cp[ax] + P.tk[i] * s.cv[ax];
          const int i0 = (int)floor(((b + P.hk[i] * P.ua[0]) - org) * M.inv_res);
          const int i1 = (int)floor(((b + P.hk[i] * P.ua[na - 1]) - org) * M.inv_res);
          lo = min(lo, min(i0, i1)); hi = max(hi, max(i0, i1));
        }
        const int ax_l = (lane < 3 * K) ? lane / K : 3;  // with 3 K <= 32 every lane serves one axis
        bool fits = use_tma != 0;
        int org3[3];
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
          int l = INT_MAX, h = INT_MIN;
          if (3 * K <= 32) {
            l = __reduce_min_sync(FULL, ax_l == ax ? lo : INT_MAX);
            h = __reduce_max_sync(FULL, ax_l == ax ? hi : INT_MIN);
          } else {  // lanes mix axes: recompute serially (not reached with the launch-file checkpoint count)
            const double org = ax == 0 ? M.ox : (ax == 1 ? M.oy : M.oz);
            for (int i = 0; i < K; i++) {
              const double b = s.cp[ax] + P.tk[i] * s.cv[ax];
              const int i0 = (int)floor(((b + P.hk[i] * P.ua[0]) - org) * M.inv_res);
              const int i1 = (int)floor(((b + P.hk[i] * P.ua[na - 1]) - org) * M.inv_res);
              l = min(l, min(i0, i1)); h = max(h, max(i0, i1));
            }
          }
          if (ax == 2) l &= ~15;  // the copy is blocked by 16 voxels in z (and UTMALDG needs 16 B aligned inner coordinates)
          if (h - l + 1 > (ax == 2 ? TBZ : TB)) fits = false;
          org3[ax] = l;
        }
        if (lane == 0) {
          s.to[0] = org3[0]; s.to[1] = org3[1]; s.to[2] = org3[2];
          s.tile_ok = fits ? 1 : 0;
          if (fits) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s.mbar, TB * TB * TBZ);
            tma_load_3d(s.a.tile, &tmap, org3[1] * 4, org3[2] >> 4, org3[0], &s.mbar);  // (y * 4 words, z block, x)
          }
        }
      }
      for (int e = tid; e < (K + 1) * 3 * na; e += KT) {
        const int i = e / (3 * na), ax = (e / na) % 3, a = e % na;
        const double u = P.ua[a];
        const double t = (i < K) ? P.tk[i] : P.tau, h = (i < K) ? P.hk[i] : P.htau;
        const double x = (s.cp[ax] + t * s.cv[ax]) + h * u;
        const double v = s.cv[ax] + t * u;
        const double org = ax == 0 ? M.ox : (ax == 1 ? M.oy : M.oz);
        const int idx = (int)floor((x - org) * M.inv_res);
        if (i < K) {
          const double lo = ax == 0 ? M.lox : (ax == 1 ? M.loy : M.loz), hi = ax == 0 ? M.hix : (ax == 1 ? M.hiy : M.hiz);
          const bool ok = !(x < lo) && !(x > hi) && !(v < -P.vmax) && !(v > P.vmax);
          s.X[i][ax][a] = x; s.XI[i][ax][a] = idx;
          if (!ok) s.axok[ax][a] = 0;  // feasible on this axis only if every checkpoint is (all writers store 0: benign)
        } else {
          s.EX[ax][a] = x; s.EV[ax][a] = v; s.EI[ax][a] = idx;
        }
      }
      __syncthreads();
      const bool tile_ok = s.tile_ok != 0;
      if (tile_ok) { mbar_wait(&s.mbar, tma_parity); tma_parity ^= 1u; }

      // ---- A1. grid + velocity + in-map for every primitive from shared memory (:172-211 minus the ellipsoid) ----
      int my_need = 0;
      {
        const int tx = s.to[0], ty = s.to[1], tz = s.to[2];
        for (int p = tid; p < nprim; p += KT) {
          const int a = p / (na * na), b = (p / na) % na, c = p % na;
          uint8_t st = ST_REJECT;
          unsigned need = 0;
          if (s.axok[0][a] && s.axok[1][b] && s.axok[2][c]) {
            bool ok = true;
            for (int i = 0; i < K; i++) {
              const int ix = s.XI[i][0][a], iy = s.XI[i][1][b], iz = s.XI[i][2][c];
              const unsigned fl = tile_ok ? s.a.tile[((((ix - tx) * (TBZ / 16) + ((iz - tz) >> 4)) * TB + (iy - ty)) << 4) + ((iz - tz) & 15)] : map_flags(M, ix, iy, iz);
              my_occ++;
              if (P.ctype == 1 && (fl & 1u)) { ok = false; break; }
              if (fl & 4u) need |= 1u << i;
            }
            if (ok) {
              st = ST_FEASIBLE;
              // the successor classification will probe this voxel's hash slot: start pulling its sector towards L2 now,
              // under the cloud tests (a primitive the ellipsoid test rejects costs one wasted prefetch)
              bool kok;
              const unsigned long long k = pack_key(s.EI[0][a], s.EI[1][b], s.EI[2][c], kok);
              asm volatile("prefetch.global.L2 [%0];" ::"l"(table + hash_key(k, table_bits)));
            }
          }
          s.state[p] = st;
          s.a.need[p] = (st == ST_FEASIBLE) ? (uint16_t)need : (uint16_t)0;
          if (st == ST_FEASIBLE && need) my_need = 1;
        }
      }
      const int any_need = __syncthreads_or(my_need);
      // work units of the cloud test: (checkpoint i, lattice column a, b) with the set of c whose checkpoint is flagged.
      // The nine centres of a unit share x and y, so one pass over the candidate points serves all of them.  Ordered
      // compaction keeps neighbouring units on neighbouring lanes (similar candidate sets -> similar trip counts).
      if (any_need) {
        const int nU = K * na * na;
        const int per = (nU + KT - 1) / KT;
        uint32_t mine[8];
        int cnt = 0;
        for (int j = 0; j < per && j < 8; j++) {
          const int u = tid * per + j;
          uint32_t m = 0;
          if (u < nU) {
            const int i = u / (na * na), ab = u % (na * na);
            for (int c = 0; c < na; c++) m |= (uint32_t)((s.a.need[ab * na + c] >> i) & 1u) << c;
          }
          mine[j] = m ? ((uint32_t)u | (m << 11)) : 0u;
          cnt += m ? 1 : 0;
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s.wsum[warp] = incl;
        __syncthreads();
        int off = incl - cnt;
        for (int w = 0; w < warp; w++) off += s.wsum[w];
        for (int j = 0; j < per && j < 8; j++) if (mine[j]) s.units[off++] = mine[j];
        if (tid == KT - 1) s.nT = off;
      }
      __syncthreads();
      PH_MARK(2);

      // ---- A2. SE(3) ellipsoid vs cloud for the checkpoints that pass near obstacles (:721-758) ----------------------
      const int nT = s.nT;
      if (nT > 0) {
        // stage the cell list of the region the feasible checkpoints can touch: cell_start entries, then the points
        if (tid < 3) {
          // position extent of the feasible checkpoints on this axis -> cell range of the staging region
          double xl = 1e300, xh = -1e300;
          for (int a = 0; a < na; a++)
            if (s.axok[tid][a])
              for (int i = 0; i < K; i++) { xl = fmin(xl, s.X[i][tid][a]); xh = fmax(xh, s.X[i][tid][a]); }
          s.xlo[tid] = xl; s.xhi[tid] = xh;
          const double co = tid == 0 ? M.cox : (tid == 1 ? M.coy : M.coz);
          const int cn = tid == 0 ? M.cnx : (tid == 1 ? M.cny : M.cnz);
          s.rc0[tid] = max((int)floor((xl - P.box_r - co) * M.inv_cell), 0);
          s.rc1[tid] = min((int)floor((xh + P.box_r - co) * M.inv_cell), cn - 1);
        }
        __syncthreads();
        const int ncx = s.rc1[0] - s.rc0[0] + 1, ncy = s.rc1[1] - s.rc0[1] + 1, ncz1 = s.rc1[2] - s.rc0[2] + 2;
        const int ncols = ncx * ncy;
        bool staged = (ncx > 0 && ncy > 0 && ncz1 > 1) && ncols <= COL_CAP && ncols * ncz1 <= CS_CAP;
        if (staged) {
          for (int e = tid; e < ncols * ncz1; e += KT) {
            const int col = e / ncz1, lz = e % ncz1;
            const int gx = s.rc0[0] + col / ncy, gy = s.rc0[1] + col % ncy, gz = s.rc0[2] + lz;
            s.a.cstart[e] = __ldg(M.cell_start + ((gx * M.cny + gy) * M.cnz + gz));
          }
          __syncthreads();
          if (warp == 0) {  // exclusive scan of the column sizes -> shared-memory offset of each column
            int run = 0;
            for (int cb = 0; cb < ncols; cb += 32) {
              const int col = cb + lane;
              const int cnt = col < ncols ? s.a.cstart[col * ncz1 + ncz1 - 1] - s.a.cstart[col * ncz1] : 0;
              int incl = cnt;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
              if (col < ncols) s.a.col_delta[col] = run + incl - cnt - s.a.cstart[col * ncz1];
              run += __shfl_sync(FULL, incl, 31);
            }
            if (lane == 0) s.npts = run;
          }
          __syncthreads();
          staged = s.npts <= PTS_CAP;
          if (staged) {
            for (int col = warp; col < ncols; col += KT / 32) {
              const int g0 = s.a.cstart[col * ncz1], g1 = s.a.cstart[col * ncz1 + ncz1 - 1], d = s.a.col_delta[col];
              for (int g = g0 + lane; g < g1; g += 32) s.a.pts[g + d] = __ldg(M.pts + g);
            }
          }
          __syncthreads();
        }
        PH_MARK(8);  // staging time
        if (prof && tid == 0) { if (staged) s.ph[9] += 1; s.ph[11] += (unsigned long long)(staged ? s.npts : 0); s.ph[12] += (unsigned long long)nT; }
        unsigned my_hits = 0;
        (void)my_hits;
        if (staged) {
          // Unit-major sweep over the staged cell list.  Each thread keeps ONE unit's nine centres in registers and walks
          // the staged points of the cell columns its bounding box (centre +- box_r, the cell walk of the reference-
          // equivalent fallback below) touches, restricted to the z cells its centres can reach; neighbouring lanes hold
          // neighbouring lattice columns of the same checkpoint, so a warp walks (almost) the same columns and the loads
          // are shared-memory broadcasts.  With few units, G lanes share a unit and split each column's points.
          const float cullf = P.cullf;
          const int rcx = s.rc0[0], rcy = s.rc0[1], rcz = s.rc0[2], rcx1 = s.rc1[0], rcy1 = s.rc1[1], rcz1 = s.rc1[2];
          for (int base = 0, slots = 0; base < nT; base += slots) {
            // G is chosen per round from the units that are left (a short last round spreads over the whole CTA): the
            // largest split that fits them in one round, or twice that when the units it leaves over then fit a round
            // of four times the split (1/(2G) + 1/(4G) < 1/G of a single round)
            constexpr int GMAX = KT / 32;
            const int rem = nT - base;
            int G = 1;
            while (G < GMAX && rem * (G * 2) <= KT) G *= 2;
            if (4 * G <= GMAX) {
              const int rem2 = rem - KT / (2 * G);
              if (rem2 > 0 && rem2 * (4 * G) <= KT) G *= 2;
            }
            slots = KT / G;
            const int e = base + tid / G, r = tid % G;
            if (e >= nT) continue;
            const uint32_t un = s.units[e];
            const int u = (int)(un & 2047u);
            const int i = u / (na * na), ab = u % (na * na), a = ab / na, b = ab % na;
            uint32_t mask = 0;
            for (int c = 0; c < na; c++)   // drop primitives another checkpoint has rejected meanwhile
              if (((un >> (11 + c)) & 1u) && s.state[ab * na + c] == ST_FEASIBLE) mask |= 1u << c;
            const double px = s.X[i][0][a], py = s.X[i][1][b];
            const float fx = (float)px, fy = (float)py;
            float fz[UAVMP_MAXNA];
#pragma unroll
            for (int c = 0; c < UAVMP_MAXNA; c++) fz[c] = (c < na) ? (float)s.X[i][2][c] : 1e30f;
            // two points per trip: their cull chains are independent, which roughly halves the dependent latency
            // Float slab test without a body-axis table.  The ellipsoid is diag(r, r, h) in the body frame, so with w = d . b3:
            // || E^-1 d ||^2 = |d|^2 / r^2 + w^2 (1/h^2 - 1/r^2), and b3 = n / |n| with n = (ux, uy, uz + 9.81) (:724-726), hence
            //   || E^-1 d ||^2 <= v   <=>   (d . n)^2 (1/h^2 - 1/r^2) <= (v - |d|^2 / r^2) |n|^2
            // where d . n and |n|^2 are sums of per-axis terms (registers for the unit's a and b, a 9-entry shared-memory table
            // for c): no global table load inside the per-candidate loop.  v = 1 + slab_margin: above it the pair is a miss;
            // v = 1 - slab_margin: below it a hit — both final, the float error is far below the margin (1 %, widened by the
            // host for maps whose coordinates are coarse in float).  Only the shell in between takes the f64 path.
            const float nxf = (float)P.ua[a], nyf = (float)P.ua[b];
            const float nab2 = nxf * nxf + nyf * nyf;
            const float kq = P.inv_h2f - P.inv_r2f, v_hi = 1.f + slab_margin, v_lo = 1.f - slab_margin;
            auto decide = [&](const float4& q, float dx, float dy, float dxy2, uint32_t cand) {
              const float dn_ab = dx * nxf + dy * nyf;
              while (cand) {
                const int c = __ffs(cand) - 1;
                cand &= cand - 1;
                if (!((mask >> c) & 1u)) continue;
                const float dz = q.z - fz[c];
                const float t = dn_ab + dz * s.nzf[c];
                const float lhs = t * t * kq, rr = (dxy2 + dz * dz) * P.inv_r2f, n2 = nab2 + s.nz2[c];
                if (lhs > (v_hi - rr) * n2) continue;                                                               // clear miss
                if (lhs < (v_lo - rr) * n2) { s.state[ab * na + c] = ST_REJECT; mask &= ~(1u << c); continue; }     // clear hit
                EllipsoidTest t64;
                make_test(t64, P, lat, ab * na + c, px, py, s.X[i][2][c]);
                if (point_hits(t64, q)) { s.state[ab * na + c] = ST_REJECT; mask &= ~(1u << c); }
              }
            };
            if (!mask) continue;
            int trip = 0;
            constexpr int NP = 4;  // points per trip
            double zl = 1e300, zh = -1e300;
            for (int c = 0; c < na; c++)
              if ((mask >> c) & 1u) { zl = fmin(zl, s.X[i][2][c]); zh = fmax(zh, s.X[i][2][c]); }
            const int bx0 = max((int)floor((px - P.box_r - M.cox) * M.inv_cell), rcx), bx1 = min((int)floor((px + P.box_r - M.cox) * M.inv_cell), rcx1);
            const int by0 = max((int)floor((py - P.box_r - M.coy) * M.inv_cell), rcy), by1 = min((int)floor((py + P.box_r - M.coy) * M.inv_cell), rcy1);
            const int bz0 = max((int)floor((zl - P.box_r - M.coz) * M.inv_cell), rcz), bz1 = min((int)floor((zh + P.box_r - M.coz) * M.inv_cell), rcz1);
            if (bz0 > bz1) continue;
            for (int cx = bx0; cx <= bx1 && mask; cx++)
            for (int cy = by0; cy <= by1 && mask; cy++) {
            const int lcol = (cx - rcx) * ncy + (cy - rcy);
            const int cdel = s.a.col_delta[lcol];
            const int k1 = s.a.cstart[lcol * ncz1 + (bz1 + 1 - rcz)] + cdel;
            for (int g = s.a.cstart[lcol * ncz1 + (bz0 - rcz)] + cdel + r; g < k1 && mask; g += NP * G, trip++) {
              if ((trip & 15) == 15) {  // somebody else may have rejected these primitives in the meantime
                for (int c = 0; c < na; c++) if (s.state[ab * na + c] != ST_FEASIBLE) mask &= ~(1u << c);
              }
              float4 q[NP];
              float dx[NP], dy[NP], d2[NP];
              bool in[NP];
              bool any_in = false;
#pragma unroll
              for (int k = 0; k < NP; k++) {
                const bool ok = g + k * G < k1;
                q[k] = s.a.pts[ok ? g + k * G : g];
                dx[k] = q[k].x - fx; dy[k] = q[k].y - fy;
                d2[k] = dx[k] * dx[k] + dy[k] * dy[k];
                in[k] = ok && d2[k] <= cullf;
                any_in = any_in || in[k];
              }
              if (!any_in) continue;  // outside the cylinder every centre of the unit lives in
              uint32_t cm[NP];
#pragma unroll
              for (int k = 0; k < NP; k++) cm[k] = 0;
#pragma unroll
              for (int c = 0; c < UAVMP_MAXNA; c++) {
#pragma unroll
                for (int k = 0; k < NP; k++) {
                  const float z = q[k].z - fz[c];
                  if (d2[k] + z * z <= cullf) cm[k] |= 1u << c;
                }
              }
#pragma unroll
              for (int k = 0; k < NP; k++)
                if (in[k]) { my_cloud += (unsigned)__popc(cm[k] & mask); decide(q[k], dx[k], dy[k], d2[k], cm[k] & mask); }
            }
            }
          }
        } else {
          for (int e = tid; e < nT; e += KT) {
            const uint32_t un = s.units[e];
            const int u = (int)(un & 2047u);
            const int i = u / (na * na), ab = u % (na * na), a = ab / na, b = ab % na;
            for (int c = 0; c < na; c++) {
              if (!((un >> (11 + c)) & 1u) || s.state[ab * na + c] != ST_FEASIBLE) continue;
              EllipsoidTest t;
              make_test(t, P, lat, ab * na + c, s.X[i][0][a], s.X[i][1][b], s.X[i][2][c]);
              if (ellipsoid_hit_global(M, P, t, my_cloud)) { s.state[ab * na + c] = ST_REJECT; my_hits++; }
            }
          }
        }
        __syncthreads();
      }
      // survivors: end state -> voxel key (:216, posToIndex :302-310); the dedup table lives where the tile was
      for (int i = tid; i < TAB_SIZE; i += KT) s.b.tab[i] = 0;
      for (int pb = 0; pb < nprim; pb += KT) {  // one shared-memory atomic per warp, not per survivor
        const int p = pb + tid;
        const bool live = p < nprim && s.state[p] == ST_FEASIBLE;
        if (live) {
          const int a = p / (na * na), b = (p / na) % na, c = p % na;
          bool kok;
          s.key[p] = pack_key(s.EI[0][a], s.EI[1][b], s.EI[2][c], kok);
          if (!kok) atomicOr(bt.error_flag, 1);
        }
        const unsigned m = __ballot_sync(FULL, live);
        int at = 0;
        if (lane == 0 && m) at = atomicAdd(&s.n1, __popc(m));
        at = __shfl_sync(FULL, at, 0);
        if (live) s.list1[at + __popc(m & ((1u << lane) - 1u))] = (uint16_t)p;
      }
      __syncthreads();
      PH_MARK(3);
      const int n1 = s.n1;

      // ---- B1. group identical voxel keys inside this expansion --------------------------------------
      for (int e = tid; e < n1; e += KT) {
        const int p = s.list1[e];
        const unsigned long long k = s.key[p];
        uint32_t h = hash_key(k, 11);
        for (;;) {
          uint32_t cur = s.b.tab[h];
          if (cur == 0) {
            cur = atomicCAS(&s.b.tab[h], 0u, (uint32_t)p + 1u);
            if (cur == 0) break;
          }
          if (s.key[cur - 1] == k) { atomicMin(&s.b.tab[h], (uint32_t)p + 1u); break; }
          h = (h + 1) & (TAB_SIZE - 1);
        }
        s.id[p] = h;
      }
      __syncthreads();
      // ---- B2. leaders probe the global table (:220-225): one 32 B sector per probe --------------------
      for (int e = tid; e < n1; e += KT) {
        const int p = s.list1[e];
        const int leader = (int)s.b.tab[s.id[p]] - 1;
        if (leader != p) { s.id[p] = (uint32_t)leader; s.state[p] = ST_FOLLOW_NOCAND; continue; }
        const unsigned long long k = s.key[p];
        const double gp = s.cg + ginc_of(P, p);
        uint32_t h = hash_key(k, table_bits);
        uint8_t st;
        s.b.winm[p] = -1;
        s.b.inun[p] = 0;
        for (;;) {
          const uint4 w0 = __ldcg(reinterpret_cast<const uint4*>(&table[h]));      // key | id | heap_pos
          const uint4 w1 = __ldcg(reinterpret_cast<const uint4*>(&table[h]) + 1);  // g | closed | pad (same sector: no
                                                                                   // second round trip on a key match)
          const unsigned long long hk = ((unsigned long long)w0.y << 32) | w0.x;
          if ((uint32_t)(hk & EPOCH_MASK) != epoch) { st = ST_NEW; s.b.gcur[p] = gp; break; }
          if ((hk >> EPOCH_BITS) == k) {
            if (w1.z) {
              st = ST_CLOSED;
            } else {
              const double gold = __longlong_as_double(((long long)w1.y << 32) | (long long)w1.x);
              s.b.gcur[p] = gold;
              s.id[p] = w0.z;
              s.b.hs[p] = h;
              s.b.hpos[p] = w0.w;
              st = (gp < gold) ? ST_OPEN_CAND : ST_OPEN_NOCAND;
            }
            break;
          }
          h = (h + 1) & tmask;
        }
        s.state[p] = st;
      }
      __syncthreads();
      // ---- C. ordered id assignment for new nodes (== use_node_num_++ in lattice order) ----------------
      {
        const int p0 = tid * 3;
        int c = 0;  // low 16 bits: new nodes, high 16 bits: all commit candidates (new + improving), both in lattice order
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const int p = p0 + j;
          if (p < nprim) {
            uint8_t st = s.state[p];
            if (st == ST_FOLLOW_NOCAND) {  // a follower takes its verdict from its group leader (final since the probe)
              const int leader = (int)s.id[p];
              if (s.state[leader] == ST_CLOSED) st = ST_CLOSED;
              else if (s.cg + ginc_of(P, p) < s.b.gcur[leader]) st = ST_FOLLOW_CAND;
              s.state[p] = st;
            }
            if (st == ST_NEW) c += 0x10001;
            else if (st == ST_OPEN_CAND || st == ST_FOLLOW_CAND) c += 0x10000;
          }
        }
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int v = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += v;
        }
        if (lane == 31) s.wsum[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; w++) woff += s.wsum[w];
        int excl = woff + incl - c;
        const int base = s.use_num;
        int en = excl & 0xffff, ec = excl >> 16;
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const int p = p0 + j;
          if (p < nprim) {
            const uint8_t st = s.state[p];
            if (st == ST_NEW) { s.id[p] = (uint32_t)(base + en++); s.list2[ec++] = (uint16_t)p; }
            else if (st == ST_OPEN_CAND || st == ST_FOLLOW_CAND) s.list2[ec++] = (uint16_t)p;
          }
        }
        if (tid == KT - 1) { s.n_new = (woff + incl) & 0xffff; s.n2 = (woff + incl) >> 16; }
      }
      __syncthreads();
      PH_MARK(4);
      const int n_new = s.n_new;
      const int n2 = s.n2;
      if (s.use_num + n_new >= P.allocated) {
        // pool exhausted while committing this expansion (:243-247): the reference returns on the spot
        if (tid == 0) {
          // The reference stops at the primitive whose new node takes the last pool slot; its counters stop there too
          // (probes, inserts, in-place updates and heuristic calls of the lattice prefix up to that primitive).  Once per
          // query at most, so one thread replays the prefix.
          const uint32_t stop_id = (uint32_t)(P.allocated - 1);
          int pstar = nprim;
          for (int e = 0; e < n2; e++) {
            const int p = s.list2[e];
            if (s.state[p] == ST_NEW && s.id[p] == stop_id) { pstar = p; break; }
          }
          int probes = 0, upd = 0;
          for (int e = 0; e < n1; e++) probes += ((int)s.list1[e] <= pstar) ? 1 : 0;
          for (int e = 0; e < n2; e++) {
            const int p = s.list2[e];
            if (p >= pstar) break;
            const uint8_t st = s.state[p];
            if (st != ST_OPEN_CAND && st != ST_FOLLOW_CAND) continue;
            const int leader = (st == ST_FOLLOW_CAND) ? (int)s.id[p] : p;
            const double g = s.cg + ginc_of(P, p);
            if (g < s.b.gcur[leader]) { s.b.gcur[leader] = g; upd++; }  // tmp_g_cost < old_node->g_cost (:254)
          }
          const int ins = P.allocated - s.use_num;
          s.cnt[3] += probes; s.cnt[4] += ins; s.cnt[5] += upd; s.cnt[6] += ins + upd;
          s.use_num = P.allocated;
          s.status = UAVMP_NO_PATH_FOUND;
        }
        __syncthreads();
        break;
      }
      // The ordered replay (phase D) will stage the ancestors of the new leaves l0+1 .. l0+n_new: pull their sectors towards L2 now,
      // under the heuristic.  The arenas' open lists (1.6 MB each, 296 of them) do not stay in L2 between expansions.
      if (n_new > 0) {
        const int l0 = s.heap_len, m = min(n_new, PUSH_BATCH);
        for (int t = tid; t < m; t += KT) {  // levels 1 .. 5 of every leaf (32 B sectors: two slots each)
          const int leaf = l0 + 1 + t;
#pragma unroll
          for (int d = 1; d <= 5; d++) if ((leaf >> d) >= HTOP) asm volatile("prefetch.global.L2 [%0];" ::"l"(H + (leaf >> d)));
        }
        if (tid < 32 && ((l0 + 1) >> (6 + tid)) >= HTOP) {  // the levels above are shared by (almost) all leaves of the batch
          asm volatile("prefetch.global.L2 [%0];" ::"l"(H + ((l0 + 1) >> (6 + tid))));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(H + ((l0 + m) >> (6 + tid))));
        }
      }
      // ---- B3. heuristic for every candidate (:232,:259) ----------------------------------------------
      for (int e = tid; e < n2; e += KT) {
        const int p = s.list2[e];
        const int a = p / (na * na), b = (p / na) % na, c = p % na;
        double topt;
        const double h = d_heuristic(P, s.EX[0][a], s.EX[1][b], s.EX[2][c], s.EV[0][a], s.EV[1][b], s.EV[2][c], s.gp[0],
                                     s.gp[1], s.gp[2], s.gv[0], s.gv[1], s.gv[2], topt);
        const double gp = s.cg + ginc_of(P, p);
        s.b.f[p] = gp + P.lambda * h;
        s.b.gpc[p] = gp;
        s.b.topt[p] = topt;
      }
      // ---- C2. node records + hash slots of the new nodes, in parallel -----------------------------------
      for (int e = tid; e < n2; e += KT) {
        const int p = s.list2[e];
        if (s.state[p] != ST_NEW) continue;
        const int a = p / (na * na), b = (p / na) % na, c = p % na;
        const uint32_t nid = s.id[p];
        const unsigned long long k = s.key[p];
        const unsigned long long want = (k << EPOCH_BITS) | epoch;
        uint32_t h = hash_key(k, table_bits);
        for (;;) {
          unsigned long long cur = table[h].key;
          if ((uint32_t)(cur & EPOCH_MASK) != epoch) {
            unsigned long long old = atomicCAS(&table[h].key, cur, want);
            if (old == cur) break;
            continue;  // somebody else claimed it: re-read the same slot
          }
          h = (h + 1) & tmask;
        }
        const double g = s.cg + ginc_of(P, p);
        table[h].id = nid; table[h].g = g; table[h].closed = 0;
        s.b.hs[p] = h;
        KinoNode nd;
        nd.px = s.EX[0][a]; nd.py = s.EX[1][b]; nd.pz = s.EX[2][c];
        nd.vx = s.EV[0][a]; nd.vy = s.EV[1][b]; nd.vz = s.EV[2][c];
        nd.g = g;
        nd.parent = s.cur_id; nd.hslot = h; nd.input = (uint16_t)p; nd.closed = 0; nd.pad0 = 0; nd.pad1 = 0;
        nodes[nid] = nd;
      }
      __syncthreads();
      // ---- U. in-place mutations (:249-266), resolved in parallel.  Within one voxel group the reference applies the
      // candidates in lattice order, each one iff its g is below the running minimum; whether candidate e improves
      // depends only on the group's earlier candidates, the group's final state is its LAST improving candidate, and
      // `optimal_time` (Q2) is left by the last improving candidate / new node overall.  Only groups whose node is an
      // ancestor of a new leaf (or a new leaf itself) must still touch the open list in order, in phase D. -------------
      int my_ev = -1;
      for (int e = tid; e < n2; e += KT) {
        const int p = s.list2[e];
        const uint8_t st = s.state[p];
        bool event = false;  // does this candidate change optimal_time / count as an update?
        HeapSlot rec;        // what the ordered commit (phase D) replays for this candidate
        rec.f = s.b.f[p]; rec.id = UAVMP_NONE; rec.hs = 0;
        if (st == ST_NEW) {
          event = true;
          rec.id = s.id[p]; rec.hs = s.b.hs[p];
        } else {
          const int leader = (st == ST_FOLLOW_CAND) ? (int)s.id[p] : p;
          double run = s.b.gcur[leader];  // g of the existing node (or of the new leader) before this expansion's updates
          for (int e2 = 0; e2 < e; e2++) {
            const int p2 = s.list2[e2];
            const uint8_t st2 = s.state[p2];
            if (st2 == ST_NEW) continue;
            const int l2 = (st2 == ST_FOLLOW_CAND) ? (int)s.id[p2] : p2;
            if (l2 == leader) run = fmin(run, s.b.gpc[p2]);
          }
          if (s.b.gpc[p] < run) {  // tmp_g_cost < old_node->g_cost (:254)
            event = true;
            atomicMax(&s.b.winm[leader], e);
            atomicAdd(&s.n_upd, 1);
            bool in_union = s.state[leader] == ST_NEW;
            if (!in_union && n_new > 0) {
              const int pos1 = (int)s.b.hpos[leader] + 1, l0 = s.heap_len;
              for (int d = 0; d < 32; d++) {
                const int lo = (l0 + 1) >> d, hi = (l0 + n_new) >> d;
                if (lo < 1) break;
                if (pos1 >= lo && pos1 <= hi) { in_union = true; break; }
              }
            }
            if (in_union) { s.b.inun[leader] = 1; rec.id = s.id[leader]; rec.hs = s.b.hs[leader] | EV_SETKEY; }
            s.b.imp[p] = in_union ? 2 : 1;
          } else {
            s.b.imp[p] = 0;
          }
        }
        s.b.ev[e] = rec;
        if (event && s.b.topt[p] >= 0.0) my_ev = e;  // e grows with the trip count: the last one is this thread's maximum
      }
      my_ev = __reduce_max_sync(FULL, my_ev);
      if (lane == 0 && my_ev >= 0) atomicMax(&s.last_ev, my_ev);
      __syncthreads();
      PH_MARK(5);

      // ---- D. ordered part of the commit (:225-266): warp 0 replays, in lattice order, the heap pushes of the new nodes
      // (through the staged ancestor closure: one push = one ballot) and the key mutations of nodes that are ancestors
      // of new leaves.  Everything else was resolved in U and is written back in D2. -------------------------------------
      if (warp == 0) {
        const long long t_d0 = prof ? clock64() : 0;
        int pushed = 0, len = s.heap_len, batch_left = 0;
        Closure cl;
        cl.lo = cl.off = cl.lo_m1 = cl.off_m1 = cl.total = 0; cl.active = false;
        ChainCache cc;
        cc.e.f = 0.0; cc.e.id = 0; cc.e.hs = 0; cc.ix = 0;
        const unsigned lm = (2u << lane) - 1u;
        int nev = 0;  // the events, compacted (order kept)
        for (int cb = 0; cb < n2; cb += 32) {
          const bool in = cb + lane < n2;
          HeapSlot r;
          r.f = 0.0; r.id = UAVMP_NONE; r.hs = 0;
          if (in) r = s.b.ev[cb + lane];
          const bool isev = r.id != UAVMP_NONE;
          const unsigned m = __ballot_sync(FULL, isev);
          if (isev) s.b.evc[nev + __popc(m & (lm >> 1))] = r;
          nev += __popc(m);
        }
        __syncwarp();
        int k = 0;
        while (k < nev) {
          HeapSlot ev = s.b.evc[k];
          if (ev.hs & EV_SETKEY) {
            const long long t0 = prof ? clock64() : 0;
            if (cl.active) chain_writeback(s, cc);
            heap_set_key_slow(s, H, table, ev.id, ev.hs & HS_MASK, ev.f, lane, cl);
            if (cl.active && batch_left) chain_fill(s, cl, cc, len + 1, lane);
            if (prof && lane == 0) s.ph[14] += (unsigned long long)(clock64() - t0);
            k++;
            continue;
          }
          if (batch_left == 0) {
            const long long t0 = prof ? clock64() : 0;
            if (cl.active) chain_writeback(s, cc);
            closure_flush(s, H, table, lane, cl);
            batch_left = min(min(PUSH_BATCH, n_new - pushed), level_room(len));
            closure_load(s, H, len, batch_left, lane, cl);
            chain_fill(s, cl, cc, len + 1, lane);
            if (prof && lane == 0) s.ph[13] += (unsigned long long)(clock64() - t0);
          }
          for (;;) {  // a run of pushes inside one staged batch; the next record is fetched under the current push
            k++;
            HeapSlot nx;
            nx.f = 0.0; nx.id = 0; nx.hs = EV_SETKEY;
            if (k < nev) nx = s.b.evc[k];
            chain_push(cc, ev.f, ev.id, ev.hs, lm);
            len++; pushed++; batch_left--;
            if (batch_left == 0 || (nx.hs & EV_SETKEY)) break;
            chain_advance(s, cc, len + 1, lane);
            ev = nx;
          }
        }
        {
          const long long t0 = prof ? clock64() : 0;
          if (cl.active) chain_writeback(s, cc);
          closure_flush(s, H, table, lane, cl);
          if (prof && lane == 0) s.ph[13] += (unsigned long long)(clock64() - t0);
        }
        if (prof && lane == 0) s.ph[10] += (unsigned long long)(clock64() - t_d0);  // diagnostics: the whole ordered replay
        if (lane == 0) {
          if (len > 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + s.htop[1].id));  // the next pop's node record
          s.heap_len = len;
          s.use_num += n_new;
          if (s.last_ev >= 0) s.opt_time = s.b.topt[s.list2[s.last_ev]];
          s.cnt[3] += n1; s.cnt[4] += n_new; s.cnt[5] += s.n_upd; s.cnt[6] += n_new + s.n_upd;
        }
      }
      else {
        // ---- D2 (warps 1 .. 7, concurrently with the ordered replay on warp 0): the recorded mutations that cannot move
        // during the expansion — node state, g in the hash slot, and the cached heap key of nodes that are NOT ancestors of a
        // new leaf (the replay never reads or writes their slots) ------------------------------------------------------------
        const long long t_d2 = (prof && tid == 32) ? clock64() : 0;
        for (int e = tid - 32; e < n1; e += KT - 32) {
          const int p = s.list1[e];
          const uint8_t st = s.state[p];
          if (st != ST_NEW && st != ST_OPEN_CAND && st != ST_OPEN_NOCAND) continue;  // leaders only
          const int we = s.b.winm[p];
          if (we < 0) continue;
          const int w = s.list2[we];
          const int a = w / (na * na), b = (w / na) % na, c = w % na;
          KinoNode* nd = nodes + s.id[p];
          nd->px = s.EX[0][a]; nd->py = s.EX[1][b]; nd->pz = s.EX[2][c];
          nd->vx = s.EV[0][a]; nd->vy = s.EV[1][b]; nd->vz = s.EV[2][c];
          const double g = s.cg + ginc_of(P, w);
          nd->g = g; nd->parent = s.cur_id; nd->input = (uint16_t)w;
          table[s.b.hs[p]].g = g;
          if (st != ST_NEW && !s.b.inun[p]) hstore_key(s, H, (int)s.b.hpos[p] + 1, s.b.f[w]);
        }
        if (prof && tid == 32) s.ph[15] += (unsigned long long)(clock64() - t_d2);
      }
      __syncthreads();
      PH_MARK(6);
    }  // main loop

    // ---- query epilogue ------------------------------------------------------------------------------
    atomicAdd(&s.cnt[1], (unsigned long long)my_occ);
    atomicAdd(&s.cnt[2], (unsigned long long)my_cloud);
    __syncthreads();
    if (tid == 0) {
      bt.status[q] = s.status;
      bt.use_node_num[q] = s.use_num;
      bt.n_pop[q] = s.n_pop;
      bt.pop_hash[q] = s.pop_hash;
      if (s.status != UAVMP_REACH_END) bt.n_path[q] = 0;
      s.cnt[0] = (unsigned long long)s.n_pop;
      if (prof && bt.query_cycles) bt.query_cycles[q] = clock64() - q_t0;
      if (prof && bt.query_phase) { for (int k = 0; k < 16; k++) { bt.query_phase[(size_t)q * 16 + k] = s.ph[k] - s.phq[k]; s.phq[k] = s.ph[k]; } }
    }
    __syncthreads();
    if (tid < 8) atomicAdd(&bt.counters[tid], s.cnt[tid]);
    if (qp.enabled) {
      // a truncated path (error bit 4) never reaches the QP: the batch call reports ECAP and the query is left unsolved
      const bool go = (s.status == UAVMP_REACH_END) && bt.n_path[q] >= 1 && !s.trunc;
      if (!go) {
        double* out = qp.coef + (size_t)q * 3 * qp.n;
        for (int j = tid; j < 3 * qp.n; j += KT) out[j] = 0.0;
      }
      if (tid == 0) {
        qp.qp_solved[q] = go ? 1 : 0;
        if (go) { s.pend[s.npend] = 3 * q; s.pend[s.npend + 1] = 3 * q + 1; s.pend[s.npend + 2] = 3 * q + 2; s.npend += 3; }
      }
      __syncthreads();
      while (s.npend >= qp.warps) qp_round(s, smem_raw, pl, qp, qps, bt, tid);
    }
  }
  __syncthreads();
  if (tid == 0) { __threadfence(); atomicExch(arena_busy + s.arena, 0); }  // release: the next owner sees this CTA's arena writes
  PH_MARK(7);
  __syncthreads();
  if (prof && tid < 16) atomicAdd(&bt.phase_cycles[tid], s.ph[tid]);
}

// ---- map preprocessing -------------------------------------------------------------------------------
__global__ void k_flags_base(const int8_t* occ, uint8_t* flags, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { int8_t o = occ[i]; flags[i] = (uint8_t)((o == 1 ? 1 : 0) | (o != 0 ? 2 : 0)); }
}
__global__ void k_mark_points(const float* cloud, int n, MapDev M, int margin, uint8_t* mark) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = (double)cloud[3 * i], y = (double)cloud[3 * i + 1], z = (double)cloud[3 * i + 2];
  int ix = (int)floor((x - M.ox) * M.inv_res), iy = (int)floor((y - M.oy) * M.inv_res),
      iz = (int)floor((z - M.oz) * M.inv_res);
  if (ix < -margin || iy < -margin || iz < -margin || ix >= M.nx + margin || iy >= M.ny + margin ||
      iz >= M.nz + margin)
    return;
  ix = min(max(ix, 0), M.nx - 1); iy = min(max(iy, 0), M.ny - 1); iz = min(max(iz, 0), M.nz - 1);
  mark[((size_t)ix * M.ny + iy) * M.nz + iz] = 1;
}
// near[j] = 1 for every voxel j a checkpoint could sit in while a cloud point of marked voxel i is within `reach` of it:
// a point in voxel i and a position in voxel j are farther apart than sum_a (max(|i_a - j_a| - 1, 0) res)^2, so voxels whose
// gap vector is longer than reach (in voxels, squared: reach2) need no flag.  Roughly half the cube of the same half-width.
__global__ void k_scatter_near(const uint8_t* mark, uint8_t* near, int nx, int ny, int nz, int r, double reach2) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)nx * ny * nz;
  if (i >= n || !mark[i]) return;
  const int z = (int)(i % nz), y = (int)((i / nz) % ny), x = (int)(i / ((size_t)nz * ny));
  for (int dx = -r; dx <= r; dx++) {
    const int xx = x + dx;
    if (xx < 0 || xx >= nx) continue;
    const double gx = (double)max(abs(dx) - 1, 0);
    for (int dy = -r; dy <= r; dy++) {
      const int yy = y + dy;
      if (yy < 0 || yy >= ny) continue;
      const double gy = (double)max(abs(dy) - 1, 0);
      if (gx * gx + gy * gy > reach2) continue;
      for (int dz = -r; dz <= r; dz++) {
        const int zz = z + dz;
        if (zz < 0 || zz >= nz) continue;
        const double gz = (double)max(abs(dz) - 1, 0);
        if ((gx * gx + gy * gy) + gz * gz > reach2) continue;
        near[((size_t)xx * ny + yy) * nz + zz] = 1;  // every writer stores 1
      }
    }
  }
}
__global__ void k_or_near(uint8_t* flags, const uint8_t* near, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (uint8_t)((flags[i] & 3u) | (near[i] ? 4u : 0u));
}
__global__ void k_pad_flags(const uint8_t* in, uint8_t* out, int ny, int nz, int nzp, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int z = (int)(i % nz), y = (int)((i / nz) % ny);
  const size_t x = i / ((size_t)nz * ny);
  out[(((x * (nzp >> 4) + (z >> 4)) * ny + y) << 4) + (z & 15)] = in[i];
}
__global__ void k_cell_ids(const float* cloud, int n, MapDev M, int* cell, int* idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = (double)cloud[3 * i], y = (double)cloud[3 * i + 1], z = (double)cloud[3 * i + 2];
  int cx = (int)floor((x - M.cox) * M.inv_cell), cy = (int)floor((y - M.coy) * M.inv_cell),
      cz = (int)floor((z - M.coz) * M.inv_cell);
  int c;
  if (cx < 0 || cy < 0 || cz < 0 || cx >= M.cnx || cy >= M.cny || cz >= M.cnz) c = M.cnx * M.cny * M.cnz;  // dropped
  else c = (cx * M.cny + cy) * M.cnz + cz;
  cell[i] = c;
  idx[i] = i;
}
__global__ void k_gather_pts(const float* cloud, const int* idx_sorted, int n, float4* pts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int j = idx_sorted[i];
  pts[i] = make_float4(cloud[3 * j], cloud[3 * j + 1], cloud[3 * j + 2], 0.f);
}
__global__ void k_cell_start(const int* cell_sorted, int n, int ncell, int* cell_start) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncell) return;
  int lo = 0, hi = n;  // lower_bound(cell_sorted, c)
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cell_sorted[mid] < c) lo = mid + 1; else hi = mid;
  }
  cell_start[c] = lo;
}

__global__ void k_pack_paths(const double* stage, const int* n_path, const long long* offsets, int path_cap,
                             double* out) {
  int q = blockIdx.x;
  int n = n_path[q];
  const double* src = stage + (size_t)q * path_cap * 3;
  double* dst = out + offsets[q] * 3;
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) dst[i] = src[i];
}
__global__ void k_offsets(const int* n_path, int B, long long* offsets) {
  // single-block exclusive scan (B <= a few 100k): chunked serial-per-thread + block scan
  __shared__ long long part[1024];
  int t = threadIdx.x, T = blockDim.x;
  int per = (B + T - 1) / T;
  int b0 = t * per, b1 = min(B, b0 + per);
  long long sum = 0;
  for (int i = b0; i < b1; i++) sum += n_path[i];
  part[t] = sum;
  __syncthreads();
  if (t == 0) {
    long long acc = 0;
    for (int i = 0; i < T; i++) { long long v = part[i]; part[i] = acc; acc += v; }
    offsets[B] = acc;
  }
  __syncthreads();
  long long acc = part[t];
  for (int i = b0; i < b1; i++) { offsets[i] = acc; acc += n_path[i]; }
}
__global__ void k_dist_keys(const double* sp, const double* ep, int B, float* key, int* idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  double dx = sp[3 * i] - ep[3 * i], dy = sp[3 * i + 1] - ep[3 * i + 1], dz = sp[3 * i + 2] - ep[3 * i + 2];
  key[i] = (float)sqrt(dx * dx + dy * dy + dz * dz);
  idx[i] = i;
}

__global__ void k_fpmath(int op, int npow, const double* x, double* y, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = x[i], r = 0;
  switch (op) {
    case 0: r = fpm::cbrt(v); break;
    case 1: r = fpm::acos(v); break;
    case 2: r = fpm::cos(v); break;
    case 3: r = fpm::powi(v, npow); break;
  }
  y[i] = r;
}

}  // namespace

// =====================================================================================================
// host side
// =====================================================================================================
int uavmp_fail(uavmp_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

// Host f64 below mirrors the reference's expressions; this translation unit's host code is compiled with
// -ffp-contract=off so the tables are bit-identical to what the oracle computes.
int kino_upload_params(uavmp_ctx* ctx) {
  const uavmp_kino_params& kp = ctx->kp;
  if (kp.allocated_node_num < 2) return uavmp_fail(ctx, UAVMP_EINVAL, "allocated_node_num must be >= 2");
  if (kp.collision_check_type != 1 && kp.collision_check_type != 2)
    return uavmp_fail(ctx, UAVMP_EINVAL, "collision_check_type must be 1 or 2");
  if (!(kp.time_step_size > 0) || !(kp.sample_tau > 0) || !(kp.acc_resolution > 0) || !(kp.max_accelration > 0))
    return uavmp_fail(ctx, UAVMP_EINVAL, "time_step_size, sample_tau, acc_resolution, max_accelration must be > 0");
  KinoParamsDev P;
  memset(&P, 0, sizeof(P));
  P.allocated = kp.allocated_node_num;
  P.ctype = kp.collision_check_type;
  int segment_num = (int)std::floor(kp.sample_tau / kp.time_step_size);  // kino_astar.cpp:169
  if (segment_num + 1 > UAVMP_MAXK) return uavmp_fail(ctx, UAVMP_EINVAL, "too many checkpoints per primitive");
  P.K = segment_num + 1;
  for (int i = 0; i <= segment_num; i++) {
    double t = i * kp.time_step_size;
    P.tk[i] = t;
    P.hk[i] = 0.5 * t * t;
  }
  P.rou = kp.rou_time; P.lambda = kp.lambda_heu; P.goal_tol = kp.goal_tolerance; P.step = kp.time_step_size;
  P.vmax = kp.max_velocity; P.tau = kp.sample_tau; P.htau = 0.5 * kp.sample_tau * kp.sample_tau;
  P.tie = 1.0 + (3 / 1e4);  // kino_astar.cpp:69
  P.robot_r = kp.robot_r; P.robot_h = kp.robot_h;
  P.box_r = std::max(kp.robot_r, kp.robot_h) * 1.001 + 1e-6;
  float radius = (float)(kp.robot_r + 1e-1);
  P.kd_r2 = radius * radius;
  P.cull2 = P.box_r * P.box_r;
  P.cullf = (float)((P.box_r + 2e-3) * (P.box_r + 2e-3));
  P.inv_r2 = 1.0 / (kp.robot_r * kp.robot_r);
  P.inv_h2 = 1.0 / (kp.robot_h * kp.robot_h);
  P.inv_r2f = (float)P.inv_r2; P.inv_h2f = (float)P.inv_h2;
  if (std::max(kp.robot_r, kp.robot_h) >= kp.robot_r + 0.1)
    return uavmp_fail(ctx, UAVMP_EINVAL, "robot_h >= robot_r + 0.1: the KD-tree radius of kino_astar.cpp:744 would cut the ellipsoid");

  // acceleration lattice by float accumulation, ax outer / az inner (kino_astar.cpp:158-160)
  std::vector<double> ux, uy, uz;
  const double inv_acc_res = 1.0 / kp.acc_resolution;
  const double amax = kp.max_accelration;
  for (double ax = -amax; ax <= amax + 1e-3; ax += inv_acc_res * amax)
    for (double ay = -amax; ay <= amax + 1e-3; ay += inv_acc_res * amax)
      for (double az = -amax; az <= amax + 1e-3; az += inv_acc_res * amax) {
        ux.push_back(ax); uy.push_back(ay); uz.push_back(az);
        if (ux.size() > UAVMP_MAXPRIM) return uavmp_fail(ctx, UAVMP_EINVAL, "more than %d motion primitives", UAVMP_MAXPRIM);
      }
  const int n = (int)ux.size();
  P.nprim = n;
  // the lattice is the tensor product of one per-axis value list (same accumulation on every axis)
  int na = 0;
  while (na < n && (na == 0 || uz[na] > uz[na - 1]) && uy[na] == uy[0] && ux[na] == ux[0]) na++;
  if (na < 1 || na > UAVMP_MAXNA || na * na * na != n)
    return uavmp_fail(ctx, UAVMP_EINVAL, "acceleration lattice is not a (2r+1)^3 tensor grid with r <= %d", (UAVMP_MAXNA - 1) / 2);
  for (int p = 0; p < n; p++)
    if (ux[p] != uz[p / (na * na)] || uy[p] != uz[(p / na) % na] || uz[p] != uz[p % na])
      return uavmp_fail(ctx, UAVMP_EINVAL, "acceleration lattice is not a tensor grid");
  P.na = na;
  for (int a = 0; a < na; a++) P.ua[a] = uz[a];
  std::vector<double> host((size_t)n * 16);
  double* hux = host.data(); double* huy = hux + n; double* huz = huy + n; double* hg = huz + n; double* hE = hg + n;
  double* hb3 = hE + (size_t)9 * n;
  for (int p = 0; p < n; p++) {
    hux[p] = ux[p]; huy[p] = uy[p]; huz[p] = uz[p];
    double usq = (ux[p] * ux[p] + uy[p] * uy[p]) + uz[p] * uz[p];
    hg[p] = (usq + kp.rou_time) * kp.sample_tau;  // (ut.dot(ut) + rou_) * sample_tau_ (:231)
    // attitude from thrust direction and E^-1 (kino_astar.cpp:724-752)
    double a3[3] = {ux[p] + 9.81 * 0.0, uy[p] + 9.81 * 0.0, uz[p] + 9.81 * 1.0};
    auto nrm = [](const double* v, double* o) {
      double n2 = std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
      o[0] = v[0] / n2; o[1] = v[1] / n2; o[2] = v[2] / n2;
    };
    auto crs = [](const double* a, const double* b, double* o) {
      o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
    };
    double b3[3], b2[3], b1[3], t[3], c1[3] = {std::cos(0.0), std::sin(0.0), 0.0};
    nrm(a3, b3);
    crs(b3, c1, t); nrm(t, b2);
    crs(b2, b3, t); nrm(t, b1);
    double Rot[3][3] = {{b1[0], b2[0], b3[0]}, {b1[1], b2[1], b3[1]}, {b1[2], b2[2], b3[2]}};
    double Pd[3] = {kp.robot_r, kp.robot_r, kp.robot_h};
    double RP[3][3], E[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) RP[i][j] = Rot[i][j] * Pd[j];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) E[i][j] = (RP[i][0] * Rot[j][0] + RP[i][1] * Rot[j][1]) + RP[i][2] * Rot[j][2];
    auto cof = [&](int i, int j) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return E[i1][j1] * E[i2][j2] - E[i1][j2] * E[i2][j1];
    };
    double cc0 = cof(0, 0), cc1 = cof(1, 0), cc2 = cof(2, 0);
    double det = (cc0 * E[0][0] + cc1 * E[1][0]) + cc2 * E[2][0];
    double invdet = 1.0 / det;
    hb3[3 * p] = b3[0]; hb3[3 * p + 1] = b3[1]; hb3[3 * p + 2] = b3[2];
    double* Ei = hE + 9 * p;
    Ei[0] = cc0 * invdet; Ei[1] = cc1 * invdet; Ei[2] = cc2 * invdet;
    Ei[3] = cof(0, 1) * invdet; Ei[4] = cof(1, 1) * invdet; Ei[5] = cof(2, 1) * invdet;
    Ei[6] = cof(0, 2) * invdet; Ei[7] = cof(1, 2) * invdet; Ei[8] = cof(2, 2) * invdet;
  }
  {
    std::vector<float> hb3f((size_t)n * 4, 0.f);
    for (int p = 0; p < n; p++) for (int j = 0; j < 3; j++) hb3f[4 * p + j] = (float)hb3[3 * p + j];
    if (ctx->d_b3f) cudaFree(ctx->d_b3f);
    UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_b3f, hb3f.size() * sizeof(float)));
    UAVMP_CUDA(ctx, cudaMemcpy(ctx->d_b3f, hb3f.data(), hb3f.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (ctx->d_lattice) cudaFree(ctx->d_lattice);
  UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_lattice, host.size() * sizeof(double)));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(ctx->d_lattice, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (!ctx->d_kparams) UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_kparams, sizeof(KinoParamsDev)));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(ctx->d_kparams, &P, sizeof(P), cudaMemcpyHostToDevice, ctx->stream));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->nprim = n;
  ctx->params_dirty = false;
  ctx->flags_dirty = true;  // the "cloud nearby" dilation radius depends on robot_r / robot_h
  return UAVMP_OK;
}

static inline unsigned nblk(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

int kino_build_map(uavmp_ctx* ctx) {
  // derived device structures: flag grid (+ dilated "cloud nearby" bit) and the cell list over the cloud
  const size_t nvox = (size_t)ctx->nx * ctx->ny * ctx->nz;
  MapDev& M = ctx->map_host;
  M.nx = ctx->nx; M.ny = ctx->ny; M.nz = ctx->nz;
  M.ox = ctx->origin[0]; M.oy = ctx->origin[1]; M.oz = ctx->origin[2];
  // grid_map.cpp:72-73 and grid_map.h:372-379
  M.lox = ctx->origin[0] + 1e-4; M.loy = ctx->origin[1] + 1e-4; M.loz = ctx->origin[2] + 1e-4;
  M.hix = (ctx->origin[0] + ctx->map_size[0]) - 1e-4;
  M.hiy = (ctx->origin[1] + ctx->map_size[1]) - 1e-4;
  M.hiz = (ctx->origin[2] + ctx->map_size[2]) - 1e-4;
  M.inv_res = 1.0 / ctx->resolution;
  const double box_r = std::max(ctx->kp.robot_r, ctx->kp.robot_h) * 1.001 + 1e-6;
  const double cell = std::max(5.0 * ctx->resolution, 0.25);
  const double margin = box_r + cell;
  M.cox = ctx->origin[0] - margin; M.coy = ctx->origin[1] - margin; M.coz = ctx->origin[2] - margin;
  M.inv_cell = 1.0 / cell;
  M.cnx = (int)std::ceil((ctx->map_size[0] + 2 * margin) / cell) + 1;
  M.cny = (int)std::ceil((ctx->map_size[1] + 2 * margin) / cell) + 1;
  M.cnz = (int)std::ceil((ctx->map_size[2] + 2 * margin) / cell) + 1;
  M.n_cloud = ctx->n_cloud;
  const int ncell = M.cnx * M.cny * M.cnz;

  if (!ctx->d_flags) UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_flags, nvox));
  if (!ctx->d_tmp) UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_tmp, 2 * nvox));
  uint8_t* t0 = ctx->d_tmp; uint8_t* t1 = ctx->d_tmp + nvox;
  cudaStream_t st = ctx->stream;
  k_flags_base<<<nblk(nvox, 256), 256, 0, st>>>(ctx->d_occ, ctx->d_flags, nvox);
  UAVMP_CUDA(ctx, cudaMemsetAsync(t0, 0, nvox, st));
  // a checkpoint x in voxel i and a cloud point p in voxel j with |p - x| <= box_r on every axis satisfy
  // (|j - i| - 1) res < box_r, i.e. |j - i| <= floor(box_r / res) + 1: that is the dilation the "near a cloud point" bit needs
  const int dil = (int)std::floor(box_r * M.inv_res) + 1;
  if (ctx->n_cloud > 0) {
    k_mark_points<<<nblk(ctx->n_cloud, 256), 256, 0, st>>>(ctx->d_cloud, ctx->n_cloud, M, dil, t0);
    // ... and, inside that cube, only the voxels whose gap to the point's voxel is within box_r (Euclidean)
    UAVMP_CUDA(ctx, cudaMemsetAsync(t1, 0, nvox, st));
    const double reach = box_r * M.inv_res;
    k_scatter_near<<<nblk(nvox, 256), 256, 0, st>>>(t0, t1, M.nx, M.ny, M.nz, dil, reach * reach * (1.0 + 1e-9));
    k_or_near<<<nblk(nvox, 256), 256, 0, st>>>(ctx->d_flags, t1, nvox);
  }
  // cell list
  if (ctx->d_cell_start) cudaFree(ctx->d_cell_start);
  if (ctx->d_pts) cudaFree(ctx->d_pts);
  ctx->d_cell_start = nullptr; ctx->d_pts = nullptr;
  UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_cell_start, (size_t)(ncell + 2) * sizeof(int)));
  const int n = ctx->n_cloud;
  UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_pts, (size_t)std::max(n, 1) * sizeof(float4)));
  if (n > 0) {
    int *ck, *ci, *cks, *cis;
    UAVMP_CUDA(ctx, cudaMalloc(&ck, (size_t)4 * n * sizeof(int)));
    ci = ck + n; cks = ci + n; cis = cks + n;
    k_cell_ids<<<nblk(n, 256), 256, 0, st>>>(ctx->d_cloud, n, M, ck, ci);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, ck, cks, ci, cis, n, 0, 32, st);
    void* tmp;
    UAVMP_CUDA(ctx, cudaMalloc(&tmp, tb));
    cub::DeviceRadixSort::SortPairs(tmp, tb, ck, cks, ci, cis, n, 0, 32, st);
    k_gather_pts<<<nblk(n, 256), 256, 0, st>>>(ctx->d_cloud, cis, n, ctx->d_pts);
    k_cell_start<<<nblk(ncell + 1, 256), 256, 0, st>>>(cks, n, ncell, ctx->d_cell_start);
    UAVMP_CUDA(ctx, cudaStreamSynchronize(st));
    cudaFree(tmp);
    cudaFree(ck);
  } else {
    UAVMP_CUDA(ctx, cudaMemsetAsync(ctx->d_cell_start, 0, (size_t)(ncell + 2) * sizeof(int), st));
  }
  // z-blocked copy [nx][nzp / 16][ny][16] (nzp = nz rounded up to 16) + the tensor map over it
  M.nzp = (M.nz + 15) & ~15;
  const size_t npad = (size_t)M.nx * M.ny * M.nzp;
  if (ctx->d_flags_pad) { cudaFree(ctx->d_flags_pad); ctx->d_flags_pad = nullptr; }
  UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_flags_pad, npad + 256));
  UAVMP_CUDA(ctx, cudaMemsetAsync(ctx->d_flags_pad, 0, npad + 256, st));
  k_pad_flags<<<nblk(nvox, 256), 256, 0, st>>>(ctx->d_flags, ctx->d_flags_pad, M.ny, M.nz, M.nzp, nvox);
  M.flags = ctx->d_flags_pad;
  M.cell_start = ctx->d_cell_start;
  M.pts = ctx->d_pts;
  if (!ctx->d_map) UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_map, sizeof(MapDev)));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(ctx->d_map, &M, sizeof(M), cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(st));
  UAVMP_CUDA(ctx, cudaGetLastError());
  // cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
  ctx->have_tmap = false;
  {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
        qres == cudaDriverEntryPointSuccess) {
      static_assert(sizeof(CUtensorMap) <= sizeof(ctx->tmap_bytes), "tensor map storage");
      // the copy is [x][z / 16][y][16]: seen as 32-bit words, dimension 0 runs over (y, 4 words) so that one box row is
      // 32 y x 16 z = 512 contiguous bytes (a [x][y][z] box would be 1 024 rows of 48 B, measured ~4.5 k cycles per box)
      cuuint64_t dims[3] = {(cuuint64_t)M.ny * 4, (cuuint64_t)(M.nzp >> 4), (cuuint64_t)M.nx};
      cuuint64_t strides[2] = {(cuuint64_t)M.ny * 16, (cuuint64_t)(M.nzp >> 4) * M.ny * 16};
      cuuint32_t box[3] = {TB * 4, TBZ / 16, TB};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = ((EncodeFn)fn)((CUtensorMap*)ctx->tmap_bytes, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, ctx->d_flags_pad, dims,
                                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      ctx->have_tmap = (r == CUDA_SUCCESS);
    }
    if (!ctx->have_tmap) memset(ctx->tmap_bytes, 0, sizeof(ctx->tmap_bytes));
  }
  ctx->flags_dirty = false;
  return UAVMP_OK;
}

static int kino_ctas_per_sm() {
  int n = 0;
  cudaFuncSetAttribute(kino_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SearchSmem));
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kino_search_kernel, KT, sizeof(SearchSmem));
  return n;
}

int kino_qp_overlay_bytes() { return (int)QP_OVERLAY_BYTES; }

int kino_ensure_arenas(uavmp_ctx* ctx) {
  const int nodes = ctx->kp.allocated_node_num;
  int per_sm = kino_ctas_per_sm();
  if (per_sm < 1) return uavmp_fail(ctx, UAVMP_ECUDA, "search kernel does not fit on an SM");
  int want = ctx->sm_count * per_sm;
  int bits = 1;
  while ((1 << bits) < 2 * nodes) bits++;
  if (bits < 12) bits = 12;
  const int tsize = 1 << bits;
  if (ctx->d_arenas && ctx->n_arenas == want && ctx->arena_nodes == nodes && ctx->table_size == tsize) return UAVMP_OK;
  // the pool is shared by every batch in flight: nothing may be running while it is rebuilt
  UAVMP_CUDA(ctx, cudaDeviceSynchronize());
  if (ctx->d_arena_mem) { cudaFree(ctx->d_arena_mem); ctx->d_arena_mem = nullptr; }
  if (ctx->d_arenas) { cudaFree(ctx->d_arenas); ctx->d_arenas = nullptr; }
  if (ctx->d_arena_busy) { cudaFree(ctx->d_arena_busy); ctx->d_arena_busy = nullptr; }
  ctx->n_arenas = 0;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t sz_nodes = up((size_t)nodes * sizeof(KinoNode));
  const size_t sz_heap = up((size_t)(nodes + 4) * sizeof(HeapSlot));
  const size_t sz_tab = up((size_t)tsize * sizeof(HashSlot));
  const size_t sz_ep = 256;
  const size_t per = sz_nodes + sz_heap + sz_tab + sz_ep;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  while (want > ctx->sm_count && per * (size_t)want > free_b / 2) want -= ctx->sm_count;
  if (per * (size_t)want > free_b) return uavmp_fail(ctx, UAVMP_ENOMEM, "not enough device memory for search arenas");
  UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_arena_mem, per * (size_t)want));
  std::vector<KinoArena> ha(want);
  char* base = (char*)ctx->d_arena_mem;
  for (int i = 0; i < want; i++) {
    char* b = base + per * (size_t)i;
    ha[i].nodes = (KinoNode*)b;
    ha[i].heap = (HeapSlot*)(b + sz_nodes);
    ha[i].table = (HashSlot*)(b + sz_nodes + sz_heap);
    ha[i].epoch = (uint32_t*)(b + sz_nodes + sz_heap + sz_tab);
    // hash tables and epoch words start at zero (epoch 0 == never used)
    UAVMP_CUDA(ctx, cudaMemsetAsync(b + sz_nodes + sz_heap, 0, sz_tab + sz_ep, ctx->stream));
  }
  UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_arenas, sizeof(KinoArena) * want));
  UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_arena_busy, sizeof(int) * want));
  UAVMP_CUDA(ctx, cudaMemsetAsync(ctx->d_arena_busy, 0, sizeof(int) * want, ctx->stream));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(ctx->d_arenas, ha.data(), sizeof(KinoArena) * want, cudaMemcpyHostToDevice, ctx->stream));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->n_arenas = want; ctx->arena_nodes = nodes; ctx->table_size = tsize;
  return UAVMP_OK;
}

void kino_free_slot(PlanSlot& sl) {
  void* ptrs[] = {sl.d_q, sl.d_order, sl.d_status, sl.d_use, sl.d_npop, sl.d_hash, sl.d_npath, sl.d_path_stage, sl.d_trace,
                  sl.d_offsets, sl.d_misc, sl.d_counters, sl.d_cub_tmp, sl.d_wp, sl.d_qp_int, sl.d_qp_out, sl.d_plan_out,
                  sl.d_plan_io, sl.qp_scr.ws};
  for (void* p : ptrs) if (p) cudaFree(p);
  sl.d_q = nullptr; sl.d_order = nullptr; sl.d_status = nullptr; sl.d_use = nullptr; sl.d_npop = nullptr; sl.d_hash = nullptr;
  sl.d_npath = nullptr; sl.d_path_stage = nullptr; sl.d_trace = nullptr; sl.d_offsets = nullptr; sl.d_misc = nullptr;
  sl.d_counters = nullptr; sl.d_cub_tmp = nullptr; sl.d_wp = nullptr; sl.d_qp_int = nullptr; sl.d_qp_out = nullptr;
  sl.d_plan_out = nullptr; sl.d_plan_io = nullptr; sl.qp_scr.ws = nullptr;
  sl.cub_tmp_bytes = sl.wp_bytes = sl.qp_int_bytes = sl.qp_out_bytes = sl.plan_out_bytes = sl.plan_io_bytes = sl.qp_scr.ws_bytes = 0;
  sl.cap = 0;
}

// (re)allocate the search buffers of a slot for B queries; a failed allocation leaves the slot empty, never dangling
int kino_ensure_slot(uavmp_ctx* ctx, PlanSlot& sl, int B) {
  if (B <= sl.cap && sl.path_cap == ctx->path_cap && sl.pop_cap == ctx->pop_cap) return UAVMP_OK;
  const int cap = std::max(B, sl.cap);
  auto fr = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
  fr(sl.d_q); fr(sl.d_order); fr(sl.d_status); fr(sl.d_use); fr(sl.d_npop); fr(sl.d_hash); fr(sl.d_npath); fr(sl.d_path_stage);
  fr(sl.d_trace); fr(sl.d_offsets);
  sl.cap = 0;
  UAVMP_CUDA(ctx, cudaMalloc(&sl.d_q, (size_t)cap * 12 * sizeof(double)));
  UAVMP_CUDA(ctx, cudaMalloc(&sl.d_order, (size_t)cap * 4 * sizeof(int)));  // idx | idx_sorted | key | key_sorted
  UAVMP_CUDA(ctx, cudaMalloc(&sl.d_status, (size_t)cap * sizeof(int)));
  UAVMP_CUDA(ctx, cudaMalloc(&sl.d_use, (size_t)cap * sizeof(int)));
  UAVMP_CUDA(ctx, cudaMalloc(&sl.d_npop, (size_t)cap * sizeof(int)));
  UAVMP_CUDA(ctx, cudaMalloc(&sl.d_hash, (size_t)cap * sizeof(unsigned long long)));
  UAVMP_CUDA(ctx, cudaMalloc(&sl.d_npath, (size_t)cap * sizeof(int)));
  UAVMP_CUDA(ctx, cudaMalloc(&sl.d_path_stage, (size_t)cap * ctx->path_cap * 3 * sizeof(double)));
  UAVMP_CUDA(ctx, cudaMalloc(&sl.d_offsets, (size_t)(cap + 1) * sizeof(long long)));
  if (ctx->pop_cap > 0) UAVMP_CUDA(ctx, cudaMalloc(&sl.d_trace, (size_t)cap * ctx->pop_cap * 3 * sizeof(int)));
  if (!sl.d_misc) UAVMP_CUDA(ctx, cudaMalloc(&sl.d_misc, 64));
  if (!sl.d_counters) UAVMP_CUDA(ctx, cudaMalloc(&sl.d_counters, 8 * sizeof(unsigned long long)));
  if (!sl.h_info) UAVMP_CUDA(ctx, cudaHostAlloc((void**)&sl.h_info, sizeof(*sl.h_info), cudaHostAllocDefault));
  {  // radix-sort scratch for the processing order
    int* idx = sl.d_order; int* idx_s = idx + cap; float* key = (float*)(idx_s + cap); float* key_s = key + cap;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, key, key_s, idx, idx_s, cap, 0, 32, sl.stream);
    if (tb > sl.cub_tmp_bytes) {
      fr(sl.d_cub_tmp); sl.cub_tmp_bytes = 0;
      UAVMP_CUDA(ctx, cudaMalloc(&sl.d_cub_tmp, tb));
      sl.cub_tmp_bytes = tb;
    }
  }
  sl.cap = cap; sl.path_cap = ctx->path_cap; sl.pop_cap = ctx->pop_cap;
  return UAVMP_OK;
}

// margin of the float ellipsoid filters: 1 % (2 mm for the distance cull), or more when the map's coordinates are coarse in float
// (the filters subtract float coordinates: error ~ 2 ulp of the largest |coordinate|, amplified by 2 R / min(r, h)^2 in the
// quadratic form)
static void float_filter_margins(const uavmp_ctx* ctx, float& slab_margin, double& delta) {
  double maxc = 0.0;
  for (int ax = 0; ax < 3; ax++) maxc = std::max(maxc, std::max(std::fabs(ctx->origin[ax]), std::fabs(ctx->origin[ax] + ctx->map_size[ax])));
  delta = 2.0 * maxc * 1.1920928955078125e-7;
  const double R = std::max(ctx->kp.robot_r, ctx->kp.robot_h);
  const double mn = std::min(ctx->kp.robot_r, ctx->kp.robot_h);
  const double err = 2.0 * R * delta / (mn * mn);
  slab_margin = (float)std::max(0.01, 4.0 * err + 1e-4);
}

int kino_launch_search(uavmp_ctx* ctx, PlanSlot& sl, int B, const double* d_sp, const double* d_sv, const double* d_ep,
                       const double* d_ev, bool sort_order, bool profile, const KinoQpDev* qp_in, const QpPlanDev* plan,
                       const uavmp_osqp_settings* settings) {
  cudaStream_t st = sl.stream;
  int* order = nullptr;
  sl.launches_aux = 0;
  if (sort_order && B > 1) {
    int* idx = sl.d_order; int* idx_s = idx + B; float* key = (float*)(idx_s + B); float* key_s = key + B;
    k_dist_keys<<<nblk(B, 256), 256, 0, st>>>(d_sp, d_ep, B, key, idx);
    size_t tb = sl.cub_tmp_bytes;
    cub::DeviceRadixSort::SortPairsDescending(sl.d_cub_tmp, tb, key, key_s, idx, idx_s, B, 0, 32, st);
    order = idx_s;
    sl.launches_aux = 1;  // k_dist_keys (+ cub's radix-sort kernels, library code)
  }
  UAVMP_CUDA(ctx, cudaMemsetAsync(sl.d_misc, 0, 64, st));
  UAVMP_CUDA(ctx, cudaMemsetAsync(sl.d_counters, 0, 8 * sizeof(unsigned long long), st));
  KinoBatchDev bt;
  bt.B = B; bt.start_pt = d_sp; bt.start_vel = d_sv; bt.end_pt = d_ep; bt.end_vel = d_ev; bt.order = order;
  bt.status = sl.d_status; bt.use_node_num = sl.d_use; bt.n_pop = sl.d_npop; bt.pop_hash = sl.d_hash;
  bt.n_path = sl.d_npath; bt.path_stage = sl.d_path_stage; bt.path_cap = sl.path_cap;
  bt.pop_trace = sl.d_trace; bt.pop_cap = sl.pop_cap;
  bt.error_flag = sl.d_misc; bt.next_query = sl.d_misc + 1; bt.counters = sl.d_counters;
  bt.phase_cycles = nullptr; bt.query_cycles = nullptr; bt.query_phase = nullptr;
  if (profile) {
    if (!ctx->d_phase) UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_phase, 16 * sizeof(unsigned long long)));
    if (ctx->query_cycles_cap < B) {
      if (ctx->d_query_cycles) cudaFree(ctx->d_query_cycles);
      ctx->d_query_cycles = nullptr; ctx->query_cycles_cap = 0;
      UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_query_cycles, (size_t)B * 17 * sizeof(long long)));
      ctx->query_cycles_cap = B;
    }
    UAVMP_CUDA(ctx, cudaMemsetAsync(ctx->d_phase, 0, 16 * sizeof(unsigned long long), st));
    bt.phase_cycles = ctx->d_phase; bt.query_cycles = ctx->d_query_cycles;
    bt.query_phase = (unsigned long long*)(ctx->d_query_cycles + B);
  }
  LatticeDev lat;
  const int n = ctx->nprim;
  lat.ux = ctx->d_lattice; lat.uy = lat.ux + n; lat.uz = lat.uy + n; lat.ginc = lat.uz + n; lat.Einv = lat.ginc + n; lat.b3 = lat.Einv + (size_t)9 * n;
  lat.b3f = ctx->d_b3f;
  int bits = 0;
  while ((1 << bits) < ctx->table_size) bits++;
  const int grid = std::min(ctx->n_arenas, B);
  ctx->last_grid = grid;
  cudaFuncSetAttribute(kino_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SearchSmem));
  CUtensorMap tm;
  memcpy(&tm, ctx->tmap_bytes, sizeof(tm));
  float slab_margin = 0.01f;
  double delta = 0.0;
  float_filter_margins(ctx, slab_margin, delta);
  const double box_r = std::max(ctx->kp.robot_r, ctx->kp.robot_h) * 1.001 + 1e-6;
  const float cullf = (float)((box_r + 2e-3 + 4.0 * delta) * (box_r + 2e-3 + 4.0 * delta));
  KinoQpDev qp;
  QpPlanDev pl;
  uavmp_osqp_settings qs;
  memset(&qp, 0, sizeof(qp)); memset(&pl, 0, sizeof(pl)); memset(&qs, 0, sizeof(qs));
  if (qp_in && qp_in->enabled) { qp = *qp_in; pl = *plan; qs = *settings; }
  kino_search_kernel<<<grid, KT, sizeof(SearchSmem), st>>>(ctx->d_kparams, lat, ctx->d_map, ctx->d_arenas, ctx->d_arena_busy,
                                                          ctx->n_arenas, bt, bits, tm,
                                                          (ctx->have_tmap && !getenv("UAVMP_NO_TMA")) ? 1 : 0, slab_margin, cullf, qp, pl, qs);
  UAVMP_CUDA(ctx, cudaGetLastError());
  sl.launches_search = 1;
  return UAVMP_OK;
}

int kino_pack_paths(uavmp_ctx* ctx, PlanSlot& sl, int B) {
  cudaStream_t st = sl.stream;
  k_offsets<<<1, 1024, 0, st>>>(sl.d_npath, B, sl.d_offsets);
  long long total = 0;
  UAVMP_CUDA(ctx, cudaMemcpyAsync(&total, sl.d_offsets + B, sizeof(long long), cudaMemcpyDeviceToHost, st));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(st));
  if (total > ctx->path_packed_cap) {
    if (ctx->d_path_packed) cudaFree(ctx->d_path_packed);
    ctx->d_path_packed = nullptr; ctx->path_packed_cap = 0;
    long long cap = std::max(total, (long long)1024);
    UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_path_packed, (size_t)cap * 3 * sizeof(double)));
    ctx->path_packed_cap = cap;
  }
  if (total > 0) k_pack_paths<<<B, 128, 0, st>>>(sl.d_path_stage, sl.d_npath, sl.d_offsets, sl.path_cap, ctx->d_path_packed);
  UAVMP_CUDA(ctx, cudaGetLastError());
  ctx->last_total_path = total;
  sl.launches_aux += 2;
  return UAVMP_OK;
}

int kino_fpmath_eval(uavmp_ctx* ctx, int op, int npow, const double* x, double* y, long long n) {
  double *dx, *dy;
  UAVMP_CUDA(ctx, cudaMalloc(&dx, n * sizeof(double)));
  UAVMP_CUDA(ctx, cudaMalloc(&dy, n * sizeof(double)));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(dx, x, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  k_fpmath<<<nblk(n, 256), 256, 0, ctx->stream>>>(op, npow, dx, dy, n);
  UAVMP_CUDA(ctx, cudaMemcpyAsync(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(dx); cudaFree(dy);
  return UAVMP_OK;
}
