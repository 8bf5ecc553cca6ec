// astar_kernel.cu — batched grid A* (first half): one warp per query.
//
// Replaces path_searching::Astar::search and its callees
// (reference: src/planner/path_searching/src/a_star.cpp:48-154 search, :156-170 heuristics, :180-190 retrievePath;
//  include/path_searching/a_star.h:19-104 node / comparator / hash table; plan_env/grid_map.h:350-385 the two lookups).
//
// What has to be reproduced for identical expanded-node sets (the reference's quirks, kept bug for bug):
//   * nodes are keyed by their exact POSITION (three doubles, a_star.h:66-69), not by a voxel index: two ways of summing
//     resolution steps that land in the same voxel with different bit patterns are different nodes;
//   * the 27 neighbour offsets come from `for (x = -res; x <= res; x += res)` (a_star.cpp:91-93), the centre (0,0,0) included
//     (it is always found in the close list);
//   * an expanded neighbour with a smaller g is overwritten in place (g, parent, f — :141-146) with no re-heapify, so the pop order
//     is whatever libstdc++'s push_heap / pop_heap produce on the live keys (same emulation as the kinodynamic search, on cached
//     keys that are written through the node's heap position);
//   * the pool check (`use_node_num_ >= allocated_node_num_`, :134-138) aborts in the middle of an expansion;
//   * the goal test is per axis, |pos - end| < resolution, on POP (:77-89); an end point outside the map returns at once (:52-56).
// Every lane evaluates one neighbour (in-map, occupancy byte, hash probe, g, heuristic); lane 0 then replays the 27 verdicts in
// lattice order against the open list.  This first version keeps the heap walk on one lane (a dependent chain of L2 round trips
// per level); it is a correctness-first "next row", not a tuned kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "uavmp_internal.h"

#define AW 4  // warps (queries in flight) per CTA
#define AFULL 0xffffffffu

namespace {

struct __align__(16) ANode {
  double px, py, pz, g;
  uint32_t parent, hpos;  // node id or UAVMP_NONE; 1-based position in the open list while open
  uint32_t closed, pad;
};
struct __align__(16) AHeap { double f; uint32_t id, pad; };
struct __align__(32) ASlot { unsigned long long k0, k1, k2; uint32_t id, tag; };  // tag == the arena's epoch: valid

struct AArena { ANode* nodes; AHeap* heap; ASlot* table; uint32_t* epoch; };

struct AParams {
  int allocated, nx, ny, nz, table_bits;
  double lambda, tie, res, ox, oy, oz, inv_res, lox, loy, loz, hix, hiy, hiz;
};

struct ABatch {
  int B;
  const double* start; const double* end;
  int* status; int* use_num; int* n_pop; unsigned long long* pop_hash; int* n_path; double* path_stage; int path_cap;
  int* next_query; int* error_flag;
};

__device__ __forceinline__ unsigned long long amix(unsigned long long h, unsigned long long v) {
  h ^= v; h *= 0x100000001b3ull; h ^= h >> 29; return h;
}
__device__ __forceinline__ unsigned long long dbits(double x) { return (unsigned long long)__double_as_longlong(x == 0.0 ? 0.0 : x); }  // -0.0 == 0.0 as keys
__device__ __forceinline__ uint32_t ahash(unsigned long long a, unsigned long long b, unsigned long long c, int bits) {
  unsigned long long h = a * 0x9E3779B97F4A7C15ull;
  h = (h ^ (h >> 32)) + b * 0xC2B2AE3D27D4EB4Full;
  h = (h ^ (h >> 29)) + c * 0x165667B19E3779F9ull;
  h ^= h >> 31;
  return (uint32_t)((h * 0x9E3779B97F4A7C15ull) >> (64 - bits));
}
// getDiagonalHeu (a_star.cpp:161-170)
__device__ __forceinline__ double a_heu(const AParams& P, double x, double y, double z, double ex, double ey, double ez) {
  const double dx = fabs(x - ex), dy = fabs(y - ey), dz = fabs(z - ez);
  const double mn = fmin(fmin(dx, dy), dz);
  const double h = ((dx + dy) + dz) + (sqrt(3.0) - 3.0) * mn;
  return P.tie * h;
}
__device__ __forceinline__ bool a_in_map(const AParams& P, double x, double y, double z) {  // grid_map.h:370-385
  if (x < P.lox || y < P.loy || z < P.loz) return false;
  if (x > P.hix || y > P.hiy || z > P.hiz) return false;
  return true;
}

// std::push_heap of (f, id) at the end of the open list (1-based array H[1..len]); every move records the node's new position
__device__ void aheap_push(AHeap* H, ANode* nodes, int& len, double f, uint32_t id) {
  int hole = ++len;
  while (hole > 1) {
    const int parent = hole >> 1;
    const AHeap pe = H[parent];
    if (!(pe.f > f)) break;
    H[hole] = pe;
    nodes[pe.id].hpos = (uint32_t)hole;
    hole = parent;
  }
  AHeap e; e.f = f; e.id = id; e.pad = 0;
  H[hole] = e;
  nodes[id].hpos = (uint32_t)hole;
}
// std::pop_heap + pop_back (bits/stl_heap.h __adjust_heap: the hole walks to a leaf taking the smaller child, the right one on
// ties; a lone left child moves up; the former last element is then pushed up from the hole).
// Called by the WHOLE warp: the walk down is a chain of dependent loads, so the warp fetches three levels below the hole per round
// trip (lane j < 14 loads the (j + 2)-th entry of the hole's binary subtree: children 2h, 2h+1, grandchildren 4h .. 4h+3, great-
// grandchildren 8h .. 8h+7 — each level one contiguous, aligned block) and then every lane walks those levels from the shuffled
// values; only lane 0 stores.  Same comparisons on the same values as the one-level loop, a third of the round trips (measured: +3 %
// on an 8 192-query batch, whose time is the tail of its longest queries — profiles/r02_reading.md).
__device__ uint32_t aheap_pop(AHeap* H, ANode* nodes, int& len, int lane) {
  const uint32_t top = H[1].id;
  const int old_len = len;
  len = old_len - 1;
  if (old_len <= 1) return top;
  const int n = len;
  const AHeap value = H[old_len];
  int hole = 1;
  while (2 * hole + 1 <= n) {
    // subtree entry j + 2 (1-based, root = hole): depth d = floor(log2(j + 2)), index hole * 2^d + (j + 2 - 2^d)
    const int r1 = lane + 2, d = 31 - __clz(r1), idx = (hole << d) + (r1 - (1 << d));
    AHeap e; e.f = 0; e.id = 0; e.pad = 0;
    if (lane < 14 && idx <= n) e = H[idx];
    int rel = 1, h = hole;  // position inside the fetched subtree / absolute index of the hole; invariant: 2 h + 1 <= n
    for (int lv = 0; lv < 3; lv++) {
      const double fl = __shfl_sync(AFULL, e.f, 2 * rel - 2), fr = __shfl_sync(AFULL, e.f, 2 * rel - 1);
      const uint32_t il = __shfl_sync(AFULL, e.id, 2 * rel - 2), ir = __shfl_sync(AFULL, e.id, 2 * rel - 1);
      const bool right = !(fr > fl);
      if (lane == 0) {
        AHeap pick; pick.f = right ? fr : fl; pick.id = right ? ir : il; pick.pad = 0;
        H[h] = pick;
        nodes[pick.id].hpos = (uint32_t)h;
      }
      rel = 2 * rel + (right ? 1 : 0);
      h = 2 * h + (right ? 1 : 0);
      if (2 * h + 1 > n) break;
    }
    hole = h;
  }
  if (lane == 0) {
    if (2 * hole == n) {
      const AHeap l = H[n];
      H[hole] = l;
      nodes[l.id].hpos = (uint32_t)hole;
      hole = n;
    }
    while (hole > 1) {
      const int parent = hole >> 1;
      const AHeap pe = H[parent];
      if (!(pe.f > value.f)) break;
      H[hole] = pe;
      nodes[pe.id].hpos = (uint32_t)hole;
      hole = parent;
    }
    H[hole] = value;
    nodes[value.id].hpos = (uint32_t)hole;
  }
  return top;
}

template <int MINB>
__global__ void __launch_bounds__(AW * 32, MINB) astar_search_kernel(AParams P, const int8_t* __restrict__ occ, const AArena* arenas, ABatch bt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const AArena ar = arenas[blockIdx.x * AW + warp];
  ANode* nodes = ar.nodes;
  AHeap* H = ar.heap;
  ASlot* table = ar.table;
  const uint32_t tmask = (1u << P.table_bits) - 1u;
  // neighbour offsets exactly as the reference's loops produce them: -res, -res + res, (-res + res) + res
  const double o0 = -P.res, o1 = o0 + P.res, o2 = o1 + P.res;
  const double ddx = (lane / 9 == 0) ? o0 : ((lane / 9 == 1) ? o1 : o2);
  const double ddy = ((lane / 3) % 3 == 0) ? o0 : (((lane / 3) % 3 == 1) ? o1 : o2);
  const double ddz = (lane % 3 == 0) ? o0 : ((lane % 3 == 1) ? o1 : o2);
  const double dnorm = sqrt((ddx * ddx + ddy * ddy) + ddz * ddz);  // Eigen::Vector3d(x, y, z).norm() (:124)

  for (;;) {
    int q = 0;
    if (lane == 0) q = atomicAdd(bt.next_query, 1);
    q = __shfl_sync(AFULL, q, 0);
    if (q >= bt.B) break;
    const double sx = bt.start[3 * q], sy = bt.start[3 * q + 1], sz = bt.start[3 * q + 2];
    const double ex = bt.end[3 * q], ey = bt.end[3 * q + 1], ez = bt.end[3 * q + 2];
    uint32_t epoch = 0;
    if (lane == 0) { epoch = *ar.epoch + 1; *ar.epoch = epoch; }  // epoch 0 = never used (the table starts zeroed); 2^32 queries per arena
    epoch = __shfl_sync(AFULL, epoch, 0);
    int status = 0, use_num = 0, len = 0, n_pop = 0, n_path = 0;
    unsigned long long ph = 0xcbf29ce484222325ull;
    if (!a_in_map(P, ex, ey, ez)) {
      status = UAVMP_NO_PATH_FOUND;  // "end_pt is out of map" (:52-56)
    } else if (lane == 0) {
      ANode nd;
      nd.px = sx; nd.py = sy; nd.pz = sz; nd.g = 0.0; nd.parent = UAVMP_NONE; nd.hpos = 0; nd.closed = 0; nd.pad = 0;
      nodes[0] = nd;
      const unsigned long long k0 = dbits(sx), k1 = dbits(sy), k2 = dbits(sz);
      uint32_t h = ahash(k0, k1, k2, P.table_bits);
      ASlot sl; sl.k0 = k0; sl.k1 = k1; sl.k2 = k2; sl.id = 0; sl.tag = epoch;
      table[h] = sl;
      aheap_push(H, nodes, len, P.lambda * a_heu(P, sx, sy, sz, ex, ey, ez), 0);
      use_num = 1;
    }
    len = __shfl_sync(AFULL, len, 0);
    use_num = __shfl_sync(AFULL, use_num, 0);
    __syncwarp();

    while (!status) {
      // ---- pop (:71-75) ------------------------------------------------------------------------------------------------
      if (len == 0) { status = UAVMP_NO_PATH_FOUND; break; }  // open list empty (:150-153); len is warp-uniform
      const uint32_t cur = aheap_pop(H, nodes, len, lane);
      if (lane == 0) nodes[cur].closed = 1;
      __syncwarp();
      const ANode cn = nodes[cur];
      n_pop++;
      ph = amix(ph, dbits(cn.px)); ph = amix(ph, dbits(cn.py)); ph = amix(ph, dbits(cn.pz)); ph = amix(ph, (unsigned long long)__double_as_longlong(cn.g));
      // ---- goal test (:77-89) ------------------------------------------------------------------------------------------------
      if (fabs(cn.px - ex) < P.res && fabs(cn.py - ey) < P.res && fabs(cn.pz - ez) < P.res) {
        if (lane == 0) {  // retrievePath (:180-190): walk the parents, then reverse
          int n = 0;
          for (uint32_t c = cur; c != UAVMP_NONE; c = nodes[c].parent) n++;
          n_path = n;
          if (n > bt.path_cap) { atomicOr(bt.error_flag, 4); n_path = 0; }
          else {
            double* out = bt.path_stage + (size_t)q * bt.path_cap * 3;
            int i = n - 1;
            for (uint32_t c = cur; c != UAVMP_NONE; c = nodes[c].parent, i--) { out[3 * i] = nodes[c].px; out[3 * i + 1] = nodes[c].py; out[3 * i + 2] = nodes[c].pz; }
          }
        }
        status = UAVMP_REACH_END;
        break;
      }
      // ---- the 27 neighbours, one per lane (:91-123) -----------------------------------------------------------------------
      int verdict = 0;  // 0 skip, 1 new, 2 smaller g on an expanded node
      uint32_t nid = 0, hslot = 0;
      double nx = 0, ny = 0, nz = 0, ng = 0, nf = 0;
      unsigned long long k0 = 0, k1 = 0, k2 = 0;
      if (lane < 27) {
        nx = cn.px + ddx; ny = cn.py + ddy; nz = cn.pz + ddz;
        if (a_in_map(P, nx, ny, nz)) {
          const int ix = (int)floor((nx - P.ox) * P.inv_res), iy = (int)floor((ny - P.oy) * P.inv_res), iz = (int)floor((nz - P.oz) * P.inv_res);
          if (__ldg(occ + ((size_t)ix * P.ny + iy) * P.nz + iz) != 1) {  // getInflateOccupancy(...) == true (:106)
            k0 = dbits(nx); k1 = dbits(ny); k2 = dbits(nz);
            uint32_t h = ahash(k0, k1, k2, P.table_bits);
            ng = cn.g + dnorm;
            verdict = 1;
            for (;;) {
              const ASlot sl = table[h];
              if (sl.tag != epoch) break;  // not expanded yet
              if (sl.k0 == k0 && sl.k1 == k1 && sl.k2 == k2) {
                nid = sl.id;
                const ANode on = nodes[nid];
                verdict = on.closed ? 0 : ((ng < on.g) ? 2 : 0);  // close_list_.find (:113) / tmp_g_cost < g_cost (:141)
                if (verdict == 2) { nx = on.px; ny = on.py; nz = on.pz; }  // f uses tmp_node->position (:145), == as keys
                break;
              }
              h = (h + 1) & tmask;
            }
            hslot = h;
            if (verdict) nf = ng + P.lambda * a_heu(P, nx, ny, nz, ex, ey, ez);
          }
        }
      }
      // ---- ordered replay (lane 0), lattice order == lane order ---------------------------------------------------------
      unsigned evm = __ballot_sync(AFULL, verdict != 0);
      if (evm) {  // the pushes of this expansion append leaves len + 1 .. : pull their parents / grandparents / great-grandparents
                  // into L1 in one round trip (a push usually stops within a level or two of its leaf) instead of one miss per push
        const int leaf = len + 1 + (lane & 15);
        const int a1 = (lane < 16) ? (leaf >> 1) : (leaf >> 2), a3 = leaf >> 3;
        double sink = (a1 >= 1) ? H[a1].f : 0.0;
        if (lane < 16 && a3 >= 1) sink += H[a3].f;
        asm volatile("" ::"d"(sink));
      }
      while (evm && !status) {
        const int l = __ffs(evm) - 1;
        evm &= evm - 1;
        const int v = __shfl_sync(AFULL, verdict, l);
        const double f = __shfl_sync(AFULL, nf, l), g = __shfl_sync(AFULL, ng, l);
        const double px = __shfl_sync(AFULL, nx, l), py = __shfl_sync(AFULL, ny, l), pz = __shfl_sync(AFULL, nz, l);
        const uint32_t id_l = __shfl_sync(AFULL, nid, l);
        uint32_t hs = __shfl_sync(AFULL, hslot, l);
        const unsigned long long a0 = __shfl_sync(AFULL, k0, l), a1 = __shfl_sync(AFULL, k1, l), a2 = __shfl_sync(AFULL, k2, l);
        if (lane == 0) {
          if (v == 1) {
            const uint32_t id = (uint32_t)use_num;
            use_num++;
            ANode nd;
            nd.px = px; nd.py = py; nd.pz = pz; nd.g = g; nd.parent = cur; nd.hpos = 0; nd.closed = 0; nd.pad = 0;
            nodes[id] = nd;
            while (table[hs].tag == epoch) hs = (hs + 1) & tmask;  // slots taken by earlier neighbours of this expansion
            ASlot sl; sl.k0 = a0; sl.k1 = a1; sl.k2 = a2; sl.id = id; sl.tag = epoch;
            table[hs] = sl;
            aheap_push(H, nodes, len, f, id);
            if (use_num >= P.allocated) status = UAVMP_NO_PATH_FOUND;  // "allocated_node_num is too small" (:134-138)
          } else {
            ANode* on = nodes + id_l;
            on->g = g; on->parent = cur;
            H[on->hpos].f = f;  // the live key of a node that stays where it is in the open list (:143-145)
          }
        }
        status = __shfl_sync(AFULL, status, 0);
      }
      len = __shfl_sync(AFULL, len, 0);
      use_num = __shfl_sync(AFULL, use_num, 0);
      __syncwarp();
    }
    if (lane == 0) {
      bt.status[q] = status; bt.use_num[q] = use_num; bt.n_pop[q] = n_pop; bt.pop_hash[q] = ph;
      bt.n_path[q] = (status == UAVMP_REACH_END) ? n_path : 0;
    }
    __syncwarp();
  }
}

}  // namespace

// =====================================================================================================
struct AstarState {
  int ctas = 4;
  int allocated = 0, n_arenas = 0, table_bits = 0;
  void* mem = nullptr;
  AArena* d_arenas = nullptr;
  int cap = 0, path_cap = 0;
  double* d_q = nullptr;
  int *d_status = nullptr, *d_use = nullptr, *d_npop = nullptr, *d_npath = nullptr, *d_misc = nullptr;
  unsigned long long* d_hash = nullptr;
  double* d_path_stage = nullptr;
  long long* d_offsets = nullptr;
  double* d_packed = nullptr; long long packed_cap = 0, last_total = 0;
};

static void astar_free(AstarState* a) {
  void* ptrs[] = {a->mem, a->d_arenas, a->d_q, a->d_status, a->d_use, a->d_npop, a->d_npath, a->d_misc, a->d_hash, a->d_path_stage,
                  a->d_offsets, a->d_packed};
  for (void* p : ptrs) if (p) cudaFree(p);
  *a = AstarState();
}
void astar_destroy(uavmp_ctx* ctx) {
  if (ctx->astar) { astar_free(ctx->astar); delete ctx->astar; ctx->astar = nullptr; }
}

static int astar_ensure(uavmp_ctx* ctx, int B) {
  if (!ctx->astar) ctx->astar = new AstarState();
  AstarState& a = *ctx->astar;
  const int allocated = ctx->astar_allocated;
  if (a.allocated != allocated) {
    astar_free(&a);
    int bits = 12;
    while ((1 << bits) < 2 * allocated) bits++;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t sz_nodes = up((size_t)allocated * sizeof(ANode)), sz_heap = up((size_t)(allocated + 2) * sizeof(AHeap)),
                 sz_tab = up(((size_t)1 << bits) * sizeof(ASlot)), per = sz_nodes + sz_heap + sz_tab + 256;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const char* env = getenv("UAVMP_ASTAR_CTAS");  // experiment knob: resident CTAs per SM (4 or 8)
    a.ctas = (env && atoi(env) == 8) ? 8 : 4;
    int want = ctx->sm_count * a.ctas * AW;  // four CTAs of AW warps per SM
    const long long fit = (long long)((free_b / 2) / per);
    if (fit < AW) return uavmp_fail(ctx, UAVMP_ENOMEM, "not enough device memory for A* arenas of %d nodes", allocated);
    if (want > fit) want = (int)(fit / AW) * AW;
    UAVMP_CUDA(ctx, cudaMalloc(&a.mem, per * (size_t)want));
    std::vector<AArena> ha(want);
    for (int i = 0; i < want; i++) {
      char* b = (char*)a.mem + per * (size_t)i;
      ha[i].nodes = (ANode*)b; ha[i].heap = (AHeap*)(b + sz_nodes); ha[i].table = (ASlot*)(b + sz_nodes + sz_heap);
      ha[i].epoch = (uint32_t*)(b + sz_nodes + sz_heap + sz_tab);
      UAVMP_CUDA(ctx, cudaMemsetAsync(b + sz_nodes + sz_heap, 0, sz_tab + 256, ctx->stream));
    }
    UAVMP_CUDA(ctx, cudaMalloc(&a.d_arenas, sizeof(AArena) * want));
    UAVMP_CUDA(ctx, cudaMemcpyAsync(a.d_arenas, ha.data(), sizeof(AArena) * want, cudaMemcpyHostToDevice, ctx->stream));
    UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    a.allocated = allocated; a.n_arenas = want; a.table_bits = bits;
  }
  if (B > a.cap || a.path_cap != ctx->astar_path_cap) {
    void* ptrs[] = {a.d_q, a.d_status, a.d_use, a.d_npop, a.d_npath, a.d_misc, a.d_hash, a.d_path_stage, a.d_offsets};
    for (void* p : ptrs) if (p) cudaFree(p);
    a.d_q = nullptr; a.d_status = a.d_use = a.d_npop = a.d_npath = a.d_misc = nullptr; a.d_hash = nullptr; a.d_path_stage = nullptr; a.d_offsets = nullptr;
    a.cap = 0;
    const int cap = std::max(B, a.cap);
    UAVMP_CUDA(ctx, cudaMalloc(&a.d_q, (size_t)cap * 6 * sizeof(double)));
    UAVMP_CUDA(ctx, cudaMalloc(&a.d_status, (size_t)cap * sizeof(int)));
    UAVMP_CUDA(ctx, cudaMalloc(&a.d_use, (size_t)cap * sizeof(int)));
    UAVMP_CUDA(ctx, cudaMalloc(&a.d_npop, (size_t)cap * sizeof(int)));
    UAVMP_CUDA(ctx, cudaMalloc(&a.d_npath, (size_t)cap * sizeof(int)));
    UAVMP_CUDA(ctx, cudaMalloc(&a.d_misc, 64));
    UAVMP_CUDA(ctx, cudaMalloc(&a.d_hash, (size_t)cap * sizeof(unsigned long long)));
    UAVMP_CUDA(ctx, cudaMalloc(&a.d_path_stage, (size_t)cap * ctx->astar_path_cap * 3 * sizeof(double)));
    UAVMP_CUDA(ctx, cudaMalloc(&a.d_offsets, (size_t)(cap + 1) * sizeof(long long)));
    a.cap = cap; a.path_cap = ctx->astar_path_cap;
  }
  return UAVMP_OK;
}

namespace {
__global__ void k_astar_offsets(const int* n_path, int B, long long* offsets) {  // B is small enough for one thread's scan
  long long acc = 0;
  for (int i = 0; i < B; i++) { offsets[i] = acc; acc += n_path[i]; }
  offsets[B] = acc;
}
__global__ void k_astar_pack(const double* stage, const int* n_path, const long long* offsets, int path_cap, double* out) {
  const int q = blockIdx.x, n = n_path[q];
  const double* src = stage + (size_t)q * path_cap * 3;
  double* dst = out + offsets[q] * 3;
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) dst[i] = src[i];
}
}  // namespace

long long astar_search_batch(uavmp_ctx* ctx, int B, const double* start_pt, const double* end_pt, int* status, int* use_node_num,
                             long long* path_offsets, uint64_t* pop_hash, int* n_pop) {
  int r = astar_ensure(ctx, B);
  if (r) return r;
  AstarState& a = *ctx->astar;
  cudaStream_t st = ctx->stream;
  const size_t nb = (size_t)B * 3 * sizeof(double);
  UAVMP_CUDA(ctx, cudaMemcpyAsync(a.d_q, start_pt, nb, cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(a.d_q + 3 * (size_t)B, end_pt, nb, cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemsetAsync(a.d_misc, 0, 64, st));
  AParams P;
  P.allocated = a.allocated; P.nx = ctx->nx; P.ny = ctx->ny; P.nz = ctx->nz; P.table_bits = a.table_bits;
  P.lambda = ctx->astar_lambda; P.tie = 1.0 + 1e-4;  // a_star.cpp:16
  P.res = ctx->resolution;                           // init() overwrites astar/resolution with the grid map's (:31)
  P.ox = ctx->origin[0]; P.oy = ctx->origin[1]; P.oz = ctx->origin[2];
  P.inv_res = 1.0 / ctx->resolution;
  P.lox = ctx->origin[0] + 1e-4; P.loy = ctx->origin[1] + 1e-4; P.loz = ctx->origin[2] + 1e-4;
  P.hix = (ctx->origin[0] + ctx->map_size[0]) - 1e-4; P.hiy = (ctx->origin[1] + ctx->map_size[1]) - 1e-4; P.hiz = (ctx->origin[2] + ctx->map_size[2]) - 1e-4;
  ABatch bt;
  bt.B = B; bt.start = a.d_q; bt.end = a.d_q + 3 * (size_t)B;
  bt.status = a.d_status; bt.use_num = a.d_use; bt.n_pop = a.d_npop; bt.pop_hash = a.d_hash; bt.n_path = a.d_npath;
  bt.path_stage = a.d_path_stage; bt.path_cap = a.path_cap; bt.next_query = a.d_misc + 1; bt.error_flag = a.d_misc;
  const int grid = std::min(a.n_arenas / AW, (B + AW - 1) / AW);
  if (a.ctas == 8) astar_search_kernel<8><<<grid, AW * 32, 0, st>>>(P, ctx->d_occ, a.d_arenas, bt);
  else astar_search_kernel<4><<<grid, AW * 32, 0, st>>>(P, ctx->d_occ, a.d_arenas, bt);
  UAVMP_CUDA(ctx, cudaGetLastError());
  k_astar_offsets<<<1, 1, 0, st>>>(a.d_npath, B, a.d_offsets);
  long long total = 0;
  int flag = 0;
  UAVMP_CUDA(ctx, cudaMemcpyAsync(&total, a.d_offsets + B, sizeof(long long), cudaMemcpyDeviceToHost, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(&flag, a.d_misc, sizeof(int), cudaMemcpyDeviceToHost, st));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(st));
  if (flag & 4) return uavmp_fail(ctx, UAVMP_ECAP, "A* path longer than %d nodes (uavmp_astar_set_params: path_cap)", a.path_cap);
  if (total > a.packed_cap) {
    if (a.d_packed) cudaFree(a.d_packed);
    a.d_packed = nullptr; a.packed_cap = 0;
    const long long cap = std::max(total, (long long)1024);
    UAVMP_CUDA(ctx, cudaMalloc(&a.d_packed, (size_t)cap * 3 * sizeof(double)));
    a.packed_cap = cap;
  }
  if (total > 0) k_astar_pack<<<B, 128, 0, st>>>(a.d_path_stage, a.d_npath, a.d_offsets, a.path_cap, a.d_packed);
  UAVMP_CUDA(ctx, cudaMemcpyAsync(status, a.d_status, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (use_node_num) UAVMP_CUDA(ctx, cudaMemcpyAsync(use_node_num, a.d_use, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (path_offsets) UAVMP_CUDA(ctx, cudaMemcpyAsync(path_offsets, a.d_offsets, (size_t)(B + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
  if (pop_hash) UAVMP_CUDA(ctx, cudaMemcpyAsync(pop_hash, a.d_hash, (size_t)B * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  if (n_pop) UAVMP_CUDA(ctx, cudaMemcpyAsync(n_pop, a.d_npop, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(st));
  a.last_total = total;
  return total;
}

int astar_get_paths(uavmp_ctx* ctx, double* path_xyz, long long cap_points) {
  if (!ctx->astar) return uavmp_fail(ctx, UAVMP_ESTATE, "no A* batch has run");
  AstarState& a = *ctx->astar;
  if (cap_points < a.last_total) return uavmp_fail(ctx, UAVMP_ECAP, "path buffer too small");
  if (a.last_total > 0) UAVMP_CUDA(ctx, cudaMemcpyAsync(path_xyz, a.d_packed, (size_t)a.last_total * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UAVMP_OK;
}
