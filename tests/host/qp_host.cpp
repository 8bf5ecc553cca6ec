// tests/host/qp_host.cpp — TEST HARNESS (never part of libuavmp.so): compiles the product's K2 body
// (uav_motion_planning_b200/csrc/qp_body.h) and symbolic plan (qp_symbolic.cpp) for the HOST, so that the restated
// OSQP algorithm and the fill-reducing plan can be compared with the reference OSQP (oracle/_ref) without a GPU.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/host_qp.py does it on demand).
#include <vector>

#include "../../uav_motion_planning_b200/csrc/qp_body.h"

extern "C" int host_qp_solve_c(int order, int S, int Kc, int B, const double* pos, const double* bv, const double* ba,
                               const double* bj, const double* T, const double* lo, const double* hi,
                               const uavmp_osqp_settings* st, double* coef, int* solved, int* status, int* iters, int* stats6) {
  QpPlanHost* H = qp_plan_build(order, S, Kc);
  std::vector<int> ints;
  std::vector<double> dbls;
  QpPlanOffsets off;
  QpPlanDev D;
  qp_plan_pack(*H, ints, dbls, off);
  qp_plan_bind(*H, off, ints.data(), dbls.data(), D);
  if (stats6) { stats6[0] = H->n; stats6[1] = H->m; stats6[2] = H->nnzP; stats6[3] = H->nnzA; stats6[4] = H->nnzK; stats6[5] = H->nnzL; }
  const int stride = (B + 31) & ~31;
  // poison the workspace: any read-before-write inside the body shows up as NaN coefficients
  std::vector<double> ws((size_t)D.ws_doubles * stride, fpm::from_bits(0x7ff8000000000000ull));
  QpIo io;
  io.pos = pos; io.bv = bv; io.ba = ba; io.bj = bj ? bj : ba; io.T = T; io.lo = lo; io.hi = hi;
  io.coef = coef; io.solved = solved; io.status = status; io.iters = iters; io.B = B; io.stride = stride;
  for (int b = 0; b < B; b++) qp_solve_one(D, io, *st, ws.data(), b, nullptr);
  delete H;
  return 0;
}
extern "C" int host_qp_solve(int order, int S, int B, const double* pos, const double* bv, const double* ba,
                             const double* bj, const double* T, const uavmp_osqp_settings* st, double* coef,
                             int* solved, int* status, int* iters, int* stats6) {
  return host_qp_solve_c(order, S, 0, B, pos, bv, ba, bj, T, nullptr, nullptr, st, coef, solved, status, iters, stats6);
}

// unpermuted KKT pattern of one (order, S) family, for tests/golden/make_amd_tables.py
extern "C" int host_qp_kkt_pattern_c(int order, int S, int Kc, int* N_out, int* nnz_out, long long* Kp, long long* Ki, int cap,
                                     int* from_table) {
  QpPlanHost* H = qp_plan_build(order, S, Kc);
  *N_out = H->N;
  *nnz_out = (int)H->Ki0.size();
  if (from_table) *from_table = H->perm_from_table ? 1 : 0;
  if (Kp && Ki && cap >= (int)H->Ki0.size()) {
    for (int i = 0; i <= H->N; i++) Kp[i] = H->Kp0[i];
    for (size_t i = 0; i < H->Ki0.size(); i++) Ki[i] = H->Ki0[i];
  }
  delete H;
  return 0;
}

extern "C" int host_qp_kkt_pattern(int order, int S, int* N_out, int* nnz_out, long long* Kp, long long* Ki, int cap,
                                   int* from_table) {
  return host_qp_kkt_pattern_c(order, S, 0, N_out, nnz_out, Kp, Ki, cap, from_table);
}

// ---- the warp-per-problem body (qp_body_warp.h) as "one lane": forward and reversed parallel-loop order ------------------
#define QPW_HOST 1
namespace fwd {
#include "../../uav_motion_planning_b200/csrc/qp_body_warp.h"
}
#define QPW_REVERSED 1
namespace rev {
#include "../../uav_motion_planning_b200/csrc/qp_body_warp.h"
}

static int host_qp_solve_warp_impl(int reversed, int order, int S, int Kc, int B, const double* pos, const double* bv, const double* ba,
                                   const double* bj, const double* T, const double* lo, const double* hi,
                                   const uavmp_osqp_settings* st, double* coef, int* solved, int* status, int* iters) {
  QpPlanHost* H = qp_plan_build(order, S, Kc);
  std::vector<int> ints;
  std::vector<double> dbls;
  QpPlanOffsets off;
  QpPlanDev D;
  qp_plan_pack(*H, ints, dbls, off);
  qp_plan_bind(*H, off, ints.data(), dbls.data(), D);
  D.Sidx = H->Sidx.data(); D.Sch = H->Sch.data();
  QpIo io;
  io.pos = pos; io.bv = bv; io.ba = ba; io.bj = bj ? bj : ba; io.T = T; io.lo = lo; io.hi = hi;
  io.coef = coef; io.solved = solved; io.status = status; io.iters = iters; io.B = B; io.stride = 0;
  for (int b = 0; b < B; b++) {
    std::vector<double> w((size_t)D.ws_warp, fpm::from_bits(0x7ff8000000000000ull));  // poisoned
    if (reversed) rev::qp_warp_solve_one(D, io, *st, w.data(), b, H->Sidx.data()); else fwd::qp_warp_solve_one(D, io, *st, w.data(), b, H->Sidx.data());
  }
  delete H;
  return 0;
}
extern "C" int host_qp_solve_warp(int reversed, int order, int S, int B, const double* pos, const double* bv, const double* ba,
                                  const double* bj, const double* T, const uavmp_osqp_settings* st, double* coef, int* solved,
                                  int* status, int* iters) {
  return host_qp_solve_warp_impl(reversed, order, S, 0, B, pos, bv, ba, bj, T, nullptr, nullptr, st, coef, solved, status, iters);
}
extern "C" int host_qp_solve_warp_c(int reversed, int order, int S, int Kc, int B, const double* pos, const double* bv,
                                    const double* ba, const double* bj, const double* T, const double* lo, const double* hi,
                                    const uavmp_osqp_settings* st, double* coef, int* solved, int* status, int* iters) {
  return host_qp_solve_warp_impl(reversed, order, S, Kc, B, pos, bv, ba, bj, T, lo, hi, st, coef, solved, status, iters);
}
