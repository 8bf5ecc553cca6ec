// qp_body_warp.h — K2w: the same OSQP restatement as qp_body.h, executed by ONE WARP per 1-D problem with the whole
// workspace in shared memory.  Every floating-point operation and its order per element are those of qp_body.h (and hence of
// the reference's OSQP, see there for the file:line citations); what changes is who executes them:
//   * element-wise loops, the column scatter of the forward solve, the rows of the mat-vecs: one element / entry / row per lane;
//   * the backward solve: the columns of one elimination-tree level are independent — one column (a sequential dot) per lane;
//   * max-norms: lane-local max + butterfly (max is exact in any order); the few genuine sums stay sequential on every lane.
// The kernel's duration is then the latency of ~2 * (#levels + #columns) shared-memory steps per ADMM iteration instead of one
// thread's ~2 000 dependent global-memory accesses.
//
// Compiled three ways: device (qp_kernel.cu), host "one lane" forward and host "one lane" REVERSED (tests/host/qp_host.cpp):
// a parallel loop whose result depended on the iteration order would show up as a bit difference between the two host builds.
// (no include guard: tests/host/qp_host.cpp includes this file twice, in two namespaces, with different loop orders)
#include <math.h>
#include <stddef.h>

#include "fpmath.h"
#include "qp_plan.h"
#include "uavmp.h"

#if defined(__CUDACC__) && !defined(QPW_HOST)
#define QPW_HD __device__ __forceinline__
#define QPW_LANE ((int)(threadIdx.x & 31))
#define QPW_PFOR(i, lo, hi) for (int i = (lo) + QPW_LANE; i < (hi); i += 32)
#define QPW_SYNC() __syncwarp()
#define QPW_LDG(p) __ldg(p)
QPW_HD double qpw_max(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
QPW_HD int qpw_any(int v) { return __any_sync(0xffffffffu, v); }
#define QPW_DEVICE_WARP 1
#else
#define QPW_HD static inline
#define QPW_LANE 0
#if defined(QPW_REVERSED)
#define QPW_PFOR(i, lo, hi) for (int i = (hi)-1; i >= (lo); i--)
#else
#define QPW_PFOR(i, lo, hi) for (int i = (lo); i < (hi); i++)
#endif
#define QPW_SYNC() ((void)0)
#define QPW_LDG(p) (*(p))
QPW_HD double qpw_max(double v) { return v; }
QPW_HD int qpw_any(int v) { return v; }
#endif
#define QPW_LANE0 (QPW_LANE == 0)

#define QW_INFTY 1e30
#define QW_MIN_SCALING 1e-4
#define QW_MAX_SCALING 1e4
#define QW_RHO_MIN 1e-6
#define QW_RHO_MAX 1e6
#define QW_RHO_TOL 1e-4
#define QW_RHO_EQ 1e3
#define QW_DIV_TOL (1.0 / QW_INFTY)
enum { QW_SOLVED = 1, QW_SOLVED_INACC = 2, QW_PINF = 3, QW_PINF_INACC = 4, QW_DINF = 5, QW_DINF_INACC = 6, QW_MAXITER = 7,
       QW_NONCVX = 9, QW_UNSOLVED = 11 };

#define WW(off, i) w[(off) + (i)]

QPW_HD double qw_limit(double v) {
  v = v < QW_MIN_SCALING ? 1.0 : v;
  v = v > QW_MAX_SCALING ? QW_MAX_SCALING : v;
  return v;
}

// numeric LDL' (QDLDL_factor's arithmetic, qp_body.h qp_factor): rows in order; inside a row the reach columns in order; the
// update of the scatter vector by one reach column touches distinct rows -> one entry per lane
QPW_HD int qw_factor(const QpPlanDev& pl, double* w, double sigma) {
  const int N = pl.N;
  int positive = 0;
  QPW_PFOR(i, 0, N) WW(pl.o_yw, i) = 0.0;
  QPW_SYNC();
  for (int k = 0; k < N; k++) {
    const int p0 = QPW_LDG(pl.Kp + k), p1 = QPW_LDG(pl.Kp + k + 1);
    double Dk = 0.0;
    for (int p = p0; p < p1; p++) {  // every lane evaluates the column (<= a dozen entries)
      const int i = QPW_LDG(pl.Ki + p), kind = QPW_LDG(pl.Kkind + p), idx = QPW_LDG(pl.Kidx + p);
      double v;
      if (kind == 0) v = WW(pl.o_Px, idx);
      else if (kind == 1) v = WW(pl.o_Px, idx) + sigma;
      else if (kind == 2) v = sigma;
      else if (kind == 3) v = WW(pl.o_Ax, idx);
      else v = -WW(pl.o_rhoinv, idx);
      if (i == k) Dk = v;
      else WW(pl.o_yw, i) = v;  // every lane stores the same value: benign
    }
    QPW_SYNC();
    for (int e = QPW_LDG(pl.Rp + k); e < QPW_LDG(pl.Rp + k + 1); e++) {
      const int c = QPW_LDG(pl.Rc + e), pos = QPW_LDG(pl.Rpos + e);
      const double yc = WW(pl.o_yw, c);
      QPW_SYNC();
      QPW_PFOR(j, QPW_LDG(pl.Lp + c), pos) {
        const int r = QPW_LDG(pl.Li + j);
        WW(pl.o_yw, r) = WW(pl.o_yw, r) - WW(pl.o_Lx, j) * yc;
      }
      const double lv = yc * WW(pl.o_Ddinv, c);
      Dk -= yc * lv;
      if (QPW_LANE0) { WW(pl.o_Lx, pos) = lv; WW(pl.o_yw, c) = 0.0; }
      QPW_SYNC();
    }
    if (Dk == 0.0) return -1;
    if (Dk > 0.0) positive++;
    if (QPW_LANE0) { WW(pl.o_Dd, k) = Dk; WW(pl.o_Ddinv, k) = 1.0 / Dk; }
    QPW_SYNC();
  }
  return positive;
}

// xz <- K^-1 xz (QDLDL_solve between the two permutations).  `sx` = the plan's uint16 index block (pl.sx_* offsets), staged in
// shared memory by the kernels.  Both triangular solves run level by level with one row / column per lane:
//   forward  (QDLDL_Lsolve: for i: for j in col i: x[Li[j]] -= Lx[j] * x[i]): row r receives its subtractions in the order of
//            increasing column i, each as x[r] = x[r] - Lx * x[i] — exactly the sequence of the column-oriented loop, so the row
//            form below performs the same operations on the same values (x[i] is final once every row of the levels before it
//            is done);
//   backward (QDLDL_Ltsolve: for i = N-1..0: for j in col i: x[i] -= Lx[j] * x[Li[j]]): column i's own loop.
// The dots stay sequential per lane (summation order is part of the contract); their loads are hoisted four entries ahead of
// the dependent subtract chain.
QPW_HD void qw_tri_dot(const double* w_lx, const double* w_bp, const unsigned short* ix_a, const unsigned short* ix_b, int p0, int p1,
                       double& val) {
  int p = p0;
  for (; p + 4 <= p1; p += 4) {
    const double a0 = w_lx[ix_a ? ix_a[p] : p], a1 = w_lx[ix_a ? ix_a[p + 1] : p + 1], a2 = w_lx[ix_a ? ix_a[p + 2] : p + 2],
                 a3 = w_lx[ix_a ? ix_a[p + 3] : p + 3];
    const double b0 = w_bp[ix_b[p]], b1 = w_bp[ix_b[p + 1]], b2 = w_bp[ix_b[p + 2]], b3 = w_bp[ix_b[p + 3]];
    val = val - a0 * b0; val = val - a1 * b1; val = val - a2 * b2; val = val - a3 * b3;
  }
  for (; p < p1; p++) val = val - w_lx[ix_a ? ix_a[p] : p] * w_bp[ix_b[p]];
}
// One triangular solve from the lane schedule (qp_symbolic.cpp: QpPlanHost::Sch).  A chunk holds the rows (columns) of one
// dependency level: every lane forms ONE product Lx * x (all of a level's products in a single step), then each row's leader lane
// subtracts its row's products in their sequential order, pulling them from the lanes to its right with shuffles — the summation
// order per element is the reference's (QDLDL_Lsolve / QDLDL_Ltsolve), only the loads and multiplications run side by side.
// (The in-lane loops of the first version cost ~15 dependent instructions per L entry on a warp that had two or three rows to
// work on: 85 % of the kernel, profiles/r02_reading.md.)  The schedule is read from global memory one chunk ahead.
#if defined(QPW_DEVICE_WARP)
// pb: 64 doubles of this warp's workspace (16 B aligned).  Every lane parks its product there; a leader (always an even lane)
// then reads its row's products two per load and subtracts them in order.  (Pulling them with shuffles cost two SHFL per
// entry plus the lane arithmetic: the chain was 80 % of the solve's instructions.)
QPW_HD void qw_tri_sched(const unsigned int* sch, int c0, int c1, const double* lx, double* bp, double* pb) {
  const int lane = QPW_LANE;
  uint2 rec = __ldg(reinterpret_cast<const uint2*>(sch) + (size_t)c0 * 32 + lane);
  const double2* mine = reinterpret_cast<const double2*>(pb + (lane & ~1));
  for (int c = c0; c < c1; c++) {
    uint2 nxt = rec;
    if (c + 1 < c1) nxt = __ldg(reinterpret_cast<const uint2*>(sch) + (size_t)(c + 1) * 32 + lane);
    const unsigned elx = rec.x & 0xffffu, esrc = rec.x >> 16, row = rec.y & 0xffffu;
    const int len = (int)((rec.y >> 16) & 0xffu), maxlen = (int)(rec.y >> 24);  // maxlen: a multiple of 4
    double prod = 0.0;
    if (elx != 0xffffu) prod = lx[elx] * bp[esrc];
    pb[lane] = prod;
    double val = 0.0;
    if (len) val = bp[row];
    QPW_SYNC();
    for (int k = 0; k < maxlen; k += 4) {
      const double2 p01 = mine[(k >> 1)], p23 = mine[(k >> 1) + 1];
      if (k < len) val = val - p01.x;
      if (k + 1 < len) val = val - p01.y;
      if (k + 2 < len) val = val - p23.x;
      if (k + 3 < len) val = val - p23.y;
    }
    if (len) bp[row] = val;
    QPW_SYNC();
    rec = nxt;
  }
}
#else
QPW_HD void qw_tri_sched(const unsigned int* sch, int c0, int c1, const double* lx, double* bp, double*) {  // the 32 lanes, one after the other
  for (int c = c0; c < c1; c++) {
    const unsigned int* rec = sch + (size_t)c * 64;
    double prod[32], out[32];
    for (int l = 0; l < 32; l++) {
      const unsigned elx = rec[2 * l] & 0xffffu, esrc = rec[2 * l] >> 16;
      prod[l] = (elx != 0xffffu) ? lx[elx] * bp[esrc] : 0.0;
    }
#if defined(QPW_REVERSED)
    for (int l = 31; l >= 0; l--) {
#else
    for (int l = 0; l < 32; l++) {
#endif
      const int len = (int)((rec[2 * l + 1] >> 16) & 0xffu);
      if (!len) continue;
      double val = bp[rec[2 * l + 1] & 0xffffu];
      for (int k = 0; k < len; k++) val = val - prod[l + k];
      out[l] = val;
    }
    for (int l = 0; l < 32; l++) if ((rec[2 * l + 1] >> 16) & 0xffu) bp[rec[2 * l + 1] & 0xffffu] = out[l];
  }
}
#endif

QPW_HD void qw_kkt_solve(const QpPlanDev& pl, double* w, const unsigned short* sx) {
  const int N = pl.N;
  const unsigned short *perm = sx + pl.sx_perm, *Lrp = sx + pl.sx_Lrp, *Lrc = sx + pl.sx_Lrc, *Lrx = sx + pl.sx_Lrx,
                       *LevFP = sx + pl.sx_LevFP, *LevFR = sx + pl.sx_LevFR, *Lp = sx + pl.sx_Lp, *Li = sx + pl.sx_Li,
                       *LevP = sx + pl.sx_LevP, *LevC = sx + pl.sx_LevC;
  double* bp = w + pl.o_bp;
  const double* lx = w + pl.o_Lx;
  QPW_PFOR(j, 0, N) bp[j] = WW(pl.o_xz, perm[j]);
  QPW_SYNC();
  if (pl.sch_n > 0) {
    qw_tri_sched(pl.Sch, 0, pl.sch_nf, lx, bp, w + pl.o_pb);
    QPW_PFOR(i, 0, N) bp[i] = bp[i] * WW(pl.o_Ddinv, i);
    QPW_SYNC();
    qw_tri_sched(pl.Sch, pl.sch_nf, pl.sch_n, lx, bp, w + pl.o_pb);
  } else {  // a row with more than 32 entries: level by level with one row / column per lane
    for (int lv = 0; lv < pl.nlevf; lv++) {
      QPW_PFOR(ri, LevFP[lv], LevFP[lv + 1]) {
        const int r = LevFR[ri];
        double val = bp[r];
        qw_tri_dot(lx, bp, Lrx, Lrc, Lrp[r], Lrp[r + 1], val);
        bp[r] = val;
      }
      QPW_SYNC();
    }
    QPW_PFOR(i, 0, N) bp[i] = bp[i] * WW(pl.o_Ddinv, i);
    QPW_SYNC();
    for (int lv = 0; lv < pl.nlev; lv++) {
      QPW_PFOR(ci, LevP[lv], LevP[lv + 1]) {
        const int i = LevC[ci];
        double val = bp[i];
        qw_tri_dot(lx, bp, nullptr, Li, Lp[i], Lp[i + 1], val);
        bp[i] = val;
      }
      QPW_SYNC();
    }
  }
  QPW_PFOR(j, 0, N) WW(pl.o_xz, perm[j]) = bp[j];
  QPW_SYNC();
}

// out(m) = A v : row r accumulates its entries in column order (== the CSC loop's order per row), starting from 0.0
QPW_HD void qw_A_mul(const QpPlanDev& pl, double* w, int o_v, int o_out) {
  QPW_PFOR(r, 0, pl.m) {
    double acc = 0.0;
    for (int p = QPW_LDG(pl.Arp + r); p < QPW_LDG(pl.Arp + r + 1); p++) acc = acc + WW(pl.o_Ax, QPW_LDG(pl.Arx + p)) * WW(o_v, QPW_LDG(pl.Arj + p));
    WW(o_out, r) = acc;
  }
  QPW_SYNC();
}
QPW_HD void qw_At_mul(const QpPlanDev& pl, double* w, int o_v, int o_out) {
  QPW_PFOR(c, 0, pl.n) {
    double acc = 0.0;
    for (int p = QPW_LDG(pl.Ap + c); p < QPW_LDG(pl.Ap + c + 1); p++) acc += WW(pl.o_Ax, p) * WW(o_v, QPW_LDG(pl.Ai + p));
    WW(o_out, c) = acc;
  }
  QPW_SYNC();
}
// out(n) = P v with P's upper triangle: element i receives its contributions in the order the CSC loop produces them
QPW_HD void qw_P_mul(const QpPlanDev& pl, double* w, int o_v, int o_out) {
  QPW_PFOR(i, 0, pl.n) {
    double acc = 0.0;
    for (int e = QPW_LDG(pl.Psp + i); e < QPW_LDG(pl.Psp + i + 1); e++) acc = acc + WW(pl.o_Px, QPW_LDG(pl.Psa + e)) * WW(o_v, QPW_LDG(pl.Psv + e));
    WW(o_out, i) = acc;
  }
  QPW_SYNC();
}
QPW_HD double qw_norm_inf(double* w, int o_v, int len) {
  double r = 0.0;
  QPW_PFOR(i, 0, len) r = fmax(r, fabs(WW(o_v, i)));
  return qpw_max(r);
}
QPW_HD double qw_scaled_norm_inf(double* w, int o_s, int o_v, int len) {
  double r = 0.0;
  QPW_PFOR(i, 0, len) r = fmax(r, fabs(WW(o_s, i) * WW(o_v, i)));
  return qpw_max(r);
}

struct QwResid { double prim_res, dual_res, scaled_prim, scaled_dual; };

QPW_HD void qw_update_info(const QpPlanDev& pl, double* w, double cinv, QwResid& R) {
  const int n = pl.n, m = pl.m;
  qw_A_mul(pl, w, pl.o_x, pl.o_Axv);
  double sp = 0.0, up = 0.0;
  QPW_PFOR(i, 0, m) {
    const double d = WW(pl.o_Axv, i) - WW(pl.o_z, i);
    sp = fmax(sp, fabs(d));
    up = fmax(up, fabs(WW(pl.o_Einv, i) * d));
  }
  R.scaled_prim = qpw_max(sp); R.prim_res = qpw_max(up);
  qw_P_mul(pl, w, pl.o_x, pl.o_Pxv);
  qw_At_mul(pl, w, pl.o_y, pl.o_Aty);
  double sd = 0.0, ud = 0.0;
  QPW_PFOR(i, 0, n) {
    const double d = (WW(pl.o_q, i) + WW(pl.o_Pxv, i)) + WW(pl.o_Aty, i);
    sd = fmax(sd, fabs(d));
    ud = fmax(ud, fabs(WW(pl.o_Dinv, i) * d));
  }
  R.scaled_dual = qpw_max(sd); R.dual_res = cinv * qpw_max(ud);
}

QPW_HD int qw_check_termination(const QpPlanDev& pl, double* w, const uavmp_osqp_settings& S, double c, double cinv,
                                const QwResid& R, bool approximate) {
  const int n = pl.n, m = pl.m;
  double eps_abs = S.eps_abs, eps_rel = S.eps_rel, eps_pinf = S.eps_prim_inf, eps_dinf = S.eps_dual_inf;
  if (R.prim_res > QW_INFTY || R.dual_res > QW_INFTY) return QW_NONCVX;
  if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pinf *= 10; eps_dinf *= 10; }
  bool prim_ok = false, dual_ok = false, pinf = false, dinf = false;
  {
    const double mx = fmax(qw_scaled_norm_inf(w, pl.o_Einv, pl.o_z, m), qw_scaled_norm_inf(w, pl.o_Einv, pl.o_Axv, m));
    const double eps_prim = eps_abs + eps_rel * mx;
    if (R.prim_res < eps_prim) {
      prim_ok = true;
    } else {
      QPW_PFOR(i, 0, m) {
        const double l = WW(pl.o_l, i), u = WW(pl.o_u, i);
        double dy = WW(pl.o_dy, i);
        if (u > QW_INFTY * QW_MIN_SCALING) {
          if (l < -QW_INFTY * QW_MIN_SCALING) dy = 0.0; else dy = fmin(dy, 0.0);
        } else if (l < -QW_INFTY * QW_MIN_SCALING) {
          dy = fmax(dy, 0.0);
        }
        WW(pl.o_dy, i) = dy;
      }
      QPW_SYNC();
      const double norm_dy = qw_scaled_norm_inf(w, pl.o_E, pl.o_dy, m);
      if (norm_dy > QW_DIV_TOL) {
        double lhs = 0.0, lhs2 = 0.0;  // genuine sums: sequential, every lane computes the same value
        for (int i = 0; i < m; i++) { const double dy = WW(pl.o_dy, i); lhs += WW(pl.o_u, i) * fmax(dy, 0.0); }
        for (int i = 0; i < m; i++) { const double dy = WW(pl.o_dy, i); lhs2 += WW(pl.o_l, i) * fmin(dy, 0.0); }
        lhs += lhs2;
        if (lhs < 0.0) {
          qw_At_mul(pl, w, pl.o_dy, pl.o_tn);
          pinf = qw_scaled_norm_inf(w, pl.o_Dinv, pl.o_tn, n) < eps_pinf * norm_dy;
        }
      }
    }
  }
  {
    double mx = fmax(fmax(qw_scaled_norm_inf(w, pl.o_Dinv, pl.o_q, n), qw_scaled_norm_inf(w, pl.o_Dinv, pl.o_Aty, n)),
                     qw_scaled_norm_inf(w, pl.o_Dinv, pl.o_Pxv, n));
    mx *= cinv;
    const double eps_dual = eps_abs + eps_rel * mx;
    if (R.dual_res < eps_dual) {
      dual_ok = true;
    } else {
      const double norm_dx = qw_scaled_norm_inf(w, pl.o_D, pl.o_dx, n);
      if (norm_dx > QW_DIV_TOL) {
        double qdx = 0.0;
        for (int i = 0; i < n; i++) qdx += WW(pl.o_q, i) * WW(pl.o_dx, i);
        if (qdx < 0.0) {
          qw_P_mul(pl, w, pl.o_dx, pl.o_tn);
          if (qw_scaled_norm_inf(w, pl.o_Dinv, pl.o_tn, n) < c * eps_dinf * norm_dx) {
            qw_A_mul(pl, w, pl.o_dx, pl.o_tm);
            int out_cone = 0;
            const double tol = eps_dinf * norm_dx;
            QPW_PFOR(i, 0, m) {
              const double v = WW(pl.o_Einv, i) * WW(pl.o_tm, i);
              if ((WW(pl.o_u, i) < QW_INFTY * QW_MIN_SCALING && v > tol) || (WW(pl.o_l, i) > -QW_INFTY * QW_MIN_SCALING && v < -tol)) out_cone = 1;
            }
            dinf = !qpw_any(out_cone);
          }
        }
      }
    }
  }
  if (prim_ok && dual_ok) return approximate ? QW_SOLVED_INACC : QW_SOLVED;
  if (pinf) return approximate ? QW_PINF_INACC : QW_PINF;
  if (dinf) return approximate ? QW_DINF_INACC : QW_DINF;
  return 0;
}

QPW_HD void qw_set_rho(const QpPlanDev& pl, double* w, double rho_) {
  QPW_PFOR(i, 0, pl.m) {
    const double l = WW(pl.o_l, i), u = WW(pl.o_u, i);
    double r;
    if (l < -QW_INFTY * QW_MIN_SCALING && u > QW_INFTY * QW_MIN_SCALING) r = QW_RHO_MIN;
    else if (u - l < QW_RHO_TOL) r = QW_RHO_EQ * rho_;
    else r = rho_;
    WW(pl.o_rho, i) = r;
    WW(pl.o_rhoinv, i) = 1.0 / r;
  }
  QPW_SYNC();
}

// one problem per warp: assembly -> osqp_setup -> osqp_solve -> store_solution.  `w` = this warp's workspace (pl.ws_warp doubles)
QPW_HD void qp_warp_solve_one(const QpPlanDev& pl, const QpIo& io, const uavmp_osqp_settings& S, double* w, int b, const unsigned short* sx) {
  const int n = pl.n, m = pl.m, Sg = pl.S;
  // ---- assembly (minimum_control.cpp:5-125) ----------------------------------------------------------------------------
  const double* T = io.T + (size_t)b * Sg;
  QPW_PFOR(p, 0, pl.nnzP) WW(pl.o_Px, p) = QPW_LDG(pl.P_coef + p) * fpm::powi(T[QPW_LDG(pl.P_seg + p)], QPW_LDG(pl.P_pow + p));
  QPW_PFOR(p, 0, pl.nnzA) WW(pl.o_Ax, p) = QPW_LDG(pl.A_coef + p) * fpm::powi(T[QPW_LDG(pl.A_seg + p)], QPW_LDG(pl.A_pow + p));
  QPW_PFOR(i, 0, n) { WW(pl.o_q, i) = 0.0; WW(pl.o_D, i) = 1.0; }
  QPW_PFOR(i, 0, m) {
    WW(pl.o_l, i) = qp_bound_value(io, b, Sg, QPW_LDG(pl.l_src + i));
    WW(pl.o_u, i) = qp_bound_value(io, b, Sg, QPW_LDG(pl.u_src + i));
    WW(pl.o_E, i) = 1.0;
  }
  QPW_SYNC();
  // ---- scale_data (scaling.c:49-165) ---------------------------------------------------------------------------------
  double c = 1.0;
  for (int it = 0; it < S.scaling; it++) {
    QPW_PFOR(j, 0, n) {
      double dn = 0.0;
      for (int p = QPW_LDG(pl.Pp + j); p < QPW_LDG(pl.Pp + j + 1); p++) dn = fmax(fabs(WW(pl.o_Px, p)), dn);
      double an = 0.0;
      for (int p = QPW_LDG(pl.Ap + j); p < QPW_LDG(pl.Ap + j + 1); p++) an = fmax(fabs(WW(pl.o_Ax, p)), an);
      WW(pl.o_tn, j) = 1.0 / sqrt(qw_limit(fmax(dn, an)));
    }
    QPW_PFOR(r, 0, m) {
      double rn = 0.0;
      for (int p = QPW_LDG(pl.Arp + r); p < QPW_LDG(pl.Arp + r + 1); p++) rn = fmax(fabs(WW(pl.o_Ax, QPW_LDG(pl.Arx + p))), rn);
      WW(pl.o_tm, r) = 1.0 / sqrt(qw_limit(rn));
    }
    QPW_SYNC();
    QPW_PFOR(j, 0, n) {  // P <- D P D, A <- E A D, q <- D q
      const double dj = WW(pl.o_tn, j);
      for (int p = QPW_LDG(pl.Pp + j); p < QPW_LDG(pl.Pp + j + 1); p++) {
        const double v = WW(pl.o_Px, p) * WW(pl.o_tn, QPW_LDG(pl.Pi + p));
        WW(pl.o_Px, p) = v * dj;
      }
      for (int p = QPW_LDG(pl.Ap + j); p < QPW_LDG(pl.Ap + j + 1); p++) {
        const double v = WW(pl.o_Ax, p) * WW(pl.o_tm, QPW_LDG(pl.Ai + p));
        WW(pl.o_Ax, p) = v * dj;
      }
      WW(pl.o_q, j) = WW(pl.o_q, j) * dj;
      WW(pl.o_D, j) = WW(pl.o_D, j) * dj;
    }
    QPW_PFOR(i, 0, m) WW(pl.o_E, i) = WW(pl.o_E, i) * WW(pl.o_tm, i);
    QPW_SYNC();
    QPW_PFOR(j, 0, n) {  // column norms of the scaled P, then their (sequential) sum
      double dn = 0.0;
      for (int p = QPW_LDG(pl.Pp + j); p < QPW_LDG(pl.Pp + j + 1); p++) dn = fmax(fabs(WW(pl.o_Px, p)), dn);
      WW(pl.o_tn, j) = fabs(dn);
    }
    QPW_SYNC();
    double sum = 0.0;
    for (int j = 0; j < n; j++) sum += WW(pl.o_tn, j);
    double c_temp = sum / n;
    const double inf_q = qw_limit(qw_norm_inf(w, pl.o_q, n));
    c_temp = fmax(c_temp, inf_q);
    c_temp = qw_limit(c_temp);
    c_temp = 1.0 / c_temp;
    QPW_SYNC();
    QPW_PFOR(p, 0, pl.nnzP) WW(pl.o_Px, p) = WW(pl.o_Px, p) * c_temp;
    QPW_PFOR(j, 0, n) WW(pl.o_q, j) = WW(pl.o_q, j) * c_temp;
    QPW_SYNC();
    c *= c_temp;
  }
  const double cinv = 1.0 / c;
  QPW_PFOR(j, 0, n) WW(pl.o_Dinv, j) = 1.0 / WW(pl.o_D, j);
  QPW_PFOR(i, 0, m) {
    const double e = WW(pl.o_E, i);
    WW(pl.o_Einv, i) = 1.0 / e;
    WW(pl.o_l, i) = WW(pl.o_l, i) * e;
    WW(pl.o_u, i) = WW(pl.o_u, i) * e;
  }
  QPW_SYNC();
  double rho = fmin(fmax(S.rho, QW_RHO_MIN), QW_RHO_MAX);
  qw_set_rho(pl, w, rho);

  int status = QW_UNSOLVED, iter_out = 0;
  if (qw_factor(pl, w, S.sigma) < n) status = QW_NONCVX;
  if (status == QW_UNSOLVED) {
    QPW_PFOR(i, 0, n) { WW(pl.o_x, i) = 0.0; WW(pl.o_xprev, i) = 0.0; }
    QPW_PFOR(i, 0, m) { WW(pl.o_z, i) = 0.0; WW(pl.o_zprev, i) = 0.0; WW(pl.o_y, i) = 0.0; }
    QPW_SYNC();
    const int interval = S.adaptive_rho_interval ? S.adaptive_rho_interval : (S.check_termination ? 4 * S.check_termination : 100);
    const double alpha = S.alpha, sigma = S.sigma, one_m_alpha = 1.0 - S.alpha;
    QwResid R;
    R.prim_res = R.dual_res = R.scaled_prim = R.scaled_dual = QW_INFTY;
    bool checked_last = false;
    int iter;
    for (iter = 1; iter <= S.max_iter; iter++) {
      QPW_PFOR(i, 0, n) {  // compute_rhs
        const double xv = WW(pl.o_x, i);
        WW(pl.o_xprev, i) = xv;
        WW(pl.o_xz, i) = sigma * xv + (-1.0) * WW(pl.o_q, i);
      }
      QPW_PFOR(i, 0, m) {
        const double zv = WW(pl.o_z, i);
        WW(pl.o_zprev, i) = zv;
        const double t = WW(pl.o_rhoinv, i) * WW(pl.o_y, i);
        const double rhs = (-1.0) * t + 1.0 * zv;
        WW(pl.o_xz, n + i) = rhs;
        WW(pl.o_tm, i) = rhs;
      }
      QPW_SYNC();
      qw_kkt_solve(pl, w, sx);
      QPW_PFOR(i, 0, n) {  // update_x
        const double xp = WW(pl.o_xprev, i);
        const double xn = alpha * WW(pl.o_xz, i) + one_m_alpha * xp;
        WW(pl.o_x, i) = xn;
        WW(pl.o_dx, i) = xn - xp;
      }
      QPW_PFOR(i, 0, m) {  // ztilde, update_z, update_y
        const double rv = WW(pl.o_rhoinv, i), zp = WW(pl.o_zprev, i), yv = WW(pl.o_y, i);
        const double zt = WW(pl.o_tm, i) + rv * WW(pl.o_xz, n + i);
        double zn = rv * yv;
        zn = zn + (alpha * zt + one_m_alpha * zp);  // add_scaled3's incrementing form (vector.c:441-444), see qp_body.h
        zn = fmin(fmax(zn, WW(pl.o_l, i)), WW(pl.o_u, i));
        WW(pl.o_z, i) = zn;
        double dy = (alpha * zt + one_m_alpha * zp) + (-1.0) * zn;
        dy = dy * WW(pl.o_rho, i);
        WW(pl.o_dy, i) = dy;
        WW(pl.o_y, i) = yv + dy;
      }
      QPW_SYNC();
      const bool can_check = S.check_termination && (iter % S.check_termination == 0);
      checked_last = can_check;
      if (can_check) {
        qw_update_info(pl, w, cinv, R);
        iter_out = iter;
        const int st = qw_check_termination(pl, w, S, c, cinv, R, false);
        if (st) { status = st; break; }
      }
      if (S.adaptive_rho && interval && (iter % interval == 0)) {
        if (!can_check) { qw_update_info(pl, w, cinv, R); iter_out = iter; }
        double pr = R.scaled_prim, dr = R.scaled_dual;
        const double pn = fmax(qw_norm_inf(w, pl.o_z, m), qw_norm_inf(w, pl.o_Axv, m));
        pr /= (pn + QW_DIV_TOL);
        const double dn = fmax(fmax(qw_norm_inf(w, pl.o_q, n), qw_norm_inf(w, pl.o_Aty, n)), qw_norm_inf(w, pl.o_Pxv, n));
        dr /= (dn + QW_DIV_TOL);
        double rho_new = rho * sqrt(pr / dr);
        rho_new = fmin(fmax(rho_new, QW_RHO_MIN), QW_RHO_MAX);
        if (rho_new > rho * S.adaptive_rho_tolerance || rho_new < rho / S.adaptive_rho_tolerance) {
          rho = fmin(fmax(rho_new, QW_RHO_MIN), QW_RHO_MAX);
          qw_set_rho(pl, w, rho);
          if (qw_factor(pl, w, sigma) < 0) { status = QW_NONCVX; break; }
        }
      }
    }
    if (status == QW_UNSOLVED) {
      if (!checked_last) {
        qw_update_info(pl, w, cinv, R);
        iter_out = iter - 1;
        const int st = qw_check_termination(pl, w, S, c, cinv, R, false);
        if (st) status = st;
      }
      if (status == QW_UNSOLVED) {
        const int st = qw_check_termination(pl, w, S, c, cinv, R, true);
        status = st ? st : QW_MAXITER;
      }
    }
  }
  // ---- store_solution: x = D x_scaled, NaN when there is no solution ----------------------------------------------------
  const bool has_sol = !(status == QW_PINF || status == QW_PINF_INACC || status == QW_DINF || status == QW_DINF_INACC || status == QW_NONCVX);
  double* out = io.coef + (size_t)b * n;
  QPW_PFOR(i, 0, n) out[i] = has_sol ? WW(pl.o_D, i) * WW(pl.o_x, i) : fpm::from_bits(0x7ff8000000000000ull);
  if (QPW_LANE0) {
    io.status[b] = status;
    io.iters[b] = iter_out;
    io.solved[b] = (status == QW_SOLVED) ? 1 : 0;
  }
  QPW_SYNC();
}

#undef WW
#undef QPW_HD
#undef QPW_LANE
#undef QPW_PFOR
#undef QPW_SYNC
#undef QPW_LDG
#undef QPW_LANE0
#undef QPW_DEVICE_WARP
#undef QW_INFTY
#undef QW_MIN_SCALING
#undef QW_MAX_SCALING
#undef QW_RHO_MIN
#undef QW_RHO_MAX
#undef QW_RHO_TOL
#undef QW_RHO_EQ
#undef QW_DIV_TOL
