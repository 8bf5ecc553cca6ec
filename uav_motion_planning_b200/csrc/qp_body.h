// qp_body.h — the per-problem body of K2 (batched minimum-jerk / minimum-snap QP), written once and compiled twice:
//   * by nvcc as the device code of qp_solve_kernel (qp_kernel.cu) — the product;
//   * by g++ (-ffp-contract=off) inside tests/host/qp_host.cpp — a TEST harness that runs the identical statements
//     on the host so the symbolic plan (qp_symbolic.cpp) and the ADMM restatement can be checked without a GPU.
// The product never runs the host instantiation: libuavmp.so only contains the __global__ wrapper.
//
// Restates the algorithm of the reference's vendored OSQP 1.0.0.beta0 (file:line in the comments below).
#pragma once
#include <math.h>
#include <stddef.h>

#include "fpmath.h"
#include "qp_plan.h"
#include "uavmp.h"

#if defined(__CUDACC__)
#define QP_HD __device__ __forceinline__
#define QP_LDG(p) __ldg(p)
#define QP_TID ((int)threadIdx.x)
#else
#define QP_HD static inline
#define QP_LDG(p) (*(p))
#define QP_TID 0
#endif
#define QP_TPB 32  // threads per CTA of the device instantiation (one problem per thread)

// The two vectors every triangular solve hammers (the permuted right-hand side `bp` and the KKT vector `xz`; the
// factorisation's scatter vector shares bp's storage) can live in shared memory, [element][thread] so that a warp's
// accesses are conflict-free: sv != nullptr selects that placement, otherwise they stay in the global workspace.
#define WBP(i) (*(sv ? &sv[(size_t)(i) * QP_TPB + QP_TID] : &W(pl.o_bp, i)))
#define WXZ(i) (*(sv ? &sv[(size_t)(pl.N + (i)) * QP_TPB + QP_TID] : &W(pl.o_xz, i)))
#define WYW(i) (*(sv ? &sv[(size_t)(i) * QP_TPB + QP_TID] : &W(pl.o_yw, i)))

#define OSQP_INFTY_ 1e30
#define OSQP_MIN_SCALING_ 1e-4
#define OSQP_MAX_SCALING_ 1e4
#define OSQP_RHO_MIN_ 1e-6
#define OSQP_RHO_MAX_ 1e6
#define OSQP_RHO_TOL_ 1e-4
#define OSQP_RHO_EQ_OVER_RHO_INEQ_ 1e3
#define OSQP_DIVISION_TOL_ (1.0 / OSQP_INFTY_)

enum { ST_SOLVED = 1, ST_SOLVED_INACC = 2, ST_PINF = 3, ST_PINF_INACC = 4, ST_DINF = 5, ST_DINF_INACC = 6,
       ST_MAXITER = 7, ST_NONCVX = 9, ST_UNSOLVED = 11 };


QP_HD double limit_scaling(double v) {
  v = v < OSQP_MIN_SCALING_ ? 1.0 : v;
  v = v > OSQP_MAX_SCALING_ ? OSQP_MAX_SCALING_ : v;
  return v;
}

#define W(off, i) ws[((size_t)((off) + (i))) * stride + b]

// numeric LDL' of the permuted KKT matrix (up-looking, static reach lists) — QDLDL_factor's arithmetic
QP_HD int qp_factor(const QpPlanDev& pl, double* ws, size_t stride, int b, double sigma, double* sv) {
  const int N = pl.N;
  int positive = 0;
  // the scatter vector starts at zero (QDLDL_factor clears yVals the same way); it shares storage with bp when in
  // shared memory, and the global workspace is reused across batches without being zeroed by the host
  for (int i = 0; i < N; i++) WYW(i) = 0.0;
  for (int k = 0; k < N; k++) {
    double Dk = 0.0;
    for (int p = QP_LDG(pl.Kp + k); p < QP_LDG(pl.Kp + k + 1); p++) {
      const int i = QP_LDG(pl.Ki + p), kind = QP_LDG(pl.Kkind + p), idx = QP_LDG(pl.Kidx + p);
      double v;
      if (kind == 0) v = W(pl.o_Px, idx);
      else if (kind == 1) v = W(pl.o_Px, idx) + sigma;
      else if (kind == 2) v = sigma;
      else if (kind == 3) v = W(pl.o_Ax, idx);
      else v = -W(pl.o_rhoinv, idx);
      if (i == k) Dk = v; else WYW(i) = v;
    }
    for (int e = QP_LDG(pl.Rp + k); e < QP_LDG(pl.Rp + k + 1); e++) {
      const int c = QP_LDG(pl.Rc + e), pos = QP_LDG(pl.Rpos + e);
      const double yc = WYW(c);
      for (int j = QP_LDG(pl.Lp + c); j < pos; j++) {
        const int r = QP_LDG(pl.Li + j);
        WYW(r) = WYW(r) - W(pl.o_Lx, j) * yc;
      }
      const double lv = yc * W(pl.o_Ddinv, c);
      W(pl.o_Lx, pos) = lv;
      W(pl.o_LxT, QP_LDG(pl.Ltpos + pos)) = lv;  // second copy in the order the L' solve consumes it
      Dk -= yc * lv;
      WYW(c) = 0.0;
    }
    if (Dk == 0.0) return -1;
    if (Dk > 0.0) positive++;
    W(pl.o_Dd, k) = Dk;
    W(pl.o_Ddinv, k) = 1.0 / Dk;
  }
  return positive;
}

// xz <- K^-1 xz  (qdldl_interface.c:394-415: permute, L solve, D^-1, L' solve, permute back).
// The arithmetic and its order are QDLDL_solve's; only the FETCH of L's values is reorganised: both triangular solves
// consume L as a linear stream (Lx in column order, LxT in the L' solve's order), fetched 8 values at a time so that a
// thread has 8 loads in flight instead of one.
QP_HD void qp_kkt_solve(const QpPlanDev& pl, double* ws, size_t stride, int b, double* sv) {
  const int N = pl.N, nnzL = pl.nnzL;
  for (int j = 0; j < N; j++) WBP(j) = WXZ(QP_LDG(pl.perm + j));
  {  // QDLDL_Lsolve: for i: val = x[i]; for j in col i: x[Li[j]] -= Lx[j] * val
    int i = 0, cend = QP_LDG(pl.Lp + 1);
    double val = WBP(0);
    double nx[8];
    int nr[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { nx[u] = (u < nnzL) ? W(pl.o_Lx, u) : 0.0; nr[u] = (u < nnzL) ? QP_LDG(pl.Li + u) : 0; }
    for (int j0 = 0; j0 < nnzL; j0 += 8) {
      double lx[8];
      int lr[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { lx[u] = nx[u]; lr[u] = nr[u]; }
#pragma unroll
      for (int u = 0; u < 8; u++) {  // next batch in flight while this one is consumed
        const int j = j0 + 8 + u;
        nx[u] = (j < nnzL) ? W(pl.o_Lx, j) : 0.0;
        nr[u] = (j < nnzL) ? QP_LDG(pl.Li + j) : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int j = j0 + u;
        if (j < nnzL) {
          while (j >= cend) { i++; cend = QP_LDG(pl.Lp + i + 1); val = WBP(i); }
          WBP(lr[u]) = WBP(lr[u]) - lx[u] * val;
        }
      }
    }
  }
  for (int i0 = 0; i0 < N; i0 += 8) {
    double dv[8];
#pragma unroll
    for (int u = 0; u < 8; u++) dv[u] = (i0 + u < N) ? W(pl.o_Ddinv, i0 + u) : 0.0;
#pragma unroll
    for (int u = 0; u < 8; u++) if (i0 + u < N) WBP(i0 + u) = WBP(i0 + u) * dv[u];
  }
  {  // QDLDL_Ltsolve: for i = N-1..0: val = x[i]; for j in col i: val -= Lx[j] * x[Li[j]]; x[i] = val
    int t = 0, i = N - 1, cend = QP_LDG(pl.LtEnd);
    double val = WBP(N - 1);
    double nx[8];
    int nr[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { nx[u] = (u < nnzL) ? W(pl.o_LxT, u) : 0.0; nr[u] = (u < nnzL) ? QP_LDG(pl.LtR + u) : 0; }
    for (int k0 = 0; k0 < nnzL; k0 += 8) {
      double lx[8];
      int lr[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { lx[u] = nx[u]; lr[u] = nr[u]; }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int k = k0 + 8 + u;
        nx[u] = (k < nnzL) ? W(pl.o_LxT, k) : 0.0;
        nr[u] = (k < nnzL) ? QP_LDG(pl.LtR + k) : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int k = k0 + u;
        if (k < nnzL) {
          while (k >= cend) { WBP(i) = val; t++; i = N - 1 - t; val = WBP(i); cend = QP_LDG(pl.LtEnd + t); }
          val -= lx[u] * WBP(lr[u]);
        }
      }
    }
    WBP(i) = val;
  }
  for (int j = 0; j < N; j++) WXZ(QP_LDG(pl.perm + j)) = WBP(j);
}

// out(m) = A * v(n)
QP_HD void qp_A_mul(const QpPlanDev& pl, double* ws, size_t stride, int b, int o_v, int o_out) {
  for (int i = 0; i < pl.m; i++) W(o_out, i) = 0.0;
  for (int c = 0; c < pl.n; c++) {
    const double vc = W(o_v, c);
    for (int p = QP_LDG(pl.Ap + c); p < QP_LDG(pl.Ap + c + 1); p++) {
      const int r = QP_LDG(pl.Ai + p);
      W(o_out, r) = W(o_out, r) + W(pl.o_Ax, p) * vc;
    }
  }
}
// out(n) = A' * v(m)
QP_HD void qp_At_mul(const QpPlanDev& pl, double* ws, size_t stride, int b, int o_v, int o_out) {
  for (int c = 0; c < pl.n; c++) {
    double acc = 0.0;
    for (int p = QP_LDG(pl.Ap + c); p < QP_LDG(pl.Ap + c + 1); p++) acc += W(pl.o_Ax, p) * W(o_v, QP_LDG(pl.Ai + p));
    W(o_out, c) = acc;
  }
}
// out(n) = P * v(n), P stored as its upper triangle
QP_HD void qp_P_mul(const QpPlanDev& pl, double* ws, size_t stride, int b, int o_v, int o_out) {
  for (int i = 0; i < pl.n; i++) W(o_out, i) = 0.0;
  for (int c = 0; c < pl.n; c++) {
    const double vc = W(o_v, c);
    for (int p = QP_LDG(pl.Pp + c); p < QP_LDG(pl.Pp + c + 1); p++) {
      const int r = QP_LDG(pl.Pi + p);
      const double a = W(pl.o_Px, p);
      W(o_out, r) = W(o_out, r) + a * vc;
      if (r != c) W(o_out, c) = W(o_out, c) + a * W(o_v, r);
    }
  }
}
QP_HD double qp_norm_inf(double* ws, size_t stride, int b, int o_v, int len) {
  double r = 0.0;
  for (int i = 0; i < len; i++) r = fmax(r, fabs(W(o_v, i)));
  return r;
}
QP_HD double qp_scaled_norm_inf(double* ws, size_t stride, int b, int o_s, int o_v, int len) {
  double r = 0.0;
  for (int i = 0; i < len; i++) r = fmax(r, fabs(W(o_s, i) * W(o_v, i)));
  return r;
}

struct QpResid {
  double prim_res, dual_res, scaled_prim, scaled_dual;
};

// update_info (auxil.c:615-690): residuals of (x, z, y); leaves Ax, Px, A'y in the workspace
QP_HD void qp_update_info(const QpPlanDev& pl, double* ws, size_t stride, int b, double cinv, QpResid& R) {
  const int n = pl.n, m = pl.m;
  qp_A_mul(pl, ws, stride, b, pl.o_x, pl.o_Axv);
  double sp = 0.0, up = 0.0;
  for (int i = 0; i < m; i++) {
    const double d = W(pl.o_Axv, i) - W(pl.o_z, i);
    sp = fmax(sp, fabs(d));
    up = fmax(up, fabs(W(pl.o_Einv, i) * d));
  }
  R.scaled_prim = sp; R.prim_res = up;
  qp_P_mul(pl, ws, stride, b, pl.o_x, pl.o_Pxv);
  qp_At_mul(pl, ws, stride, b, pl.o_y, pl.o_Aty);
  double sd = 0.0, ud = 0.0;
  for (int i = 0; i < n; i++) {
    const double d = (W(pl.o_q, i) + W(pl.o_Pxv, i)) + W(pl.o_Aty, i);
    sd = fmax(sd, fabs(d));
    ud = fmax(ud, fabs(W(pl.o_Dinv, i) * d));
  }
  R.scaled_dual = sd; R.dual_res = cinv * ud;
}

// check_termination (auxil.c:736-851).  Returns the new status or 0.
QP_HD int qp_check_termination(const QpPlanDev& pl, double* ws, size_t stride, int b, const uavmp_osqp_settings& S,
                                    double c, double cinv, const QpResid& R, bool approximate) {
  const int n = pl.n, m = pl.m;
  double eps_abs = S.eps_abs, eps_rel = S.eps_rel, eps_pinf = S.eps_prim_inf, eps_dinf = S.eps_dual_inf;
  if (R.prim_res > OSQP_INFTY_ || R.dual_res > OSQP_INFTY_) return ST_NONCVX;
  if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pinf *= 10; eps_dinf *= 10; }
  bool prim_ok = false, dual_ok = false, pinf = false, dinf = false;
  {
    double mx = fmax(qp_scaled_norm_inf(ws, stride, b, pl.o_Einv, pl.o_z, m),
                     qp_scaled_norm_inf(ws, stride, b, pl.o_Einv, pl.o_Axv, m));
    const double eps_prim = eps_abs + eps_rel * mx;
    if (R.prim_res < eps_prim) {
      prim_ok = true;
    } else {
      // is_primal_infeasible (auxil.c:399-448); bounds are finite here so the polar-cone projection is the identity
      for (int i = 0; i < m; i++) {
        const double l = W(pl.o_l, i), u = W(pl.o_u, i);
        double dy = W(pl.o_dy, i);
        if (u > OSQP_INFTY_ * OSQP_MIN_SCALING_) {
          if (l < -OSQP_INFTY_ * OSQP_MIN_SCALING_) dy = 0.0; else dy = fmin(dy, 0.0);
        } else if (l < -OSQP_INFTY_ * OSQP_MIN_SCALING_) {
          dy = fmax(dy, 0.0);
        }
        W(pl.o_dy, i) = dy;
      }
      const double norm_dy = qp_scaled_norm_inf(ws, stride, b, pl.o_E, pl.o_dy, m);
      if (norm_dy > OSQP_DIVISION_TOL_) {
        double lhs = 0.0, lhs2 = 0.0;
        for (int i = 0; i < m; i++) { const double dy = W(pl.o_dy, i); lhs += W(pl.o_u, i) * fmax(dy, 0.0); }
        for (int i = 0; i < m; i++) { const double dy = W(pl.o_dy, i); lhs2 += W(pl.o_l, i) * fmin(dy, 0.0); }
        lhs += lhs2;
        if (lhs < 0.0) {
          qp_At_mul(pl, ws, stride, b, pl.o_dy, pl.o_tn);
          pinf = qp_scaled_norm_inf(ws, stride, b, pl.o_Dinv, pl.o_tn, n) < eps_pinf * norm_dy;
        }
      }
    }
  }
  {
    double mx = fmax(fmax(qp_scaled_norm_inf(ws, stride, b, pl.o_Dinv, pl.o_q, n),
                          qp_scaled_norm_inf(ws, stride, b, pl.o_Dinv, pl.o_Aty, n)),
                     qp_scaled_norm_inf(ws, stride, b, pl.o_Dinv, pl.o_Pxv, n));
    mx *= cinv;
    const double eps_dual = eps_abs + eps_rel * mx;
    if (R.dual_res < eps_dual) {
      dual_ok = true;
    } else {
      // is_dual_infeasible (auxil.c:450-528)
      const double norm_dx = qp_scaled_norm_inf(ws, stride, b, pl.o_D, pl.o_dx, n);
      if (norm_dx > OSQP_DIVISION_TOL_) {
        double qdx = 0.0;
        for (int i = 0; i < n; i++) qdx += W(pl.o_q, i) * W(pl.o_dx, i);
        if (qdx < 0.0) {
          qp_P_mul(pl, ws, stride, b, pl.o_dx, pl.o_tn);
          if (qp_scaled_norm_inf(ws, stride, b, pl.o_Dinv, pl.o_tn, n) < c * eps_dinf * norm_dx) {
            qp_A_mul(pl, ws, stride, b, pl.o_dx, pl.o_tm);
            bool in_cone = true;
            const double tol = eps_dinf * norm_dx;
            for (int i = 0; i < m; i++) {
              const double v = W(pl.o_Einv, i) * W(pl.o_tm, i);
              if ((W(pl.o_u, i) < OSQP_INFTY_ * OSQP_MIN_SCALING_ && v > tol) ||
                  (W(pl.o_l, i) > -OSQP_INFTY_ * OSQP_MIN_SCALING_ && v < -tol)) { in_cone = false; break; }
            }
            dinf = in_cone;
          }
        }
      }
    }
  }
  if (prim_ok && dual_ok) return approximate ? ST_SOLVED_INACC : ST_SOLVED;
  if (pinf) return approximate ? ST_PINF_INACC : ST_PINF;
  if (dinf) return approximate ? ST_DINF_INACC : ST_DINF;
  return 0;
}


// one problem: assembly -> osqp_setup -> osqp_solve -> store_solution
QP_HD void qp_solve_one(const QpPlanDev& pl, const QpIo& io, const uavmp_osqp_settings& S, double* ws, int b, double* sv) {
  const size_t stride = (size_t)io.stride;
  const int n = pl.n, m = pl.m, Sg = pl.S, nc = pl.nc;

  // ---- assembly: P, q, A, l, u (minimum_control.cpp:5-125) ---------------------------------------------------
  const double* T = io.T + (size_t)b * Sg;
  for (int p = 0; p < pl.nnzP; p++) W(pl.o_Px, p) = QP_LDG(pl.P_coef + p) * fpm::powi(T[QP_LDG(pl.P_seg + p)], QP_LDG(pl.P_pow + p));
  for (int p = 0; p < pl.nnzA; p++) W(pl.o_Ax, p) = QP_LDG(pl.A_coef + p) * fpm::powi(T[QP_LDG(pl.A_seg + p)], QP_LDG(pl.A_pow + p));
  for (int i = 0; i < n; i++) W(pl.o_q, i) = 0.0;
  for (int i = 0; i < m; i++) {  // equality rows: l == u (minimum_control.cpp:98-125); corridor rows (extension): lo <= . <= hi
    W(pl.o_l, i) = qp_bound_value(io, b, Sg, QP_LDG(pl.l_src + i));
    W(pl.o_u, i) = qp_bound_value(io, b, Sg, QP_LDG(pl.u_src + i));
  }

  // ---- scale_data (scaling.c:49-165) ----------------------------------------------------------------------------
  double c = 1.0;
  for (int i = 0; i < n; i++) W(pl.o_D, i) = 1.0;
  for (int i = 0; i < m; i++) W(pl.o_E, i) = 1.0;
  for (int it = 0; it < S.scaling; it++) {
    // column inf-norms of [P; A] — P's stored upper triangle only (csc_col_norm_inf ignores symmetry) — and rows of A
    for (int j = 0; j < n; j++) {
      double dn = 0.0;
      for (int p = QP_LDG(pl.Pp + j); p < QP_LDG(pl.Pp + j + 1); p++) dn = fmax(fabs(W(pl.o_Px, p)), dn);
      double an = 0.0;
      for (int p = QP_LDG(pl.Ap + j); p < QP_LDG(pl.Ap + j + 1); p++) an = fmax(fabs(W(pl.o_Ax, p)), an);
      W(pl.o_tn, j) = fmax(dn, an);
    }
    for (int i = 0; i < m; i++) W(pl.o_tm, i) = 0.0;
    for (int p = 0; p < pl.nnzA; p++) {
      const int r = QP_LDG(pl.Ai + p);
      W(pl.o_tm, r) = fmax(fabs(W(pl.o_Ax, p)), W(pl.o_tm, r));
    }
    for (int j = 0; j < n; j++) W(pl.o_tn, j) = 1.0 / sqrt(limit_scaling(W(pl.o_tn, j)));
    for (int i = 0; i < m; i++) W(pl.o_tm, i) = 1.0 / sqrt(limit_scaling(W(pl.o_tm, i)));
    // P <- D P D, A <- E A D, q <- D q
    for (int j = 0; j < n; j++) {
      const double dj = W(pl.o_tn, j);
      for (int p = QP_LDG(pl.Pp + j); p < QP_LDG(pl.Pp + j + 1); p++) {
        double v = W(pl.o_Px, p) * W(pl.o_tn, QP_LDG(pl.Pi + p));
        W(pl.o_Px, p) = v * dj;
      }
      for (int p = QP_LDG(pl.Ap + j); p < QP_LDG(pl.Ap + j + 1); p++) {
        double v = W(pl.o_Ax, p) * W(pl.o_tm, QP_LDG(pl.Ai + p));
        W(pl.o_Ax, p) = v * dj;
      }
      W(pl.o_q, j) = W(pl.o_q, j) * dj;
      W(pl.o_D, j) = W(pl.o_D, j) * dj;
    }
    for (int i = 0; i < m; i++) W(pl.o_E, i) = W(pl.o_E, i) * W(pl.o_tm, i);
    // cost normalisation
    double sum = 0.0;
    for (int j = 0; j < n; j++) {
      double dn = 0.0;
      for (int p = QP_LDG(pl.Pp + j); p < QP_LDG(pl.Pp + j + 1); p++) dn = fmax(fabs(W(pl.o_Px, p)), dn);
      sum += fabs(dn);
    }
    double c_temp = sum / n;
    double inf_q = limit_scaling(qp_norm_inf(ws, stride, b, pl.o_q, n));
    c_temp = fmax(c_temp, inf_q);
    c_temp = limit_scaling(c_temp);
    c_temp = 1.0 / c_temp;
    for (int p = 0; p < pl.nnzP; p++) W(pl.o_Px, p) = W(pl.o_Px, p) * c_temp;
    for (int j = 0; j < n; j++) W(pl.o_q, j) = W(pl.o_q, j) * c_temp;
    c *= c_temp;
  }
  const double cinv = 1.0 / c;
  for (int j = 0; j < n; j++) W(pl.o_Dinv, j) = 1.0 / W(pl.o_D, j);
  for (int i = 0; i < m; i++) {
    const double e = W(pl.o_E, i);
    W(pl.o_Einv, i) = 1.0 / e;
    W(pl.o_l, i) = W(pl.o_l, i) * e;
    W(pl.o_u, i) = W(pl.o_u, i) * e;
  }

  // ---- set_rho_vec (auxil.c:75-104) ----------------------------------------------------------------------------------
  double rho = fmin(fmax(S.rho, OSQP_RHO_MIN_), OSQP_RHO_MAX_);
  auto set_rho = [&](double rho_) {
    for (int i = 0; i < m; i++) {
      const double l = W(pl.o_l, i), u = W(pl.o_u, i);
      double r;
      if (l < -OSQP_INFTY_ * OSQP_MIN_SCALING_ && u > OSQP_INFTY_ * OSQP_MIN_SCALING_) r = OSQP_RHO_MIN_;
      else if (u - l < OSQP_RHO_TOL_) r = OSQP_RHO_EQ_OVER_RHO_INEQ_ * rho_;
      else r = rho_;
      W(pl.o_rho, i) = r;
      W(pl.o_rhoinv, i) = 1.0 / r;
    }
  };
  set_rho(rho);

  int status = ST_UNSOLVED, iter_out = 0;
  // ---- KKT factorisation (init_linsys_solver_qdldl) -----------------------------------------------------------------------
  if (qp_factor(pl, ws, stride, b, S.sigma, sv) < n) status = ST_NONCVX;  // osqp_setup fails: OSQP_NONCVX_ERROR

  if (status == ST_UNSOLVED) {
    for (int i = 0; i < n; i++) { W(pl.o_x, i) = 0.0; W(pl.o_xprev, i) = 0.0; }
    for (int i = 0; i < m; i++) { W(pl.o_z, i) = 0.0; W(pl.o_zprev, i) = 0.0; W(pl.o_y, i) = 0.0; }
    const int interval = S.adaptive_rho_interval ? S.adaptive_rho_interval
                                                 : (S.check_termination ? 4 * S.check_termination : 100);
    const double alpha = S.alpha, sigma = S.sigma, one_m_alpha = 1.0 - S.alpha;
    QpResid R;
    R.prim_res = R.dual_res = R.scaled_prim = R.scaled_dual = OSQP_INFTY_;
    bool checked_last = false;
    int iter;
    for (iter = 1; iter <= S.max_iter; iter++) {
      // x_prev <- x, z_prev <- z (the reference swaps pointers; x and z are fully overwritten below)
      // compute_rhs (auxil.c:135-157).  Loops are written in batches of 4 with all loads ahead of the stores: the values
      // and their order are unchanged, a thread merely keeps several independent loads in flight.
      for (int i0 = 0; i0 < n; i0 += 4) {
        double xv[4], qv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (i0 + u < n) { xv[u] = W(pl.o_x, i0 + u); qv[u] = W(pl.o_q, i0 + u); }
#pragma unroll
        for (int u = 0; u < 4; u++) if (i0 + u < n) { W(pl.o_xprev, i0 + u) = xv[u]; WXZ(i0 + u) = sigma * xv[u] + (-1.0) * qv[u]; }
      }
      for (int i0 = 0; i0 < m; i0 += 4) {
        double zv[4], rv[4], yv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (i0 + u < m) { zv[u] = W(pl.o_z, i0 + u); rv[u] = W(pl.o_rhoinv, i0 + u); yv[u] = W(pl.o_y, i0 + u); }
#pragma unroll
        for (int u = 0; u < 4; u++) if (i0 + u < m) {
          W(pl.o_zprev, i0 + u) = zv[u];
          const double t = rv[u] * yv[u];
          const double rhs = (-1.0) * t + 1.0 * zv[u];
          WXZ(n + i0 + u) = rhs;
          W(pl.o_tm, i0 + u) = rhs;  // keep the right-hand side of the z block (qdldl_interface.c:447-450)
        }
      }
      qp_kkt_solve(pl, ws, stride, b, sv);
      // update_x (auxil.c:171-186)
      for (int i0 = 0; i0 < n; i0 += 4) {
        double xp[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (i0 + u < n) xp[u] = W(pl.o_xprev, i0 + u);
#pragma unroll
        for (int u = 0; u < 4; u++) if (i0 + u < n) {
          const double xn = alpha * WXZ(i0 + u) + one_m_alpha * xp[u];
          W(pl.o_x, i0 + u) = xn;
          W(pl.o_dx, i0 + u) = xn - xp[u];
        }
      }
      // ztilde = rhs_z + rho^-1 * nu, then update_z / update_y (auxil.c:188-228)
      for (int i0 = 0; i0 < m; i0 += 2) {
        double tmv[2], rv[2], zp[2], yv[2], lv[2], uv[2], rh[2];
#pragma unroll
        for (int u = 0; u < 2; u++) if (i0 + u < m) {
          tmv[u] = W(pl.o_tm, i0 + u); rv[u] = W(pl.o_rhoinv, i0 + u); zp[u] = W(pl.o_zprev, i0 + u); yv[u] = W(pl.o_y, i0 + u);
          lv[u] = W(pl.o_l, i0 + u); uv[u] = W(pl.o_u, i0 + u); rh[u] = W(pl.o_rho, i0 + u);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) if (i0 + u < m) {
          const double zt = tmv[u] + rv[u] * WXZ(n + i0 + u);
          double zn = rv[u] * yv[u];
          // OSQPVectorf_add_scaled3 with x == a and sca == 1 takes its incrementing form: z += (alpha ztilde + (1 - alpha) z_prev)
          // (algebra/builtin/vector.c:441-444); the association only shows on inequality rows (equality rows are clipped to l == u)
          zn = zn + (alpha * zt + one_m_alpha * zp[u]);
          zn = fmin(fmax(zn, lv[u]), uv[u]);
          W(pl.o_z, i0 + u) = zn;
          double dy = (alpha * zt + one_m_alpha * zp[u]) + (-1.0) * zn;
          dy = dy * rh[u];
          W(pl.o_dy, i0 + u) = dy;
          W(pl.o_y, i0 + u) = yv[u] + dy;
        }
      }
      const bool can_check = S.check_termination && (iter % S.check_termination == 0);
      checked_last = can_check;
      if (can_check) {
        qp_update_info(pl, ws, stride, b, cinv, R);
        iter_out = iter;
        int st = qp_check_termination(pl, ws, stride, b, S, c, cinv, R, false);
        if (st) { status = st; break; }
      }
      if (S.adaptive_rho && interval && (iter % interval == 0)) {
        if (!can_check) { qp_update_info(pl, ws, stride, b, cinv, R); iter_out = iter; }
        // compute_rho_estimate + adapt_rho (auxil.c:14-73)
        double pr = R.scaled_prim, dr = R.scaled_dual;
        double pn = fmax(qp_norm_inf(ws, stride, b, pl.o_z, m), qp_norm_inf(ws, stride, b, pl.o_Axv, m));
        pr /= (pn + OSQP_DIVISION_TOL_);
        double dn = fmax(fmax(qp_norm_inf(ws, stride, b, pl.o_q, n), qp_norm_inf(ws, stride, b, pl.o_Aty, n)),
                         qp_norm_inf(ws, stride, b, pl.o_Pxv, n));
        dr /= (dn + OSQP_DIVISION_TOL_);
        double rho_new = rho * sqrt(pr / dr);
        rho_new = fmin(fmax(rho_new, OSQP_RHO_MIN_), OSQP_RHO_MAX_);
        if (rho_new > rho * S.adaptive_rho_tolerance || rho_new < rho / S.adaptive_rho_tolerance) {
          rho = fmin(fmax(rho_new, OSQP_RHO_MIN_), OSQP_RHO_MAX_);  // osqp_update_rho (osqp_api.c:1178-1228)
          for (int i = 0; i < m; i++) {
            // constraint classes were fixed at setup (work->constr_type)
            const double l = W(pl.o_l, i), u = W(pl.o_u, i);
            double r;
            if (l < -OSQP_INFTY_ * OSQP_MIN_SCALING_ && u > OSQP_INFTY_ * OSQP_MIN_SCALING_) r = OSQP_RHO_MIN_;
            else if (u - l < OSQP_RHO_TOL_) r = OSQP_RHO_EQ_OVER_RHO_INEQ_ * rho;
            else r = rho;
            W(pl.o_rho, i) = r;
            W(pl.o_rhoinv, i) = 1.0 / r;
          }
          if (qp_factor(pl, ws, stride, b, sigma, sv) < 0) { status = ST_NONCVX; break; }
        }
      }
    }
    if (status == ST_UNSOLVED) {
      // osqp_api.c:686-731: loop ran out
      if (!checked_last) {
        qp_update_info(pl, ws, stride, b, cinv, R);
        iter_out = iter - 1;
        int st = qp_check_termination(pl, ws, stride, b, S, c, cinv, R, false);
        if (st) status = st;
      }
      if (status == ST_UNSOLVED) {
        int st = qp_check_termination(pl, ws, stride, b, S, c, cinv, R, true);
        status = st ? st : ST_MAXITER;
      }
    }
  }

  // ---- store_solution (auxil.c:537-613): x = D x_scaled, NaN when there is no solution -------------------------------
  const bool has_sol = !(status == ST_PINF || status == ST_PINF_INACC || status == ST_DINF || status == ST_DINF_INACC ||
                         status == ST_NONCVX);
  double* out = io.coef + (size_t)b * n;
  for (int i = 0; i < n; i++) out[i] = has_sol ? W(pl.o_D, i) * W(pl.o_x, i) : fpm::from_bits(0x7ff8000000000000ull);
  io.status[b] = status;
  io.iters[b] = iter_out;
  io.solved[b] = (status == ST_SOLVED) ? 1 : 0;  // OsqpEigen::Solver::solve is true only for OSQP_SOLVED
  (void)nc;
}

#undef W
#undef WBP
#undef WXZ
#undef WYW
