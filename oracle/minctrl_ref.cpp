// oracle/minctrl_ref.cpp — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py CPU legs).
//
// CPU restatement of traj_optimization::MinimumControl (reference:
//   src/planner/traj_optimization/src/minimum_control.cpp:5-19 getHessian, :26-96 getConstraintMatrix,
//   :98-125 getBound, :127-192 solve) and of the OsqpEigen marshalling it goes through
//   (3rd/osqp-eigen/include/OsqpEigen/Data.tpp:42 upper-triangular P, SparseMatrixHelper.tpp:14-71 CSC copy that
//   keeps explicit zeros), feeding the reference's OWN vendored OSQP C code, compiled unmodified into
//   oracle/_ref/libosqp_ref.so (plus the restated QDLDL in oracle/qdldl, which the reference fetches at build time).
// order 5 (minimum jerk) is what the reference implements; order 7 (minimum snap) is the generalisation of
// SURVEY.md §9.3, labelled "extension" wherever it is reported.
// PARITY STATUS: OSQP + QDLDL are pinned by the reference's known-answer tests (tests/test_osqp_kat.py);
// MinimumControl's own matrices/coefficients are unpinned by the reference (it records no expected outputs) and
// are pinned by goldens this oracle generated (tests/golden/minctrl_*.json).
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../uav_motion_planning_b200/csrc/fpmath.h"
#include "oracle.h"
#include "oracle_osqp.h"

namespace {

void* g_ref = nullptr;
osqp_ref_solve_fn g_solve = nullptr;
osqp_ref_kkt_solve_fn g_kkt = nullptr;

bool load_ref() {
  if (g_solve) return true;
  Dl_info di;
  std::string dir = ".";
  if (dladdr((void*)&load_ref, &di) && di.dli_fname) {
    std::string p(di.dli_fname);
    size_t k = p.find_last_of('/');
    dir = (k == std::string::npos) ? "." : p.substr(0, k);
  }
  std::string path = dir + "/_ref/libosqp_ref.so";
  g_ref = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!g_ref) { fprintf(stderr, "oracle: cannot load %s: %s\n", path.c_str(), dlerror()); return false; }
  g_solve = (osqp_ref_solve_fn)dlsym(g_ref, "osqp_ref_solve");
  g_kkt = (osqp_ref_kkt_solve_fn)dlsym(g_ref, "osqp_ref_kkt_solve");
  return g_solve != nullptr;
}

// Eigen::SparseMatrix::insert keeps each column sorted by row; entries are unique -> a (col,row)-ordered map
struct Triplets {
  std::map<std::pair<long long, long long>, double> e;  // key (col, row)
  void insert(long long r, long long c, double v) { e[{c, r}] = v; }
  void toCSC(long long ncol, std::vector<long long>& p, std::vector<long long>& i, std::vector<double>& x,
             bool upper_only) const {
    p.assign(ncol + 1, 0);
    i.clear(); x.clear();
    for (auto& kv : e) {
      long long c = kv.first.first, r = kv.first.second;
      if (upper_only && r > c) continue;  // P.triangularView<Eigen::Upper>() (Data.tpp:42)
      p[c + 1]++;
      i.push_back(r);
      x.push_back(kv.second);
    }
    for (long long c = 0; c < ncol; c++) p[c + 1] += p[c];
  }
};

double mpow(int libm, double t, int n) { return libm ? std::pow(t, n) : fpm::powi(t, n); }
double falling(int j, int r) {  // j (j-1) ... (j-r+1)
  double v = 1.0;
  for (int k = 0; k < r; k++) v *= (double)(j - k);
  return v;
}

}  // namespace

extern "C" {

// Assemble P (upper CSC), q, A (CSC with the reference's explicit zeros), l, u.  d = order, k = (d+1)/2.
// Returns n (variables); *m_out constraints.  Buffers sized by the caller (see oracle_minctrl_dims).
// Kc > 0 (extension, SURVEY.md §9.3 "Corridor"): Kc inequality rows per segment, lo <= p_s(phi_j T_s) <= hi at
// phi_j = (j + 1) / (Kc + 1); row index m_eq + s Kc + j; the entry for coefficient i is phi_j^i * T_s^i.
void oracle_minctrl_dims_c(int order, int S, int Kc, int* n, int* m, int* nnzP, int* nnzA) {
  int k = (order + 1) / 2, nc = order + 1;
  *n = nc * S;
  *m = 2 * k + (k + 1) * (S - 1) + Kc * S;
  *nnzP = S * (k * (k + 1)) / 2;
  // start k ; interior: waypoint nc + sum_r (nc + r + 1) ; end k*nc
  int per = nc;
  for (int r = 0; r < k; r++) per += nc + r + 1;
  *nnzA = k + (S - 1) * per + k * nc + Kc * S * nc;
}
void oracle_minctrl_dims(int order, int S, int* n, int* m, int* nnzP, int* nnzA) { oracle_minctrl_dims_c(order, S, 0, n, m, nnzP, nnzA); }

int oracle_minctrl_assemble_c(int order, int S, int Kc, const double* pos_1d, const double* bound_vel, const double* bound_acc,
                              const double* bound_jerk, const double* T, const double* corridor_lo, const double* corridor_hi,
                              int libm_mode, long long* Pp, long long* Pi, double* Px, double* q, long long* Ap, long long* Ai,
                              double* Ax, double* l, double* u) {
  const int k = (order + 1) / 2, nc = order + 1;
  const int n = nc * S, m_eq = 2 * k + (k + 1) * (S - 1), m = m_eq + Kc * S;
  Triplets P, A;
  // getHessian (minimum_control.cpp:5-19): full symmetric block inserted; OsqpEigen keeps the upper triangle
  for (int s = 0; s < S; s++)
    for (int i = k; i < nc; i++)
      for (int j = k; j < nc; j++) {
        int e = i + j - 2 * k + 1;
        double coef = falling(i, k) * falling(j, k) / (double)e;  // 36, 72, 120, 192, 360, 720 for k = 3
        P.insert(nc * s + i, nc * s + j, coef * mpow(libm_mode, T[s], e));
      }
  for (int i = 0; i < n; i++) q[i] = 0.0;  // getGradient :21-24
  // getConstraintMatrix :26-96
  for (int r = 0; r < k; r++) A.insert(r, r, falling(r, r));  // start p, v, a(, j): A(2,2) = 2.0
  auto deriv_row = [&](long long row, int s, int r, double t) {
    // d^r/dt^r of sum_j c_j t^j on segment s; the reference also stores the zeros for j < r (:55,61,64,...)
    for (int j = 0; j < nc; j++) {
      double v = (j < r) ? 0.0 : falling(j, r) * mpow(libm_mode, t, j - r);
      A.insert(row, nc * s + j, v);
    }
  };
  for (int s = 0; s + 1 < S; s++) {
    long long base = k + (long long)(k + 1) * s;
    deriv_row(base, s, 0, T[s]);  // waypoint row :36-44
    for (int r = 0; r < k; r++) {  // continuity rows :47-77
      long long row = base + 1 + r;
      deriv_row(row, s, r, T[s]);
      for (int j = 0; j < r; j++) A.insert(row, nc * (s + 1) + j, 0.0);
      A.insert(row, nc * (s + 1) + r, -falling(r, r));
    }
  }
  {
    long long base = k + (long long)(k + 1) * (S - 1);
    for (int r = 0; r < k; r++) deriv_row(base + r, S - 1, r, T[S - 1]);  // end p, v, a(, j) :80-95
  }
  // getBound :98-125
  for (int i = 0; i < m; i++) { l[i] = 0.0; u[i] = 0.0; }
  const double* bs[4] = {nullptr, bound_vel, bound_acc, bound_jerk};
  l[0] = u[0] = pos_1d[0];
  for (int r = 1; r < k; r++) l[r] = u[r] = bs[r][0];
  long long eb = k + (long long)(k + 1) * (S - 1);
  l[eb] = u[eb] = pos_1d[S];
  for (int r = 1; r < k; r++) l[eb + r] = u[eb + r] = bs[r][1];
  for (int s = 0; s + 1 < S; s++) {
    long long row = k + (long long)(k + 1) * s;
    l[row] = u[row] = pos_1d[s + 1];
  }
  // corridor rows (extension): lo <= sum_i c_i (phi_j^i) (T_s^i) <= hi
  for (int s = 0; s < S; s++)
    for (int j = 0; j < Kc; j++) {
      const double phi = (double)(j + 1) / (double)(Kc + 1);
      const long long row = m_eq + (long long)s * Kc + j;
      for (int i = 0; i < nc; i++) A.insert(row, nc * s + i, fpm::powi(phi, i) * mpow(libm_mode, T[s], i));
      l[row] = corridor_lo[s];
      u[row] = corridor_hi[s];
    }
  std::vector<long long> p, i;
  std::vector<double> x;
  P.toCSC(n, p, i, x, true);
  memcpy(Pp, p.data(), p.size() * 8); memcpy(Pi, i.data(), i.size() * 8); memcpy(Px, x.data(), x.size() * 8);
  A.toCSC(n, p, i, x, false);
  memcpy(Ap, p.data(), p.size() * 8); memcpy(Ai, i.data(), i.size() * 8); memcpy(Ax, x.data(), x.size() * 8);
  return 0;
}

int oracle_minctrl_assemble(int order, int S, const double* pos_1d, const double* bound_vel, const double* bound_acc,
                            const double* bound_jerk, const double* T, int libm_mode, long long* Pp, long long* Pi,
                            double* Px, double* q, long long* Ap, long long* Ai, double* Ax, double* l, double* u) {
  return oracle_minctrl_assemble_c(order, S, 0, pos_1d, bound_vel, bound_acc, bound_jerk, T, nullptr, nullptr, libm_mode, Pp, Pi, Px, q,
                                   Ap, Ai, Ax, l, u);
}

// MinimumControl::solve (+ getCoef1d): returns 1 iff OSQP status == SOLVED (Solver.cpp:181-187), else 0; -1 if the
// reference library is unavailable.
int oracle_minctrl_solve_c(int order, int S, int Kc, const double* pos_1d, const double* bound_vel, const double* bound_acc,
                           const double* bound_jerk, const double* T, const double* corridor_lo, const double* corridor_hi,
                           const oracle_osqp_settings* st, int libm_mode, double* coef, oracle_osqp_info* info) {
  if (!load_ref()) return -1;
  int n, m, nnzP, nnzA;
  oracle_minctrl_dims_c(order, S, Kc, &n, &m, &nnzP, &nnzA);
  std::vector<long long> Pp(n + 1), Pi(nnzP), Ap(n + 1), Ai(nnzA);
  std::vector<double> Px(nnzP), q(n), Ax(nnzA), l(m), u(m), y(m);
  oracle_minctrl_assemble_c(order, S, Kc, pos_1d, bound_vel, bound_acc, bound_jerk, T, corridor_lo, corridor_hi, libm_mode, Pp.data(),
                            Pi.data(), Px.data(), q.data(), Ap.data(), Ai.data(), Ax.data(), l.data(), u.data());
  int rc = g_solve(n, m, Pp.data(), Pi.data(), Px.data(), q.data(), Ap.data(), Ai.data(), Ax.data(), l.data(),
                   u.data(), st, coef, y.data(), info);
  if (rc) return 0;  // "solver init failed!" (:173-177)
  return info->status_val == 1 ? 1 : 0;
}

int oracle_minctrl_solve(int order, int S, const double* pos_1d, const double* bound_vel, const double* bound_acc,
                         const double* bound_jerk, const double* T, const oracle_osqp_settings* st, int libm_mode,
                         double* coef, oracle_osqp_info* info) {
  return oracle_minctrl_solve_c(order, S, 0, pos_1d, bound_vel, bound_acc, bound_jerk, T, nullptr, nullptr, st, libm_mode, coef, info);
}

// generic QP through the reference OSQP (known-answer tests)
int oracle_osqp_solve(long long n, long long m, const long long* Pp, const long long* Pi, const double* Px,
                      const double* q, const long long* Ap, const long long* Ai, const double* Ax, const double* l,
                      const double* u, const oracle_osqp_settings* st, double* x, double* y, oracle_osqp_info* info) {
  if (!load_ref()) return -1;
  return g_solve(n, m, Pp, Pi, Px, q, Ap, Ai, Ax, l, u, st, x, y, info);
}
int oracle_osqp_kkt_solve(long long n, long long m, const long long* Pp, const long long* Pi, const double* Px,
                          const long long* Ap, const long long* Ai, const double* Ax, double sigma, double rho,
                          const double* rhs, double* sol) {
  if (!load_ref() || !g_kkt) return -1;
  return g_kkt(n, m, Pp, Pi, Px, Ap, Ai, Ax, sigma, rho, rhs, sol);
}
int oracle_have_ref(void) { return load_ref() ? 1 : 0; }

}  // extern "C"
