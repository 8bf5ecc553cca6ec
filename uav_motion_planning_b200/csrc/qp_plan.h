// qp_plan.h — pattern-only data of one (order, S) QP family, built by qp_symbolic.cpp, consumed by qp_kernel.cu.
#pragma once
#include <stddef.h>

#include <vector>

struct QpPlanHost {
  int order, S, k, nc, n, m, N, nnzP, nnzA, nnzK, nnzL;
  std::vector<int> Pp, Pi, P_seg, P_pow;  std::vector<double> P_coef;   // upper-triangular CSC of P + value recipe
  std::vector<int> Ap, Ai, A_seg, A_pow;  std::vector<double> A_coef;   // CSC of A + value recipe coef * T[seg]^pow
  std::vector<int> l_src;                                               // bound source per constraint row (-1: 0.0)
  std::vector<int> Kp0, Ki0;                                            // unpermuted upper CSC KKT pattern (form_KKT order)
  bool perm_from_table = false;                                         // true: the reference AMD's permutation (tabulated)
  std::vector<int> perm;                                                // perm[j] = original KKT index at position j
  std::vector<int> Kp, Ki, Kkind, Kidx;                                 // permuted upper CSC of the KKT matrix
  std::vector<int> Lp, Li;                                              // pattern of L (strictly lower, CSC)
  std::vector<int> Rp, Rc, Rpos;                                        // per row k: reach columns and target slots
  std::vector<int> Arp, Arj, Arx;                                       // CSR view of A: per row, column index and slot in Ax, by increasing column
  std::vector<int> Psp, Psa, Psv;                                       // per element i of P*v: (slot in Px, index into v) in the CSC loop's order
  std::vector<int> LevP, LevC; int nlev = 0;                            // L' solve: columns with entries grouped by elimination-tree level
  std::vector<int> Ltpos, LtR, LtEnd;                                   // L in the L' solve's consumption order: slot of L entry j, row index per slot, end slot per processed column
};

QpPlanHost* qp_plan_build(int order, int S);

// device view (all int arrays live in one allocation)
struct QpPlanDev {
  int order, S, k, nc, n, m, N, nnzP, nnzA, nnzK, nnzL;
  const int *Pp, *Pi, *P_seg, *P_pow;
  const int *Ap, *Ai, *A_seg, *A_pow;
  const double *P_coef, *A_coef;
  const int *l_src, *perm, *Kp, *Ki, *Kkind, *Kidx, *Lp, *Li, *Rp, *Rc, *Rpos, *Ltpos, *LtR, *LtEnd;
  const int *Arp, *Arj, *Arx, *Psp, *Psa, *Psv, *LevP, *LevC;
  int nlev, ws_warp;  // ws_warp: doubles of the one-warp-per-problem workspace (everything except the LxT copy)
  // workspace layout (offsets in doubles)
  int o_Px, o_Ax, o_q, o_l, o_u, o_D, o_Dinv, o_E, o_Einv, o_rho, o_rhoinv, o_Lx, o_LxT, o_Dd, o_Ddinv, o_yw, o_x, o_xprev,
      o_dx, o_Pxv, o_Aty, o_z, o_zprev, o_y, o_dy, o_Axv, o_xz, o_bp, o_tn, o_tm, ws_doubles;
};

// flattening of a host plan into one int array + one double array (what is uploaded), and the view over it
struct QpPlanOffsets { size_t Pp, Pi, P_seg, P_pow, Ap, Ai, A_seg, A_pow, l_src, perm, Kp, Ki, Kkind, Kidx, Lp, Li, Rp, Rc, Rpos, Ltpos, LtR, LtEnd, Arp, Arj, Arx, Psp, Psa, Psv, LevP, LevC, A_coef; };
void qp_plan_pack(const QpPlanHost& H, std::vector<int>& ints, std::vector<double>& dbls, QpPlanOffsets& off);
void qp_plan_bind(const QpPlanHost& H, const QpPlanOffsets& off, const int* ints, const double* dbls, QpPlanDev& D);

// per-batch I/O of the solve kernel
struct QpIo {
  const double* pos; const double* bv; const double* ba; const double* bj; const double* T;
  double* coef; int* solved; int* status; int* iters;
  int B, stride;
};
