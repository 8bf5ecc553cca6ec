// qp_plan.h — pattern-only data of one (order, S) QP family, built by qp_symbolic.cpp, consumed by qp_kernel.cu.
#pragma once
#include <stddef.h>

#include <vector>

struct QpPlanHost {
  int order, S, k, nc, n, m, N, nnzP, nnzA, nnzK, nnzL;
  int Kc = 0, m_eq = 0;                                                 // corridor samples per segment; rows [m_eq, m) are the corridor (inequality) rows
  std::vector<int> Pp, Pi, P_seg, P_pow;  std::vector<double> P_coef;   // upper-triangular CSC of P + value recipe
  std::vector<int> Ap, Ai, A_seg, A_pow;  std::vector<double> A_coef;   // CSC of A + value recipe coef * T[seg]^pow
  std::vector<int> l_src, u_src;                                        // bound sources per constraint row (-1: 0.0), see qp_bound_value
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
  // warp-per-problem triangular solves: L by rows (entries of a row by increasing column = the order the column-oriented
  // QDLDL_Lsolve subtracts them from that row), rows grouped by dependency level; all of it packed as uint16 in `Sidx`
  std::vector<int> Lrp, Lrc, Lrx, LevFP, LevFR; int nlevf = 0;
  // lane schedule of the triangular solves (see qw_tri_sched): per chunk 32 records {lx slot:16 | source index:16, row:16 | len:8 | maxlen:8};
  // chunks [0, sch_nf) = forward solve, [sch_nf, sch_n) = backward solve; empty when a row has more than 32 entries
  std::vector<unsigned int> Sch; int sch_nf = 0, sch_n = 0;
  std::vector<unsigned short> Sidx; int sx_Lrp = 0, sx_Lrc = 0, sx_Lrx = 0, sx_LevFP = 0, sx_LevFR = 0, sx_Lp = 0, sx_Li = 0, sx_LevP = 0, sx_LevC = 0, sx_perm = 0;
};

// Kc > 0 appends Kc corridor rows per segment (an extension): lo <= p_s(phi_j T_s) <= hi at phi_j = (j + 1) / (Kc + 1)
QpPlanHost* qp_plan_build(int order, int S, int Kc = 0);

// device view (all int arrays live in one allocation)
struct QpPlanDev {
  int order, S, k, nc, n, m, N, nnzP, nnzA, nnzK, nnzL, Kc, m_eq;
  const int *Pp, *Pi, *P_seg, *P_pow;
  const int *Ap, *Ai, *A_seg, *A_pow;
  const double *P_coef, *A_coef;
  const int *l_src, *u_src, *perm, *Kp, *Ki, *Kkind, *Kidx, *Lp, *Li, *Rp, *Rc, *Rpos, *Ltpos, *LtR, *LtEnd;
  const int *Arp, *Arj, *Arx, *Psp, *Psa, *Psv, *LevP, *LevC;
  int nlev, ws_warp;  // ws_warp: doubles of the one-warp-per-problem workspace (everything except the LxT copy)
  // the index block of the warp-per-problem triangular solves (uint16, `n_sidx` entries; staged in shared memory by the kernels)
  const unsigned int* Sch; int sch_nf, sch_n;  // lane schedule of the triangular solves (global memory, 2 words per lane per chunk)
  const unsigned short* Sidx; int n_sidx, nlevf, sx_Lrp, sx_Lrc, sx_Lrx, sx_LevFP, sx_LevFR, sx_Lp, sx_Li, sx_LevP, sx_LevC, sx_perm;
  // workspace layout (offsets in doubles)
  int o_Px, o_Ax, o_q, o_l, o_u, o_D, o_Dinv, o_E, o_Einv, o_rho, o_rhoinv, o_Lx, o_LxT, o_Dd, o_Ddinv, o_yw, o_x, o_xprev,
      o_dx, o_Pxv, o_Aty, o_z, o_zprev, o_y, o_dy, o_Axv, o_xz, o_bp, o_tn, o_tm, o_pb, ws_doubles;
};

// flattening of a host plan into one int array + one double array (what is uploaded), and the view over it
struct QpPlanOffsets { size_t Pp, Pi, P_seg, P_pow, Ap, Ai, A_seg, A_pow, l_src, u_src, perm, Kp, Ki, Kkind, Kidx, Lp, Li, Rp, Rc, Rpos, Ltpos, LtR, LtEnd, Arp, Arj, Arx, Psp, Psa, Psv, LevP, LevC, A_coef; };
void qp_plan_pack(const QpPlanHost& H, std::vector<int>& ints, std::vector<double>& dbls, QpPlanOffsets& off);
void qp_plan_bind(const QpPlanHost& H, const QpPlanOffsets& off, const int* ints, const double* dbls, QpPlanDev& D);
// D.Sidx / D.Sch are NOT set by qp_plan_bind: point them at copies of H.Sidx / H.Sch in the memory space the solver runs in

// per-batch I/O of the solve kernel
struct QpIo {
  const double* pos; const double* bv; const double* ba; const double* bj; const double* T;
  const double* lo; const double* hi;  // corridor boxes, B x S each (plans with Kc > 0), else nullptr
  double* coef; int* solved; int* status; int* iters;
  int B, stride;
};

// Bound source encoding of l_src / u_src: -1 -> 0.0; [0, S] -> pos_1d[src]; S+1 .. S+6 -> bound_vel / bound_acc / bound_jerk
// [start, end]; S+7+s -> corridor lo of segment s; 2S+7+s -> corridor hi of segment s.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline double qp_bound_value(const QpIo& io, int b, int S, int src) {
  if (src < 0) return 0.0;
  if (src <= S) return io.pos[(size_t)b * (S + 1) + src];
  if (src < S + 7) {
    const int r = (src - (S + 1)) >> 1, e = (src - (S + 1)) & 1;
    return (r == 0 ? io.bv : (r == 1 ? io.ba : io.bj))[(size_t)b * 2 + e];
  }
  if (src < 2 * S + 7) return io.lo[(size_t)b * S + (src - (S + 7))];
  return io.hi[(size_t)b * S + (src - (2 * S + 7))];
}
