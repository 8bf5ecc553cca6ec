// qp_symbolic.cpp — host-side symbolic analysis for the batched minimum-jerk / minimum-snap QP (K2).
//
// Every 1-D QP of a batch has the same sparsity pattern (it depends only on (order, S)), so everything that is
// pattern-only is computed once here and shared by all threads of the kernel:
//   * the pattern + value recipe of P (upper triangle) and A
//     (reference: src/planner/traj_optimization/src/minimum_control.cpp:5-19 getHessian, :26-96 getConstraintMatrix,
//      generalised to order 7 per SURVEY.md §9.3; the explicit 0.0 entries the reference inserts only influence
//      OSQP's fill-in / rounding, not the mathematics, and are not stored),
//   * where each constraint bound comes from (:98-125 getBound),
//   * the quasi-definite KKT matrix [[P + sigma I, A'], [A, -diag(1/rho)]] (3rd/osqp/algebra/_common/kkt.h:15-21) in a
//     fill-reducing order (our own minimum-degree ordering; OSQP uses AMD — any symmetric permutation of a
//     quasi-definite matrix has an LDL' factorisation), its elimination tree, the pattern of L and, for every row of the
//     up-looking factorisation, the etree reach in topological order (what QDLDL recomputes per call,
//     qdldl_interface.c:85-134).
#include <algorithm>
#include <cstring>
#include <map>
#include <set>
#include <vector>

#include "fpmath.h"
#include "qp_plan.h"

// amd_perm_table.inc defines: static const int* qp_amd_table_lookup(int order, int S, int N)
#include "amd_perm_table.inc"

namespace {

double falling(int j, int r) {
  double v = 1.0;
  for (int k = 0; k < r; k++) v *= (double)(j - k);
  return v;
}

struct Entry { int r, c, seg, pw; double coef; };

}  // namespace

QpPlanHost* qp_plan_build(int order, int S, int Kc) {
  QpPlanHost* pl = new QpPlanHost();
  QpPlanHost& P = *pl;
  const int k = (order + 1) / 2, nc = order + 1;
  const int n = nc * S, m_eq = 2 * k + (k + 1) * (S - 1), m = m_eq + Kc * S, N = n + m;
  P.order = order; P.S = S; P.k = k; P.nc = nc; P.n = n; P.m = m; P.N = N; P.Kc = Kc; P.m_eq = m_eq;

  // ---- P (upper) ------------------------------------------------------------------------------------------
  std::vector<Entry> pe;
  for (int s = 0; s < S; s++)
    for (int j = k; j < nc; j++)      // column
      for (int i = k; i <= j; i++) {  // row <= column
        int e = i + j - 2 * k + 1;
        pe.push_back({nc * s + i, nc * s + j, s, e, falling(i, k) * falling(j, k) / (double)e});
      }
  std::sort(pe.begin(), pe.end(), [](const Entry& a, const Entry& b) { return a.c != b.c ? a.c < b.c : a.r < b.r; });
  P.Pp.assign(n + 1, 0);
  for (auto& e : pe) {
    P.Pp[e.c + 1]++; P.Pi.push_back(e.r); P.P_seg.push_back(e.seg); P.P_pow.push_back(e.pw); P.P_coef.push_back(e.coef);
  }
  for (int c = 0; c < n; c++) P.Pp[c + 1] += P.Pp[c];
  P.nnzP = (int)pe.size();

  // ---- A -----------------------------------------------------------------------------------------------------
  // The reference inserts explicit 0.0 entries (minimum_control.cpp:55,61,64,65,70,71,83,90,91): every derivative row
  // stores all nc coefficients of its segment and the continuity rows also store the zeros in front of the -r! entry.
  // They are kept (value 0 * T^0) because OsqpEigen copies them into OSQP's CSC A, where they shape the KKT pattern,
  // the AMD ordering and therefore the order of floating-point operations in the LDL' factorisation.
  std::vector<Entry> ae;
  for (int r = 0; r < k; r++) ae.push_back({r, r, 0, 0, falling(r, r)});
  auto deriv_row = [&](int row, int s, int r) {
    for (int j = 0; j < nc; j++) {
      if (j < r) ae.push_back({row, nc * s + j, s, 0, 0.0});
      else ae.push_back({row, nc * s + j, s, j - r, falling(j, r)});
    }
  };
  for (int s = 0; s + 1 < S; s++) {
    int base = k + (k + 1) * s;
    deriv_row(base, s, 0);
    for (int r = 0; r < k; r++) {
      deriv_row(base + 1 + r, s, r);
      for (int j = 0; j < r; j++) ae.push_back({base + 1 + r, nc * (s + 1) + j, 0, 0, 0.0});
      ae.push_back({base + 1 + r, nc * (s + 1) + r, 0, 0, -falling(r, r)});
    }
  }
  {
    int base = k + (k + 1) * (S - 1);
    for (int r = 0; r < k; r++) deriv_row(base + r, S - 1, r);
  }
  // corridor rows (extension; no counterpart in minimum_control.cpp, whose rows are all equalities): the
  // position of segment s at the interior sample time phi_j T_s, phi_j = (j + 1) / (Kc + 1), as sum_i c_i phi_j^i T_s^i —
  // the entry is (phi_j^i) * T_s^i, evaluated exactly like that by oracle/minctrl_ref.cpp
  for (int s = 0; s < S; s++)
    for (int j = 0; j < Kc; j++) {
      const double phi = (double)(j + 1) / (double)(Kc + 1);
      for (int i = 0; i < nc; i++) ae.push_back({m_eq + s * Kc + j, nc * s + i, s, i, fpm::powi(phi, i)});
    }
  std::sort(ae.begin(), ae.end(), [](const Entry& a, const Entry& b) { return a.c != b.c ? a.c < b.c : a.r < b.r; });
  P.Ap.assign(n + 1, 0);
  for (auto& e : ae) {
    P.Ap[e.c + 1]++; P.Ai.push_back(e.r); P.A_seg.push_back(e.seg); P.A_pow.push_back(e.pw); P.A_coef.push_back(e.coef);
  }
  for (int c = 0; c < n; c++) P.Ap[c + 1] += P.Ap[c];
  P.nnzA = (int)ae.size();

  // ---- bounds: index into [pos_1d (S+1) | bound_vel (2) | bound_acc (2) | bound_jerk (2)], -1 -> 0.0 ------------
  P.l_src.assign(m, -1);
  P.l_src[0] = 0;
  for (int r = 1; r < k; r++) P.l_src[r] = (S + 1) + 2 * (r - 1);
  int eb = k + (k + 1) * (S - 1);
  P.l_src[eb] = S;
  for (int r = 1; r < k; r++) P.l_src[eb + r] = (S + 1) + 2 * (r - 1) + 1;
  for (int s = 0; s + 1 < S; s++) P.l_src[k + (k + 1) * s] = s + 1;
  P.u_src = P.l_src;  // equality rows: l == u (minimum_control.cpp:98-125)
  for (int s = 0; s < S; s++)
    for (int j = 0; j < Kc; j++) { P.l_src[m_eq + s * Kc + j] = (S + 7) + s; P.u_src[m_eq + s * Kc + j] = (2 * S + 7) + s; }

  // ---- KKT = [[P + sigma I, A'], [A, -diag(1/rho)]], upper-triangular CSC, entries in the order OSQP's form_KKT
  // writes them (3rd/osqp/algebra/_common/kkt.c:254-291 _kkt_assemble_csc): column c < n holds P's column c and, when P
  // has no diagonal there, a structural diagonal last; column n + r holds row r of A in increasing variable order, then
  // the diagonal.  entry sources: kind 0 P off-diagonal, 1 P diagonal (+sigma), 2 sigma only, 3 A, 4 -1/rho
  struct KE { int i, j, kind, idx; };
  std::vector<std::vector<KE>> kcol(N);
  for (int c = 0; c < n; c++) {
    bool has_diag = false;
    for (int p = P.Pp[c]; p < P.Pp[c + 1]; p++) {
      int r = P.Pi[p];
      if (r == c) { kcol[c].push_back({r, c, 1, p}); has_diag = true; }
      else kcol[c].push_back({r, c, 0, p});
    }
    if (!has_diag) kcol[c].push_back({c, c, 2, 0});
  }
  for (int c = 0; c < n; c++)
    for (int p = P.Ap[c]; p < P.Ap[c + 1]; p++) kcol[n + P.Ai[p]].push_back({c, n + P.Ai[p], 3, p});
  for (int r = 0; r < m; r++) kcol[n + r].push_back({n + r, n + r, 4, r});
  P.Kp0.assign(N + 1, 0);
  P.Ki0.clear();
  for (int c = 0; c < N; c++) {
    for (auto& e : kcol[c]) P.Ki0.push_back(e.i);
    P.Kp0[c + 1] = (int)P.Ki0.size();
  }

  // ---- fill-reducing order -------------------------------------------------------------------------------------------
  // OSQP orders the KKT matrix with AMD (qdldl_interface.c:137-160 amd_l_order).  For the (order, S) families in
  // amd_perm_table.inc the permutation AMD returns for exactly this pattern is tabulated (generated by
  // tests/golden/make_amd_tables.py from the reference's own AMD), which makes the factorisation — and with it every ADMM
  // iterate — follow the reference operation for operation.  Other families fall back to a plain minimum-degree order
  // (any symmetric permutation of a quasi-definite matrix has an LDL'; results then agree to rounding only).
  std::vector<int> perm(N), inv(N, -1);
  const int* tab = qp_amd_table_lookup(order, S, N);
  P.perm_from_table = tab != nullptr;
  if (tab) {
    for (int i = 0; i < N; i++) perm[i] = tab[i];
  } else {
    std::vector<std::set<int>> adj(N);
    for (int c = 0; c < N; c++)
      for (auto& e : kcol[c]) if (e.i != e.j) { adj[e.i].insert(e.j); adj[e.j].insert(e.i); }
    std::vector<char> done(N, 0);
    for (int step = 0; step < N; step++) {
      int best = -1; size_t bd = (size_t)-1;
      for (int v = 0; v < N; v++) if (!done[v] && adj[v].size() < bd) { bd = adj[v].size(); best = v; }
      perm[step] = best; done[best] = 1;
      std::vector<int> nb(adj[best].begin(), adj[best].end());
      for (int a : nb) adj[a].erase(best);
      for (size_t x = 0; x < nb.size(); x++)
        for (size_t y = x + 1; y < nb.size(); y++) { adj[nb[x]].insert(nb[y]); adj[nb[y]].insert(nb[x]); }
      adj[best].clear();
    }
  }
  for (int i = 0; i < N; i++) inv[perm[i]] = i;
  P.perm = perm;

  // ---- permuted upper CSC, entries in the order csc_symperm emits them (csc_utils.c:326-386): original columns in
  // order, entries in order, each appended to column max(pinv[i], pinv[j]) with row min(...) — rows are NOT sorted
  // within a column, and QDLDL's reach order (hence the summation order) follows that sequence
  std::vector<std::vector<KE>> pcol(N);
  for (int c = 0; c < N; c++)
    for (auto& e : kcol[c]) {
      int a = inv[e.i], b = inv[e.j];
      pcol[std::max(a, b)].push_back({std::min(a, b), std::max(a, b), e.kind, e.idx});
    }
  P.Kp.assign(N + 1, 0);
  P.Ki.clear(); P.Kkind.clear(); P.Kidx.clear();
  for (int c = 0; c < N; c++) {
    for (auto& e : pcol[c]) { P.Ki.push_back(e.i); P.Kkind.push_back(e.kind); P.Kidx.push_back(e.idx); }
    P.Kp[c + 1] = (int)P.Ki.size();
  }
  P.nnzK = (int)P.Ki.size();

  // ---- elimination tree + column counts of L (same algorithm QDLDL_etree runs per setup) ---------------------------------------
  std::vector<int> etree(N, -1), work(N, 0), Lnz(N, 0);
  for (int j = 0; j < N; j++) {
    work[j] = j;
    for (int p = P.Kp[j]; p < P.Kp[j + 1]; p++) {
      int i = P.Ki[p];
      while (work[i] != j) {
        if (etree[i] == -1) etree[i] = j;
        Lnz[i]++;
        work[i] = j;
        i = etree[i];
      }
    }
  }
  P.Lp.assign(N + 1, 0);
  for (int i = 0; i < N; i++) P.Lp[i + 1] = P.Lp[i] + Lnz[i];
  P.nnzL = P.Lp[N];
  P.Li.assign(P.nnzL, 0);
  // ---- row-by-row reach (topological) and the slot every new L entry lands in ---------------------------------------------------
  std::vector<int> next(N);
  for (int i = 0; i < N; i++) next[i] = P.Lp[i];
  std::vector<char> mark(N, 0);
  P.Rp.assign(N + 1, 0);
  for (int kk = 0; kk < N; kk++) {
    std::vector<int> yidx;
    for (int p = P.Kp[kk]; p < P.Kp[kk + 1]; p++) {
      int b = P.Ki[p];
      if (b == kk) continue;
      if (!mark[b]) {
        std::vector<int> chain;
        int nx = b;
        while (nx != -1 && nx < kk && !mark[nx]) { mark[nx] = 1; chain.push_back(nx); nx = etree[nx]; }
        for (int q = (int)chain.size() - 1; q >= 0; q--) yidx.push_back(chain[q]);
      }
    }
    for (int q = (int)yidx.size() - 1; q >= 0; q--) {  // visited back to front == topological order
      int c = yidx[q];
      P.Rc.push_back(c);
      P.Rpos.push_back(next[c]);
      P.Li[next[c]] = kk;
      next[c]++;
      mark[c] = 0;
    }
    P.Rp[kk + 1] = (int)P.Rc.size();
  }
  // ---- L in the order QDLDL_Ltsolve consumes it: columns N-1 .. 0, entries of a column in storage order ---------------
  P.Ltpos.assign(P.nnzL, 0); P.LtR.assign(P.nnzL, 0); P.LtEnd.assign(N + 1, 0);
  {
    int k = 0;
    for (int t = 0; t < N; t++) {
      const int i = N - 1 - t;
      for (int j = P.Lp[i]; j < P.Lp[i + 1]; j++) { P.Ltpos[j] = k; P.LtR[k] = P.Li[j]; k++; }
      P.LtEnd[t] = k;
    }
    P.LtEnd[N] = k;
  }
  // ---- views for the warp-per-problem kernel ---------------------------------------------------------------------------
  {  // CSR of A, entries of a row by increasing column (the order in which the CSC loop adds them to that row)
    P.Arp.assign(m + 1, 0);
    for (int e = 0; e < P.nnzA; e++) P.Arp[P.Ai[e] + 1]++;
    for (int r = 0; r < m; r++) P.Arp[r + 1] += P.Arp[r];
    P.Arj.assign(P.nnzA, 0); P.Arx.assign(P.nnzA, 0);
    std::vector<int> nxt(P.Arp.begin(), P.Arp.end() - 1);
    for (int c = 0; c < n; c++)
      for (int e = P.Ap[c]; e < P.Ap[c + 1]; e++) { int r = P.Ai[e]; P.Arj[nxt[r]] = c; P.Arx[nxt[r]] = e; nxt[r]++; }
  }
  {  // P*v with the upper triangle: contributions to element i in the order of the CSC loop (qp_P_mul)
    std::vector<std::vector<std::pair<int, int>>> lst(n);
    for (int c = 0; c < n; c++)
      for (int e = P.Pp[c]; e < P.Pp[c + 1]; e++) {
        int r = P.Pi[e];
        lst[r].push_back({e, c});
        if (r != c) lst[c].push_back({e, r});
      }
    P.Psp.assign(n + 1, 0);
    for (int i = 0; i < n; i++) {
      for (auto& pr : lst[i]) { P.Psa.push_back(pr.first); P.Psv.push_back(pr.second); }
      P.Psp[i + 1] = (int)P.Psa.size();
    }
  }
  {  // L' solve levels: x_i needs x_r for every row r of column i (all r > i)
    std::vector<int> lev(N, 0);
    int mx = 0;
    for (int i = N - 1; i >= 0; i--) {
      int l = 0;
      for (int j = P.Lp[i]; j < P.Lp[i + 1]; j++) l = std::max(l, lev[P.Li[j]] + 1);
      lev[i] = l; mx = std::max(mx, l);
    }
    P.nlev = mx;  // level 0 (columns without entries) needs no work
    P.LevP.assign(mx + 1, 0);
    for (int l = 1; l <= mx; l++) {
      for (int i = N - 1; i >= 0; i--) if (lev[i] == l) P.LevC.push_back(i);
      P.LevP[l] = (int)P.LevC.size();
    }
  }
  {  // L by rows + forward-solve levels + the packed uint16 index block of the warp-per-problem solves
    const int nnzL = P.nnzL;
    P.Lrp.assign(N + 1, 0);
    for (int j = 0; j < nnzL; j++) P.Lrp[P.Li[j] + 1]++;
    for (int r = 0; r < N; r++) P.Lrp[r + 1] += P.Lrp[r];
    P.Lrc.assign(nnzL, 0); P.Lrx.assign(nnzL, 0);
    std::vector<int> nxt(P.Lrp.begin(), P.Lrp.end() - 1);
    for (int i = 0; i < N; i++)  // columns ascending: every row receives its entries by increasing column
      for (int j = P.Lp[i]; j < P.Lp[i + 1]; j++) { const int r = P.Li[j]; P.Lrc[nxt[r]] = i; P.Lrx[nxt[r]] = j; nxt[r]++; }
    std::vector<int> lev(N, 0);
    int mx = 0;
    for (int r = 0; r < N; r++) {  // row r needs x[i] for every column i < r it has an entry in
      int l = 0;
      for (int p = P.Lrp[r]; p < P.Lrp[r + 1]; p++) l = std::max(l, lev[P.Lrc[p]] + 1);
      lev[r] = l; mx = std::max(mx, l);
    }
    P.nlevf = mx;  // level 0 (rows without entries) needs no work
    P.LevFP.assign(mx + 1, 0);
    P.LevFR.clear();
    for (int l = 1; l <= mx; l++) {
      for (int r = 0; r < N; r++) if (lev[r] == l) P.LevFR.push_back(r);
      P.LevFP[l] = (int)P.LevFR.size();
    }
    auto put = [&](const std::vector<int>& v) { int at = (int)P.Sidx.size(); for (int x : v) P.Sidx.push_back((unsigned short)x); return at; };
    P.Sidx.clear();
    if (N < 65536 && nnzL < 65536) {
      P.sx_Lrp = put(P.Lrp); P.sx_Lrc = put(P.Lrc); P.sx_Lrx = put(P.Lrx); P.sx_LevFP = put(P.LevFP); P.sx_LevFR = put(P.LevFR);
      P.sx_Lp = put(P.Lp); P.sx_Li = put(P.Li); P.sx_LevP = put(P.LevP); P.sx_LevC = put(P.LevC); P.sx_perm = put(P.perm);
      if (P.Sidx.size() & 1) P.Sidx.push_back(0);
    }
    // ---- lane schedule: the rows (forward) / columns (backward) of one dependency level are packed into chunks of 32 lanes, a
    // row's entries on consecutive lanes starting at its leader lane, in the order the sequential solve consumes them
    struct Item { int dst; std::vector<std::pair<int, int>> ent; };  // destination index, (Lx slot, source index) in order
    auto pack = [&](const std::vector<std::vector<Item>>& levels) {
      for (const auto& lv : levels) {
        std::vector<unsigned int> rec(64, 0);
        int used = 0;
        auto flush = [&]() {
          int mx = 0;
          for (int l = 0; l < 32; l++) mx = std::max(mx, (int)((rec[2 * l + 1] >> 16) & 0xff));
          mx = (mx + 3) & ~3;  // the chain loop runs four entries per trip
          for (int l = 0; l < 32; l++) rec[2 * l + 1] |= (unsigned)mx << 24;
          P.Sch.insert(P.Sch.end(), rec.begin(), rec.end());
          rec.assign(64, 0); used = 0;
        };
        for (int l = 0; l < 32; l++) { rec[2 * l] = 0xffffu; }
        for (const Item& it : lv) {
          const int len = (int)it.ent.size();
          if (len == 0) continue;
          used += used & 1;  // leaders sit on even lanes: the leader reads its row's products two at a time (16 B aligned)
          if (used + len > 32) { flush(); for (int l = 0; l < 32; l++) rec[2 * l] = 0xffffu; }
          for (int k = 0; k < len; k++) rec[2 * (used + k)] = (unsigned)it.ent[k].first | ((unsigned)it.ent[k].second << 16);
          rec[2 * used + 1] = (unsigned)it.dst | ((unsigned)len << 16);
          used += len;
        }
        if (used) flush();
      }
    };
    bool ok = N < 65535 && nnzL < 65535;
    for (int r = 0; r < N && ok; r++) if (P.Lrp[r + 1] - P.Lrp[r] > 32 || P.Lp[r + 1] - P.Lp[r] > 32) ok = false;
    P.Sch.clear(); P.sch_nf = P.sch_n = 0;
    if (ok) {
      std::vector<std::vector<Item>> fl(P.nlevf), bl(P.nlev);
      for (int l = 0; l < P.nlevf; l++)
        for (int q = P.LevFP[l]; q < P.LevFP[l + 1]; q++) {
          Item it; it.dst = P.LevFR[q];
          for (int e = P.Lrp[it.dst]; e < P.Lrp[it.dst + 1]; e++) it.ent.push_back({P.Lrx[e], P.Lrc[e]});
          fl[l].push_back(it);
        }
      for (int l = 0; l < P.nlev; l++)
        for (int q = P.LevP[l]; q < P.LevP[l + 1]; q++) {
          Item it; it.dst = P.LevC[q];
          for (int j = P.Lp[it.dst]; j < P.Lp[it.dst + 1]; j++) it.ent.push_back({j, P.Li[j]});
          bl[l].push_back(it);
        }
      pack(fl); P.sch_nf = (int)P.Sch.size() / 64;
      pack(bl); P.sch_n = (int)P.Sch.size() / 64;
    }
  }
  return pl;
}

void qp_plan_pack(const QpPlanHost& H, std::vector<int>& ints, std::vector<double>& dbls, QpPlanOffsets& o) {
  ints.clear();
  auto push = [&](const std::vector<int>& v) { size_t at = ints.size(); ints.insert(ints.end(), v.begin(), v.end()); return at; };
  o.Pp = push(H.Pp); o.Pi = push(H.Pi); o.P_seg = push(H.P_seg); o.P_pow = push(H.P_pow);
  o.Ap = push(H.Ap); o.Ai = push(H.Ai); o.A_seg = push(H.A_seg); o.A_pow = push(H.A_pow);
  o.l_src = push(H.l_src); o.u_src = push(H.u_src); o.perm = push(H.perm); o.Kp = push(H.Kp); o.Ki = push(H.Ki); o.Kkind = push(H.Kkind);
  o.Kidx = push(H.Kidx); o.Lp = push(H.Lp); o.Li = push(H.Li); o.Rp = push(H.Rp); o.Rc = push(H.Rc); o.Rpos = push(H.Rpos);
  o.Ltpos = push(H.Ltpos); o.LtR = push(H.LtR); o.LtEnd = push(H.LtEnd);
  o.Arp = push(H.Arp); o.Arj = push(H.Arj); o.Arx = push(H.Arx); o.Psp = push(H.Psp); o.Psa = push(H.Psa); o.Psv = push(H.Psv);
  o.LevP = push(H.LevP); o.LevC = push(H.LevC);
  dbls = H.P_coef;
  o.A_coef = dbls.size();
  dbls.insert(dbls.end(), H.A_coef.begin(), H.A_coef.end());
}

void qp_plan_bind(const QpPlanHost& H, const QpPlanOffsets& o, const int* I, const double* Dbl, QpPlanDev& D) {
  D.order = H.order; D.S = H.S; D.k = H.k; D.nc = H.nc; D.n = H.n; D.m = H.m; D.N = H.N;
  D.nnzP = H.nnzP; D.nnzA = H.nnzA; D.nnzK = H.nnzK; D.nnzL = H.nnzL; D.Kc = H.Kc; D.m_eq = H.m_eq;
  D.Pp = I + o.Pp; D.Pi = I + o.Pi; D.P_seg = I + o.P_seg; D.P_pow = I + o.P_pow;
  D.Ap = I + o.Ap; D.Ai = I + o.Ai; D.A_seg = I + o.A_seg; D.A_pow = I + o.A_pow;
  D.P_coef = Dbl; D.A_coef = Dbl + o.A_coef;
  D.l_src = I + o.l_src; D.u_src = I + o.u_src; D.perm = I + o.perm; D.Kp = I + o.Kp; D.Ki = I + o.Ki; D.Kkind = I + o.Kkind; D.Kidx = I + o.Kidx;
  D.Lp = I + o.Lp; D.Li = I + o.Li; D.Rp = I + o.Rp; D.Rc = I + o.Rc; D.Rpos = I + o.Rpos;
  D.Ltpos = I + o.Ltpos; D.LtR = I + o.LtR; D.LtEnd = I + o.LtEnd;
  D.Arp = I + o.Arp; D.Arj = I + o.Arj; D.Arx = I + o.Arx; D.Psp = I + o.Psp; D.Psa = I + o.Psa; D.Psv = I + o.Psv;
  D.LevP = I + o.LevP; D.LevC = I + o.LevC; D.nlev = H.nlev;
  D.Sidx = nullptr; D.n_sidx = (int)H.Sidx.size(); D.nlevf = H.nlevf;
  D.Sch = nullptr; D.sch_nf = H.sch_nf; D.sch_n = H.sch_n;
  D.sx_Lrp = H.sx_Lrp; D.sx_Lrc = H.sx_Lrc; D.sx_Lrx = H.sx_Lrx; D.sx_LevFP = H.sx_LevFP; D.sx_LevFR = H.sx_LevFR;
  D.sx_Lp = H.sx_Lp; D.sx_Li = H.sx_Li; D.sx_LevP = H.sx_LevP; D.sx_LevC = H.sx_LevC; D.sx_perm = H.sx_perm;
  // workspace layout (offsets in doubles; element e of problem b lives at ws[e * stride + b])
  int at = 0;
  auto take = [&](int len) { int r = at; at += len; return r; };
  const int n = H.n, m = H.m, N = H.N;
  D.o_Px = take(H.nnzP); D.o_Ax = take(H.nnzA); D.o_q = take(n); D.o_l = take(m); D.o_u = take(m);
  D.o_D = take(n); D.o_Dinv = take(n); D.o_E = take(m); D.o_Einv = take(m); D.o_rho = take(m); D.o_rhoinv = take(m);
  D.o_Lx = take(H.nnzL); D.o_Dd = take(N); D.o_Ddinv = take(N); D.o_yw = take(N);
  D.o_x = take(n); D.o_xprev = take(n); D.o_dx = take(n); D.o_Pxv = take(n); D.o_Aty = take(n);
  D.o_z = take(m); D.o_zprev = take(m); D.o_y = take(m); D.o_dy = take(m); D.o_Axv = take(m);
  D.o_xz = take(N); D.o_bp = take(N); D.o_tn = take(n); D.o_tm = take(m);
  at = (at + 1) & ~1; D.o_pb = take(64);  // the products of one schedule chunk (warp kernel; 16 B aligned)
  D.ws_warp = at;              // the warp-per-problem kernel keeps everything up to here in shared memory
  D.o_LxT = take(H.nnzL);      // thread-per-problem kernel only: L in the L' solve's consumption order
  D.ws_doubles = at;
}
