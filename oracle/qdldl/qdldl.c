/* oracle/qdldl/qdldl.c — TEST INFRASTRUCTURE.  Restatement of QDLDL 0.1.6 (see qdldl.h): elimination tree,
 * up-looking LDL^T of an upper-triangular CSC quasi-definite matrix, and the L / D / L^T solves.
 * The order of floating-point operations follows the published algorithm: row k of L is obtained by a sparse
 * triangular solve over the etree reach of column k of A, visited in topological order; contributions to D[k] are
 * subtracted in that same order. */
#include "qdldl.h"

#define Q_UNKNOWN (-1)
#define Q_USED (1)
#define Q_UNUSED (0)

QDLDL_int QDLDL_etree(const QDLDL_int n, const QDLDL_int* Ap, const QDLDL_int* Ai, QDLDL_int* work, QDLDL_int* Lnz,
                      QDLDL_int* etree) {
  QDLDL_int i, j, p, sum = 0;
  for (i = 0; i < n; i++) {
    work[i] = 0;
    Lnz[i] = 0;
    etree[i] = Q_UNKNOWN;
    if (Ap[i] == Ap[i + 1]) return -1; /* empty column: the diagonal entry is missing */
  }
  for (j = 0; j < n; j++) {
    work[j] = j;
    for (p = Ap[j]; p < Ap[j + 1]; p++) {
      i = Ai[p];
      if (i > j) return -1; /* entry below the diagonal */
      while (work[i] != j) {
        if (etree[i] == Q_UNKNOWN) etree[i] = j;
        Lnz[i]++; /* column i of L gains row j */
        work[i] = j;
        i = etree[i];
      }
    }
  }
  for (i = 0; i < n; i++) {
    if (sum > QDLDL_INT_MAX - Lnz[i]) return -2;
    sum += Lnz[i];
  }
  return sum;
}

QDLDL_int QDLDL_factor(const QDLDL_int n, const QDLDL_int* Ap, const QDLDL_int* Ai, const QDLDL_float* Ax,
                       QDLDL_int* Lp, QDLDL_int* Li, QDLDL_float* Lx, QDLDL_float* D, QDLDL_float* Dinv,
                       const QDLDL_int* Lnz, const QDLDL_int* etree, QDLDL_bool* bwork, QDLDL_int* iwork,
                       QDLDL_float* fwork) {
  QDLDL_int i, j, k, nnzY, bidx, cidx, nextIdx, nnzE, tmpIdx;
  QDLDL_int positive = 0;
  QDLDL_bool* yMarkers = bwork;
  QDLDL_int* yIdx = iwork;
  QDLDL_int* elimBuffer = iwork + n;
  QDLDL_int* LNextSpaceInCol = iwork + 2 * n;
  QDLDL_float* yVals = fwork;
  QDLDL_float yVals_cidx;

  Lp[0] = 0;
  for (i = 0; i < n; i++) {
    Lp[i + 1] = Lp[i] + Lnz[i];
    yMarkers[i] = Q_UNUSED;
    yVals[i] = 0.0;
    D[i] = 0.0;
    LNextSpaceInCol[i] = Lp[i];
  }
  /* the first column holds only its diagonal */
  D[0] = Ax[0];
  if (D[0] == 0.0) return -1;
  if (D[0] > 0.0) positive++;
  Dinv[0] = 1 / D[0];

  for (k = 1; k < n; k++) {
    nnzY = 0; /* number of non-zeros in row k of L found so far */
    tmpIdx = Ap[k + 1];
    for (i = Ap[k]; i < tmpIdx; i++) {
      bidx = Ai[i];
      if (bidx == k) {
        D[k] = Ax[i];
        continue;
      }
      yVals[bidx] = Ax[i];
      nextIdx = bidx;
      if (yMarkers[nextIdx] == Q_UNUSED) { /* climb the etree from bidx, collecting unvisited nodes below k */
        yMarkers[nextIdx] = Q_USED;
        elimBuffer[0] = nextIdx;
        nnzE = 1;
        nextIdx = etree[bidx];
        while (nextIdx != Q_UNKNOWN && nextIdx < k) {
          if (yMarkers[nextIdx] == Q_USED) break;
          yMarkers[nextIdx] = Q_USED;
          elimBuffer[nnzE] = nextIdx;
          nnzE++;
          nextIdx = etree[nextIdx];
        }
        while (nnzE) { /* append the chain in reverse so that yIdx read backwards is topological */
          yIdx[nnzY++] = elimBuffer[--nnzE];
        }
      }
    }
    for (i = nnzY - 1; i >= 0; i--) {
      cidx = yIdx[i];
      tmpIdx = LNextSpaceInCol[cidx];
      yVals_cidx = yVals[cidx];
      for (j = Lp[cidx]; j < tmpIdx; j++) yVals[Li[j]] -= Lx[j] * yVals_cidx;
      Li[tmpIdx] = k;
      Lx[tmpIdx] = yVals_cidx * Dinv[cidx];
      D[k] -= yVals_cidx * Lx[tmpIdx];
      LNextSpaceInCol[cidx]++;
      yVals[cidx] = 0.0;
      yMarkers[cidx] = Q_UNUSED;
    }
    if (D[k] == 0.0) return -1;
    if (D[k] > 0.0) positive++;
    Dinv[k] = 1 / D[k];
  }
  return positive;
}

void QDLDL_Lsolve(const QDLDL_int n, const QDLDL_int* Lp, const QDLDL_int* Li, const QDLDL_float* Lx, QDLDL_float* x) {
  QDLDL_int i, j;
  for (i = 0; i < n; i++) {
    QDLDL_float val = x[i];
    for (j = Lp[i]; j < Lp[i + 1]; j++) x[Li[j]] -= Lx[j] * val;
  }
}

void QDLDL_Ltsolve(const QDLDL_int n, const QDLDL_int* Lp, const QDLDL_int* Li, const QDLDL_float* Lx, QDLDL_float* x) {
  QDLDL_int i, j;
  for (i = n - 1; i >= 0; i--) {
    QDLDL_float val = x[i];
    for (j = Lp[i]; j < Lp[i + 1]; j++) val -= Lx[j] * x[Li[j]];
    x[i] = val;
  }
}

void QDLDL_solve(const QDLDL_int n, const QDLDL_int* Lp, const QDLDL_int* Li, const QDLDL_float* Lx,
                 const QDLDL_float* Dinv, QDLDL_float* x) {
  QDLDL_int i;
  QDLDL_Lsolve(n, Lp, Li, Lx, x);
  for (i = 0; i < n; i++) x[i] *= Dinv[i];
  QDLDL_Ltsolve(n, Lp, Li, Lx, x);
}
