/* oracle/qdldl/qdldl.h — TEST INFRASTRUCTURE.  Restatement of the public QDLDL 0.1.6 API
 * (github.com/osqp/qdldl @ 29d140419a3bec20d860052d73ba2be927faf5a1).  The source is NOT vendored in the reference
 * (it is fetched at build time, 3rd/osqp/algebra/_common/lin_sys/qdldl/qdldl.cmake:6-12); the signatures are fixed by
 * the call sites qdldl_interface.c:94,116-119,409.  Pinned by tests/test_osqp_kat.py (KKT solve vs SciPy, OSQP KATs). */
#ifndef QDLDL_H
#define QDLDL_H
#include "qdldl_types.h"
#define QDLDL_VERSION_MAJOR 0
#define QDLDL_VERSION_MINOR 1
#define QDLDL_VERSION_PATCH 6
#ifdef __cplusplus
extern "C" {
#endif
QDLDL_int QDLDL_etree(const QDLDL_int n, const QDLDL_int* Ap, const QDLDL_int* Ai, QDLDL_int* work, QDLDL_int* Lnz,
                      QDLDL_int* etree);
QDLDL_int QDLDL_factor(const QDLDL_int n, const QDLDL_int* Ap, const QDLDL_int* Ai, const QDLDL_float* Ax,
                       QDLDL_int* Lp, QDLDL_int* Li, QDLDL_float* Lx, QDLDL_float* D, QDLDL_float* Dinv,
                       const QDLDL_int* Lnz, const QDLDL_int* etree, QDLDL_bool* bwork, QDLDL_int* iwork,
                       QDLDL_float* fwork);
void QDLDL_solve(const QDLDL_int n, const QDLDL_int* Lp, const QDLDL_int* Li, const QDLDL_float* Lx,
                 const QDLDL_float* Dinv, QDLDL_float* x);
void QDLDL_Lsolve(const QDLDL_int n, const QDLDL_int* Lp, const QDLDL_int* Li, const QDLDL_float* Lx, QDLDL_float* x);
void QDLDL_Ltsolve(const QDLDL_int n, const QDLDL_int* Lp, const QDLDL_int* Li, const QDLDL_float* Lx, QDLDL_float* x);
#ifdef __cplusplus
}
#endif
#endif
