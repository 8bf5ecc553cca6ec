/* oracle/osqp_ref_driver.c — TEST INFRASTRUCTURE.  Compiled INTO oracle/_ref/libosqp_ref.so together with the
 * reference's unmodified OSQP C sources (see Makefile); exposes one flat entry point so that liboracle.so never needs
 * OSQP's headers.  It does what OsqpEigen::Solver::initSolver / solve / getSolution do
 * (3rd/osqp-eigen/src/Solver.cpp:84-153,173-190,218-225): osqp_setup, osqp_solve, copy solution->x. */
#include <stdlib.h>
#include <string.h>
#include "osqp.h"
#include "oracle_osqp.h"

/* exposes the scaled workspace for white-box checks of the GPU kernel (D, E, c, rho) */
#include "types.h"
#include "algebra_vector.h"

int osqp_ref_solve(long long n, long long m, const long long* Pp, const long long* Pi, const double* Px,
                   const double* q, const long long* Ap, const long long* Ai, const double* Ax, const double* l,
                   const double* u, const oracle_osqp_settings* os, double* x, double* y, oracle_osqp_info* info) {
  OSQPSettings* st = (OSQPSettings*)malloc(sizeof(OSQPSettings));
  OSQPSolver* solver = NULL;
  OSQPCscMatrix P, A;
  OSQPInt flag;
  osqp_set_default_settings(st);
  /* MinimumControl::solve's overrides (minimum_control.cpp:160-162) arrive through `os`; OsqpEigen leaves the rest */
  st->rho = os->rho; st->sigma = os->sigma; st->alpha = os->alpha;
  st->eps_abs = os->eps_abs; st->eps_rel = os->eps_rel; st->eps_prim_inf = os->eps_prim_inf;
  st->eps_dual_inf = os->eps_dual_inf; st->max_iter = os->max_iter; st->check_termination = os->check_termination;
  st->scaling = os->scaling; st->adaptive_rho = os->adaptive_rho; st->adaptive_rho_interval = os->adaptive_rho_interval;
  st->adaptive_rho_tolerance = os->adaptive_rho_tolerance;
  st->warm_starting = 1; /* setWarmStart(true) */
  st->verbose = 0;
  st->polishing = 0;
  P.m = n; P.n = n; P.p = (OSQPInt*)Pp; P.i = (OSQPInt*)Pi; P.x = (OSQPFloat*)Px; P.nzmax = Pp[n]; P.nz = -1;
  A.m = m; A.n = n; A.p = (OSQPInt*)Ap; A.i = (OSQPInt*)Ai; A.x = (OSQPFloat*)Ax; A.nzmax = Ap[n]; A.nz = -1;
  memset(info, 0, sizeof(*info));
  flag = osqp_setup(&solver, &P, q, &A, l, u, m, n, st);
  info->setup_flag = (int)flag;
  if (flag) { free(st); if (solver) osqp_cleanup(solver); return (int)flag; }
  info->adaptive_rho_interval_used = (int)solver->settings->adaptive_rho_interval;
  if (solver->work->scaling) {
    info->scaling_c = solver->work->scaling->c;
    if (os->want_scaling_dump && os->dump_D && os->dump_E) {
      OSQPVectorf_to_raw(os->dump_D, solver->work->scaling->D);
      OSQPVectorf_to_raw(os->dump_E, solver->work->scaling->E);
    }
  }
  flag = osqp_solve(solver);
  info->solve_flag = (int)flag;
  info->status_val = (int)solver->info->status_val;
  info->iter = (int)solver->info->iter;
  info->rho_updates = (int)solver->info->rho_updates;
  info->rho_final = solver->settings->rho;
  info->prim_res = solver->info->prim_res;
  info->dual_res = solver->info->dual_res;
  info->obj_val = solver->info->obj_val;
  if (solver->solution) {
    memcpy(x, solver->solution->x, (size_t)n * sizeof(double));
    if (y && m) memcpy(y, solver->solution->y, (size_t)m * sizeof(double));
  }
  osqp_cleanup(solver);
  free(st);
  return 0;
}

/* KKT solve at the linear-system boundary, for the solve_linsys known-answer test
 * (3rd/osqp/tests/solve_linsys/generate_problem.py): x solves [[P + sigma I, A'], [A, -1/rho I]] x = rhs */
#include "lin_alg.h"
int osqp_ref_kkt_solve(long long n, long long m, const long long* Pp, const long long* Pi, const double* Px,
                       const long long* Ap, const long long* Ai, const double* Ax, double sigma, double rho,
                       const double* rhs, double* sol) {
  OSQPSettings* st = (OSQPSettings*)malloc(sizeof(OSQPSettings));
  OSQPCscMatrix Pc, Ac;
  OSQPMatrix *Pm, *Am;
  OSQPVectorf *rv, *rho_vec;
  LinSysSolver* s = NULL;
  OSQPInt flag;
  OSQPFloat pr = 0, dr = 0;
  long long i;
  osqp_set_default_settings(st);
  st->sigma = sigma; st->rho = rho; st->rho_is_vec = 0;
  Pc.m = n; Pc.n = n; Pc.p = (OSQPInt*)Pp; Pc.i = (OSQPInt*)Pi; Pc.x = (OSQPFloat*)Px; Pc.nzmax = Pp[n]; Pc.nz = -1;
  Ac.m = m; Ac.n = n; Ac.p = (OSQPInt*)Ap; Ac.i = (OSQPInt*)Ai; Ac.x = (OSQPFloat*)Ax; Ac.nzmax = Ap[n]; Ac.nz = -1;
  Pm = OSQPMatrix_new_from_csc(&Pc, 1);
  Am = OSQPMatrix_new_from_csc(&Ac, 0);
  rv = OSQPVectorf_new(rhs, n + m);
  rho_vec = OSQPVectorf_malloc(m);
  OSQPVectorf_set_scalar(rho_vec, rho);
  flag = osqp_algebra_init_linsys_solver(&s, Pm, Am, rho_vec, st, &pr, &dr, 0);
  if (!flag) {
    s->solve(s, rv, 1);
    OSQPVectorf_to_raw(sol, rv);
    s->free(s);
  }
  (void)i;
  OSQPMatrix_free(Pm); OSQPMatrix_free(Am); OSQPVectorf_free(rv); OSQPVectorf_free(rho_vec);
  free(st);
  return (int)flag;
}
