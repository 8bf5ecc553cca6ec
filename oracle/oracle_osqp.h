/* oracle/oracle_osqp.h — flat interface between liboracle.so and oracle/_ref/libosqp_ref.so (TEST INFRASTRUCTURE). */
#ifndef ORACLE_OSQP_H
#define ORACLE_OSQP_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  int max_iter, check_termination, scaling, adaptive_rho, adaptive_rho_interval;
  double adaptive_rho_tolerance;
  int want_scaling_dump;
  double* dump_D; /* n */
  double* dump_E; /* m */
} oracle_osqp_settings;
typedef struct {
  int setup_flag, solve_flag, status_val, iter, rho_updates, adaptive_rho_interval_used;
  double rho_final, prim_res, dual_res, obj_val, scaling_c;
} oracle_osqp_info;
typedef int (*osqp_ref_solve_fn)(long long n, long long m, const long long* Pp, const long long* Pi, const double* Px,
                                 const double* q, const long long* Ap, const long long* Ai, const double* Ax,
                                 const double* l, const double* u, const oracle_osqp_settings* os, double* x, double* y,
                                 oracle_osqp_info* info);
typedef int (*osqp_ref_kkt_solve_fn)(long long n, long long m, const long long* Pp, const long long* Pi,
                                     const double* Px, const long long* Ap, const long long* Ai, const double* Ax,
                                     double sigma, double rho, const double* rhs, double* sol);
#ifdef __cplusplus
}
#endif
#endif
