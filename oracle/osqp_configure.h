/* oracle/osqp_configure.h — hand-written stand-in for the file OSQP's CMake would generate from
 * 3rd/osqp/configure/osqp_configure.h.in.  Choices and why:
 *   OSQP_ALGEBRA_BUILTIN, OSQP_USE_LONG : the reference's defaults (3rd/osqp/CMakeLists.txt:77-78,161-162)
 *   OSQP_ENABLE_PROFILING  OFF : the default ON makes the adaptive-rho interval wall-clock dependent
 *                                (osqp_api.c:610-636); OFF pins it to 4 x check_termination (osqp_api.c:394-406)
 *   OSQP_ENABLE_PRINTING   OFF : no stdout on the timed path; does not change iterates
 *   OSQP_ENABLE_INTERRUPT  OFF, derivatives / codegen OFF : unused on this path */
#ifndef OSQP_CONFIGURE_H
#define OSQP_CONFIGURE_H
#define IS_LINUX
#define OSQP_ALGEBRA_BUILTIN
#define OSQP_USE_LONG
#endif
