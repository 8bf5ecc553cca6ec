/* Restated from 3rd/osqp/algebra/_common/lin_sys/qdldl/qdldl_codegen_types.h.in: QDLDL uses OSQP's types. */
#ifndef QDLDL_TYPES_H
#define QDLDL_TYPES_H
#include <limits.h>
#include "osqp_api_types.h"
typedef OSQPInt QDLDL_int;
typedef OSQPFloat QDLDL_float;
typedef int QDLDL_bool;
#ifdef OSQP_USE_LONG
#define QDLDL_INT_MAX LLONG_MAX
#else
#define QDLDL_INT_MAX INT_MAX
#endif
#endif
