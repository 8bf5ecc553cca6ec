// placeholder until K2 lands (next commit)
#include "uavmp_internal.h"
int qp_solve_batch_dev(uavmp_ctx* ctx, int, int, int, const double*, const double*, const double*, const double*,
                       const double*, const uavmp_osqp_settings*, double*, int*, int*, int*) {
  return uavmp_fail(ctx, UAVMP_ESTATE, "QP kernel not built yet");
}
int qp_waypoints_from_paths(uavmp_ctx* ctx, int, int, double, const double*, const double*, int, double**, double**,
                            double**, double**, double**) {
  return uavmp_fail(ctx, UAVMP_ESTATE, "QP kernel not built yet");
}
int qp_scatter_plan_outputs(uavmp_ctx* ctx, int, int, int, const int*, const double*, int*, double*) {
  return uavmp_fail(ctx, UAVMP_ESTATE, "QP kernel not built yet");
}
void qp_free_plans(uavmp_ctx*) {}
