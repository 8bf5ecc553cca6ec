// capi.cu — the extern "C" boundary declared in include/uavmp.h.  No torch types, plain pointers and sizes.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "uavmp_internal.h"

int kino_fpmath_eval(uavmp_ctx* ctx, int op, int npow, const double* x, double* y, long long n);
// qp_kernel.cu
int qp_solve_batch_dev(uavmp_ctx* ctx, int order, int S, int B, const double* d_pos, const double* d_bv,
                       const double* d_ba, const double* d_bj, const double* d_T, const uavmp_osqp_settings* st,
                       double* d_coef, int* d_solved, int* d_status, int* d_iters);
int qp_waypoints_from_paths(uavmp_ctx* ctx, int B, int S, double seg_time, const double* d_sv, const double* d_ev,
                            int order, double** d_pos, double** d_bv, double** d_ba, double** d_bj, double** d_T);
int qp_scatter_plan_outputs(uavmp_ctx* ctx, int B, int order, int S, const int* d_solved3, const double* d_coef3,
                            int* d_qp_solved, double* d_coef);
void qp_free_plans(uavmp_ctx* ctx);
int qp_launch_fused(uavmp_ctx* ctx, int order, int S, int B, double seg_time, const double* d_sv, const double* d_ev,
                    const int* d_order, const uavmp_osqp_settings* st, double* d_coef, int* d_qp_solved, bool prepare_only);

int ensure_bytes(uavmp_ctx* ctx, void** p, size_t* have, size_t want) {
  if (*have >= want) return UAVMP_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *have = 0;
  UAVMP_CUDA(ctx, cudaMalloc(p, want));
  *have = want;
  return UAVMP_OK;
}

// d_occ / d_cloud are in place: record the geometry and build the derived structures (flag grid, cell list, tensor map)
int uavmp_map_commit(uavmp_ctx* ctx, int nx, int ny, int nz, const double origin[3], const double map_size[3], double resolution,
                     int n_cloud) {
  ctx->nx = nx; ctx->ny = ny; ctx->nz = nz; ctx->n_cloud = n_cloud; ctx->resolution = resolution;
  for (int i = 0; i < 3; i++) { ctx->origin[i] = origin[i]; ctx->map_size[i] = map_size[i]; }
  ctx->have_map = true;
  ctx->flags_dirty = true;
  if (ctx->params_dirty) { int r = kino_upload_params(ctx); if (r) return r; }
  return kino_build_map(ctx);
}

extern "C" {

const char* uavmp_version(void) { return "uavmp-b200 0.1 (sm_100a)"; }

void uavmp_kino_params_default(uavmp_kino_params* p) {  // kino_astar.cpp:8-19
  p->allocated_node_num = 100000; p->collision_check_type = 1; p->rou_time = 1.0; p->lambda_heu = 2.0;
  p->goal_tolerance = 2.0; p->time_step_size = 0.1; p->max_velocity = 5.0; p->max_accelration = 7.0;
  p->acc_resolution = 2.0; p->sample_tau = 0.5; p->robot_r = 0.2; p->robot_h = 0.1;
}
void uavmp_kino_params_launch(uavmp_kino_params* p) {  // test_kino_astar_searching.launch:44-57
  p->allocated_node_num = 100000; p->collision_check_type = 1; p->rou_time = 50.0; p->lambda_heu = 3.0;
  p->goal_tolerance = 2.0; p->time_step_size = 0.075; p->max_velocity = 7.0; p->max_accelration = 10.0;
  p->acc_resolution = 4.0; p->sample_tau = 0.3; p->robot_r = 0.4; p->robot_h = 0.1;
}
void uavmp_osqp_settings_default(uavmp_osqp_settings* s) {
  // osqp_api_constants.h:96-153 with minimum_control.cpp:160-162's overrides
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6; s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-3;
  s->eps_dual_inf = 1e-4; s->max_iter = 1000; s->check_termination = 25; s->scaling = 10; s->adaptive_rho = 1;
  s->adaptive_rho_interval = 0; s->adaptive_rho_tolerance = 5.0;
}

int uavmp_ctx_create(uavmp_ctx** out, int device) {
  if (!out) return UAVMP_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    fprintf(stderr, "uavmp: no CUDA device (%s) — this library has no CPU fallback\n", cudaGetErrorString(e));
    return UAVMP_ECUDA;
  }
  if (device < 0 || device >= ndev) return UAVMP_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return UAVMP_ECUDA;
  uavmp_ctx* ctx = new uavmp_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->sm_count = prop.multiProcessorCount;
  {
    // the context's stream gets the highest priority so that the overlapped QP kernel (lowest priority, second stream) only
    // takes SM space the persistent search CTAs have given up
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, greatest) != cudaSuccess) { delete ctx; return UAVMP_ECUDA; }
  }
  for (int i = 0; i < 8; i++) cudaEventCreate(&ctx->ev[i]);
  uavmp_kino_params_launch(&ctx->kp);
  memset(&ctx->tm, 0, sizeof(ctx->tm));
  memset(&ctx->map_host, 0, sizeof(ctx->map_host));
  *out = ctx;
  return UAVMP_OK;
}

void uavmp_ctx_destroy(uavmp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  void* ptrs[] = {ctx->d_kparams, ctx->d_lattice, ctx->d_occ, ctx->d_flags, ctx->d_tmp, ctx->d_cloud, ctx->d_cell_start,
                  ctx->d_pts, ctx->d_map, ctx->d_arena_mem, ctx->d_arenas, ctx->d_q, ctx->d_order, ctx->d_status,
                  ctx->d_use, ctx->d_npop, ctx->d_hash, ctx->d_npath, ctx->d_path_stage, ctx->d_trace, ctx->d_offsets,
                  ctx->d_path_packed, ctx->d_misc, ctx->d_counters, ctx->d_cub_tmp, ctx->d_qp_ws, ctx->d_qp_in,
                  ctx->d_qp_out, ctx->d_qp_int, ctx->d_plan_out, ctx->d_plan_io, ctx->d_wp, ctx->d_phase, ctx->d_query_cycles, ctx->d_flags_pad, ctx->d_b3f, ctx->d_done_flags};
  for (void* p : ptrs) if (p) cudaFree(p);
  qp_free_plans(ctx);
  for (int i = 0; i < 8; i++) cudaEventDestroy(ctx->ev[i]);
  if (ctx->fuse_ready) { cudaStreamSynchronize(ctx->stream2); cudaEventDestroy(ctx->ev_fuse[0]); cudaEventDestroy(ctx->ev_fuse[1]); cudaStreamDestroy(ctx->stream2); }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* uavmp_last_error(const uavmp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* uavmp_ctx_stream(uavmp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int uavmp_ctx_sync(uavmp_ctx* ctx) {
  if (!ctx) return UAVMP_EINVAL;
  UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UAVMP_OK;
}

int uavmp_kino_set_params(uavmp_ctx* ctx, const uavmp_kino_params* p) {
  if (!ctx || !p) return UAVMP_EINVAL;
  cudaSetDevice(ctx->device);
  ctx->kp = *p;
  ctx->params_dirty = true;
  return kino_upload_params(ctx);
}

int uavmp_map_set(uavmp_ctx* ctx, const int8_t* occ, int nx, int ny, int nz, const double origin[3],
                  const double map_size[3], double resolution, const float* cloud_xyz, int n_cloud) {
  if (!ctx || !occ || nx <= 0 || ny <= 0 || nz <= 0 || !(resolution > 0) || n_cloud < 0) return UAVMP_EINVAL;
  if (n_cloud > 0 && !cloud_xyz) return UAVMP_EINVAL;
  if (nx >= (1 << 17) || ny >= (1 << 17) || nz >= (1 << 17)) return uavmp_fail(ctx, UAVMP_EINVAL, "grid dimension too large");
  cudaSetDevice(ctx->device);
  const size_t nvox = (size_t)nx * ny * nz;
  if (ctx->d_occ) { cudaFree(ctx->d_occ); ctx->d_occ = nullptr; }
  if (ctx->d_flags) { cudaFree(ctx->d_flags); ctx->d_flags = nullptr; }
  if (ctx->d_tmp) { cudaFree(ctx->d_tmp); ctx->d_tmp = nullptr; }
  if (ctx->d_cloud) { cudaFree(ctx->d_cloud); ctx->d_cloud = nullptr; }
  UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_occ, nvox));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(ctx->d_occ, occ, nvox, cudaMemcpyHostToDevice, ctx->stream));
  if (n_cloud > 0) {
    UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_cloud, (size_t)n_cloud * 3 * sizeof(float)));
    UAVMP_CUDA(ctx, cudaMemcpyAsync(ctx->d_cloud, cloud_xyz, (size_t)n_cloud * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  }
  return uavmp_map_commit(ctx, nx, ny, nz, origin, map_size, resolution, n_cloud);
}

static int prepare_search(uavmp_ctx* ctx, int B) {
  if (!ctx->have_map) return uavmp_fail(ctx, UAVMP_ESTATE, "uavmp_map_set has not been called");
  if (ctx->params_dirty) { int r = kino_upload_params(ctx); if (r) return r; }
  if (ctx->flags_dirty) { int r = kino_build_map(ctx); if (r) return r; }
  if (ctx->kp.collision_check_type == 2 && ctx->n_cloud == 0)
    return uavmp_fail(ctx, UAVMP_ESTATE, "collision_check_type 2 needs a cloud");
  int r = kino_ensure_arenas(ctx); if (r) return r;
  return kino_ensure_batch(ctx, B);
}

static int check_error_flag(uavmp_ctx* ctx) {
  int flag = 0;
  UAVMP_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->d_misc, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (flag & 1) return uavmp_fail(ctx, UAVMP_ECAP, "voxel index outside the 18-bit key range");
  if (flag & 2) return uavmp_fail(ctx, UAVMP_ECAP, "path has more than %d nodes", UAVMP_MAXPRIM);
  if (flag & 4) return uavmp_fail(ctx, UAVMP_ECAP, "path longer than path_cap=%d points", ctx->path_cap);
  if (flag & 8) return uavmp_fail(ctx, UAVMP_ECUDA, "overlapped QP gave up waiting for a search to finish");
  return UAVMP_OK;
}

long long uavmp_kino_search_batch(uavmp_ctx* ctx, int B, const double* start_pt, const double* start_vel,
                                  const double* end_pt, const double* end_vel, int* status, int* use_node_num,
                                  long long* path_offsets, uint64_t* pop_hash, int* n_pop) {
  if (!ctx || B <= 0 || !start_pt || !start_vel || !end_pt || !end_vel || !status) return UAVMP_EINVAL;
  cudaSetDevice(ctx->device);
  int r = prepare_search(ctx, B);
  if (r) return r;
  cudaStream_t st = ctx->stream;
  const size_t nb = (size_t)B * 3 * sizeof(double);
  double* d = ctx->d_q;
  cudaEventRecord(ctx->ev[0], st);
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d, start_pt, nb, cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d + 3 * (size_t)B, start_vel, nb, cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d + 6 * (size_t)B, end_pt, nb, cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d + 9 * (size_t)B, end_vel, nb, cudaMemcpyHostToDevice, st));
  cudaEventRecord(ctx->ev[1], st);
  r = kino_launch_search(ctx, B, d, d + 3 * (size_t)B, d + 6 * (size_t)B, d + 9 * (size_t)B, true);
  if (r) return r;
  cudaEventRecord(ctx->ev[2], st);
  r = kino_pack_paths(ctx, B);
  if (r) return r;
  cudaEventRecord(ctx->ev[3], st);
  UAVMP_CUDA(ctx, cudaMemcpyAsync(status, ctx->d_status, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (use_node_num) UAVMP_CUDA(ctx, cudaMemcpyAsync(use_node_num, ctx->d_use, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (path_offsets) UAVMP_CUDA(ctx, cudaMemcpyAsync(path_offsets, ctx->d_offsets, (size_t)(B + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
  if (pop_hash) UAVMP_CUDA(ctx, cudaMemcpyAsync(pop_hash, ctx->d_hash, (size_t)B * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  if (n_pop) UAVMP_CUDA(ctx, cudaMemcpyAsync(n_pop, ctx->d_npop, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  cudaEventRecord(ctx->ev[4], st);
  UAVMP_CUDA(ctx, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&ctx->tm.h2d_ms, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->tm.search_ms, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->tm.path_ms, ctx->ev[2], ctx->ev[3]);
  cudaEventElapsedTime(&ctx->tm.d2h_ms, ctx->ev[3], ctx->ev[4]);
  cudaEventElapsedTime(&ctx->tm.total_ms, ctx->ev[0], ctx->ev[4]);
  ctx->tm.qp_ms = 0; ctx->tm.qp_launches = 0;
  ctx->last_B = B;
  r = check_error_flag(ctx);
  if (r) return r;
  return ctx->last_total_path;
}

int uavmp_kino_get_paths(uavmp_ctx* ctx, double* path_xyz, long long cap_points) {
  if (!ctx || !path_xyz) return UAVMP_EINVAL;
  cudaSetDevice(ctx->device);
  if (cap_points < ctx->last_total_path) return uavmp_fail(ctx, UAVMP_ECAP, "path buffer too small");
  if (ctx->last_total_path > 0)
    UAVMP_CUDA(ctx, cudaMemcpyAsync(path_xyz, ctx->d_path_packed, (size_t)ctx->last_total_path * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UAVMP_OK;
}

int uavmp_kino_set_trace(uavmp_ctx* ctx, int pop_cap) {
  if (!ctx || pop_cap < 0) return UAVMP_EINVAL;
  cudaSetDevice(ctx->device);
  ctx->pop_cap = pop_cap;
  ctx->batch_cap = 0;  // force re-allocation of the batch buffers
  return UAVMP_OK;
}

int uavmp_kino_get_trace(uavmp_ctx* ctx, int q, int32_t* pop_idx_xyz, int cap) {
  if (!ctx || !pop_idx_xyz || q < 0 || q >= ctx->last_B) return UAVMP_EINVAL;
  if (!ctx->d_trace || ctx->pop_cap <= 0) return uavmp_fail(ctx, UAVMP_ESTATE, "tracing is off");
  cudaSetDevice(ctx->device);
  int n = std::min(cap, ctx->pop_cap);
  UAVMP_CUDA(ctx, cudaMemcpyAsync(pop_idx_xyz, ctx->d_trace + (size_t)q * ctx->pop_cap * 3, (size_t)n * 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UAVMP_OK;
}

int uavmp_kino_get_counters(uavmp_ctx* ctx, uavmp_kino_counters* out) {
  if (!ctx || !out || !ctx->d_counters) return UAVMP_EINVAL;
  cudaSetDevice(ctx->device);
  unsigned long long c[8];
  UAVMP_CUDA(ctx, cudaMemcpyAsync(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out->n_pop = c[0]; out->n_occ_lookup = c[1]; out->n_cloud_pts_tested = c[2]; out->n_hash_probe = c[3];
  out->n_insert = c[4]; out->n_update = c[5]; out->n_heuristic = c[6]; out->n_shot = c[7];
  return UAVMP_OK;
}

int uavmp_kino_set_profile(uavmp_ctx* ctx, int on) {
  if (!ctx) return UAVMP_EINVAL;
  ctx->profile_phases = on != 0;
  return UAVMP_OK;
}

int uavmp_kino_get_profile(uavmp_ctx* ctx, unsigned long long phase_cycles[16], long long* query_cycles, int cap, int* grid) {
  if (!ctx || !phase_cycles) return UAVMP_EINVAL;
  if (!ctx->d_phase) return uavmp_fail(ctx, UAVMP_ESTATE, "profiling was off for the last search");
  cudaSetDevice(ctx->device);
  UAVMP_CUDA(ctx, cudaMemcpyAsync(phase_cycles, ctx->d_phase, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  // cap >= 17 * B: the B per-query totals followed by the B x 16 per-query phase cycles; else only the totals
  if (query_cycles && cap > 0) {
    const size_t nq = (cap >= 17 * ctx->last_B) ? (size_t)17 * ctx->last_B : (size_t)std::min(cap, ctx->last_B);
    UAVMP_CUDA(ctx, cudaMemcpyAsync(query_cycles, ctx->d_query_cycles, nq * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  }
  UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (grid) *grid = ctx->last_grid;
  return UAVMP_OK;
}

int uavmp_debug_overlap(uavmp_ctx* ctx, unsigned long long out[4]) {
  if (!ctx || !ctx->dbg_ptr) return UAVMP_ESTATE;
  cudaDeviceSynchronize();
  cudaMemcpy(out, ctx->dbg_ptr, 32, cudaMemcpyDeviceToHost);
  return UAVMP_OK;
}

int uavmp_get_timings(uavmp_ctx* ctx, uavmp_timings* out) {
  if (!ctx || !out) return UAVMP_EINVAL;
  if (ctx->tm_pending_dev) {
    // uavmp_plan_batch_dev is asynchronous: resolve its events (search start / search end / pipeline end) now
    cudaSetDevice(ctx->device);
    UAVMP_CUDA(ctx, cudaEventSynchronize(ctx->ev[7]));
    cudaEventElapsedTime(&ctx->tm.search_ms, ctx->ev[5], ctx->ev[6]);
    cudaEventElapsedTime(&ctx->tm.qp_ms, ctx->ev[6], ctx->ev[7]);
    cudaEventElapsedTime(&ctx->tm.total_ms, ctx->ev[5], ctx->ev[7]);
    ctx->tm.h2d_ms = 0; ctx->tm.d2h_ms = 0; ctx->tm.path_ms = 0;
    ctx->tm_pending_dev = false;
  }
  *out = ctx->tm;
  return UAVMP_OK;
}

int uavmp_fpmath_eval(uavmp_ctx* ctx, int op, int n_pow, const double* x, double* y, long long n) {
  if (!ctx || !x || !y || n <= 0) return UAVMP_EINVAL;
  cudaSetDevice(ctx->device);
  return kino_fpmath_eval(ctx, op, n_pow, x, y, n);
}

// ---- hot path (b) --------------------------------------------------------------------------------------

int uavmp_minctrl_solve_batch(uavmp_ctx* ctx, int order, int S, int B, const double* pos_1d, const double* bound_vel,
                              const double* bound_acc, const double* bound_jerk, const double* time_vec,
                              const uavmp_osqp_settings* settings, double* coef, int* solved, int* osqp_status,
                              int* iters) {
  if (!ctx || B <= 0 || S <= 0 || !pos_1d || !bound_vel || !bound_acc || !time_vec || !coef) return UAVMP_EINVAL;
  if (order != 5 && order != 7) return uavmp_fail(ctx, UAVMP_EINVAL, "order must be 5 (jerk) or 7 (snap)");
  if (order == 7 && !bound_jerk) return uavmp_fail(ctx, UAVMP_EINVAL, "order 7 needs bound_jerk");
  cudaSetDevice(ctx->device);
  uavmp_osqp_settings def;
  if (!settings) { uavmp_osqp_settings_default(&def); settings = &def; }
  cudaStream_t st = ctx->stream;
  const int n = (order + 1) * S;
  const size_t in_d = (size_t)B * ((S + 1) + 6 + S);
  int r = ensure_bytes(ctx, (void**)&ctx->d_qp_in, &ctx->qp_in_bytes, in_d * sizeof(double)); if (r) return r;
  r = ensure_bytes(ctx, (void**)&ctx->d_qp_out, &ctx->qp_out_bytes, (size_t)B * n * sizeof(double)); if (r) return r;
  r = ensure_bytes(ctx, (void**)&ctx->d_qp_int, &ctx->qp_int_bytes, (size_t)B * 3 * sizeof(int)); if (r) return r;
  double* d_pos = ctx->d_qp_in; double* d_bv = d_pos + (size_t)B * (S + 1); double* d_ba = d_bv + 2 * (size_t)B;
  double* d_bj = d_ba + 2 * (size_t)B; double* d_T = d_bj + 2 * (size_t)B;
  cudaEventRecord(ctx->ev[0], st);
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d_pos, pos_1d, (size_t)B * (S + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d_bv, bound_vel, (size_t)B * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d_ba, bound_acc, (size_t)B * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  if (order == 7) UAVMP_CUDA(ctx, cudaMemcpyAsync(d_bj, bound_jerk, (size_t)B * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d_T, time_vec, (size_t)B * S * sizeof(double), cudaMemcpyHostToDevice, st));
  cudaEventRecord(ctx->ev[1], st);
  int* d_solved = ctx->d_qp_int; int* d_stat = d_solved + B; int* d_it = d_stat + B;
  r = qp_solve_batch_dev(ctx, order, S, B, d_pos, d_bv, d_ba, order == 7 ? d_bj : nullptr, d_T, settings, ctx->d_qp_out,
                         d_solved, d_stat, d_it);
  if (r) return r;
  cudaEventRecord(ctx->ev[2], st);
  UAVMP_CUDA(ctx, cudaMemcpyAsync(coef, ctx->d_qp_out, (size_t)B * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (solved) UAVMP_CUDA(ctx, cudaMemcpyAsync(solved, d_solved, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (osqp_status) UAVMP_CUDA(ctx, cudaMemcpyAsync(osqp_status, d_stat, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (iters) UAVMP_CUDA(ctx, cudaMemcpyAsync(iters, d_it, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  cudaEventRecord(ctx->ev[3], st);
  UAVMP_CUDA(ctx, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&ctx->tm.h2d_ms, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->tm.qp_ms, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->tm.d2h_ms, ctx->ev[2], ctx->ev[3]);
  cudaEventElapsedTime(&ctx->tm.total_ms, ctx->ev[0], ctx->ev[3]);
  ctx->tm.search_ms = 0; ctx->tm.path_ms = 0; ctx->tm.search_launches = 0;
  return UAVMP_OK;
}

// ---- pipeline ---------------------------------------------------------------------------------------------
int uavmp_plan_batch_dev(uavmp_ctx* ctx, int B, const double* d_sp, const double* d_sv, const double* d_ep,
                         const double* d_ev, int order, int S, double seg_time, const uavmp_osqp_settings* settings,
                         int* d_search_status, int* d_qp_solved, double* d_coef) {
  if (!ctx || B <= 0 || !d_sp || !d_sv || !d_ep || !d_ev || !d_search_status || !d_qp_solved || !d_coef) return UAVMP_EINVAL;
  if (order != 5 && order != 7) return uavmp_fail(ctx, UAVMP_EINVAL, "order must be 5 or 7");
  cudaSetDevice(ctx->device);
  uavmp_osqp_settings def;
  if (!settings) { uavmp_osqp_settings_default(&def); settings = &def; }
  int r = prepare_search(ctx, B);
  if (r) return r;
  cudaStream_t st = ctx->stream;
  const bool overlap = !getenv("UAVMP_NO_OVERLAP");
  if (overlap) {
    // overlapped pipeline: search on the context's stream, the QP kernel on a second low-priority stream; every QP thread
    // waits for its own query's completion flag, so the QP runs on the SMs the search's long tail leaves idle
    if (!ctx->fuse_ready) {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      UAVMP_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, lo));
      cudaEventCreateWithFlags(&ctx->ev_fuse[0], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ctx->ev_fuse[1], cudaEventDisableTiming);
      ctx->fuse_ready = true;
    }
    if (ctx->done_flags_cap < B) {
      if (ctx->d_done_flags) cudaFree(ctx->d_done_flags);
      ctx->d_done_flags = nullptr;
      UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_done_flags, (size_t)B * sizeof(int)));
      ctx->done_flags_cap = B;
    }
    r = qp_launch_fused(ctx, order, S, B, seg_time, d_sv, d_ev, nullptr, settings, d_coef, d_qp_solved, /*prepare_only=*/true);
    if (r) return r;
    cudaEventRecord(ctx->ev[5], st);
    UAVMP_CUDA(ctx, cudaMemsetAsync(ctx->d_done_flags, 0, (size_t)B * sizeof(int), st));
    // kino_launch_search records ev_fuse[0] right before the search kernel (after the flags are zero and the processing order
    // is sorted): the QP stream waits for that event, NOT for the search kernel
    ctx->fuse_flags = ctx->d_done_flags; ctx->fuse_qp_solved = d_qp_solved;
    r = kino_launch_search(ctx, B, d_sp, d_sv, d_ep, d_ev, true);
    ctx->fuse_flags = nullptr; ctx->fuse_qp_solved = nullptr;
    if (r) return r;
    cudaEventRecord(ctx->ev[6], st);
    UAVMP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fuse[0], 0));
    r = qp_launch_fused(ctx, order, S, B, seg_time, d_sv, d_ev, B > 1 ? ctx->d_order + B : nullptr, settings, d_coef, d_qp_solved, false);
    if (r) return r;
    cudaEventRecord(ctx->ev_fuse[1], ctx->stream2);
    UAVMP_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_fuse[1], 0));  // everything after this call on `st` sees the QP's results
    UAVMP_CUDA(ctx, cudaMemcpyAsync(d_search_status, ctx->d_status, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, st));
    cudaEventRecord(ctx->ev[7], st);
    ctx->last_B = B;
    ctx->tm_pending_dev = true;
    ctx->tm.aux_launches = 1;  // k_dist_keys (+ cub's radix-sort kernels, library code)
    return UAVMP_OK;
  }
  cudaEventRecord(ctx->ev[5], st);
  r = kino_launch_search(ctx, B, d_sp, d_sv, d_ep, d_ev, true);
  if (r) return r;
  cudaEventRecord(ctx->ev[6], st);
  double *w_pos, *w_bv, *w_ba, *w_bj, *w_T;
  r = qp_waypoints_from_paths(ctx, B, S, seg_time, d_sv, d_ev, order, &w_pos, &w_bv, &w_ba, &w_bj, &w_T);
  if (r) return r;
  const int n = (order + 1) * S;
  r = ensure_bytes(ctx, (void**)&ctx->d_qp_out, &ctx->qp_out_bytes, (size_t)3 * B * n * sizeof(double)); if (r) return r;
  r = ensure_bytes(ctx, (void**)&ctx->d_qp_int, &ctx->qp_int_bytes, (size_t)3 * B * 3 * sizeof(int)); if (r) return r;
  int* d_solved3 = ctx->d_qp_int; int* d_stat3 = d_solved3 + 3 * B; int* d_it3 = d_stat3 + 3 * B;
  r = qp_solve_batch_dev(ctx, order, S, 3 * B, w_pos, w_bv, w_ba, order == 7 ? w_bj : nullptr, w_T, settings,
                         ctx->d_qp_out, d_solved3, d_stat3, d_it3);
  if (r) return r;
  r = qp_scatter_plan_outputs(ctx, B, order, S, d_solved3, ctx->d_qp_out, d_qp_solved, d_coef);
  if (r) return r;
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d_search_status, ctx->d_status, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, st));
  cudaEventRecord(ctx->ev[7], st);
  ctx->last_B = B;
  ctx->tm_pending_dev = true;
  ctx->tm.aux_launches = 3;  // k_dist_keys, k_waypoints, k_scatter_plan (+ cub's radix-sort kernels, library code)
  return UAVMP_OK;
}

int uavmp_plan_batch(uavmp_ctx* ctx, int B, const double* start_pt, const double* start_vel, const double* end_pt,
                     const double* end_vel, int order, int S, double seg_time, const uavmp_osqp_settings* settings,
                     int* search_status, int* qp_solved, double* coef) {
  if (!ctx || B <= 0 || !start_pt || !start_vel || !end_pt || !end_vel || !search_status || !qp_solved || !coef) return UAVMP_EINVAL;
  cudaSetDevice(ctx->device);
  int r = prepare_search(ctx, B);
  if (r) return r;
  cudaStream_t st = ctx->stream;
  const size_t nb = (size_t)B * 3 * sizeof(double);
  const int n = (order + 1) * S;
  double* d = ctx->d_q;
  r = ensure_bytes(ctx, (void**)&ctx->d_plan_out, &ctx->plan_out_bytes, (size_t)B * 3 * n * sizeof(double)); if (r) return r;
  r = ensure_bytes(ctx, (void**)&ctx->d_plan_io, &ctx->plan_io_bytes, (size_t)B * 2 * sizeof(int)); if (r) return r;
  double* d_out = ctx->d_plan_out;
  int* d_io = ctx->d_plan_io;
  cudaEventRecord(ctx->ev[0], st);
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d, start_pt, nb, cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d + 3 * (size_t)B, start_vel, nb, cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d + 6 * (size_t)B, end_pt, nb, cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d + 9 * (size_t)B, end_vel, nb, cudaMemcpyHostToDevice, st));
  cudaEventRecord(ctx->ev[1], st);
  r = uavmp_plan_batch_dev(ctx, B, d, d + 3 * (size_t)B, d + 6 * (size_t)B, d + 9 * (size_t)B, order, S, seg_time, settings,
                           d_io, d_io + B, d_out);
  if (r) return r;
  UAVMP_CUDA(ctx, cudaMemcpyAsync(search_status, d_io, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(qp_solved, d_io + B, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(coef, d_out, (size_t)B * 3 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  cudaEventRecord(ctx->ev[4], st);
  UAVMP_CUDA(ctx, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&ctx->tm.h2d_ms, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->tm.search_ms, ctx->ev[5], ctx->ev[6]);
  cudaEventElapsedTime(&ctx->tm.qp_ms, ctx->ev[6], ctx->ev[7]);
  cudaEventElapsedTime(&ctx->tm.d2h_ms, ctx->ev[7], ctx->ev[4]);
  cudaEventElapsedTime(&ctx->tm.total_ms, ctx->ev[0], ctx->ev[4]);
  ctx->tm.path_ms = 0;
  ctx->tm_pending_dev = false;
  return check_error_flag(ctx);
}

}  // extern "C"
