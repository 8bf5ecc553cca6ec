// aux_kernels.cu — the two "next" rows of SURVEY.md §8(f) that sit directly on either side of the hot path:
//   K3 map_inflate : GridMap::cloudCallback's inflation, cloud -> occupancy_buffer_inflate_ on the device
//                    (reference: src/planner/plan_env/src/grid_map.cpp:733-785), so the map never crosses PCIe as a grid;
//   K4 polytraj    : PolyTraj::evaluatePos / evaluateVel / evaluateAcc batched over trajectories and sample times
//                    (reference: src/planner/traj_utils/include/traj_utils/poly_traj.hpp:74-168), the direct consumer of
//                    the QP's coefficients (segment-major, ascending power, per-segment local time).
#include <math.h>

#include "uavmp_internal.h"

int uavmp_map_commit(uavmp_ctx* ctx, int nx, int ny, int nz, const double origin[3], const double map_size[3], double resolution,
                     int n_cloud);

namespace {

__global__ void k_inflate(const float* __restrict__ cloud, int n, int8_t* occ, int nx, int ny, int nz, double ox, double oy,
                          double oz, double res, double inv_res, int inf_step) {
  // one thread per (point, dx, dy): the 3 dz stamps are written by the same thread; all writes store the value 1, so the
  // result does not depend on the order (the reference's loop is sequential, grid_map.cpp:761-785)
  const int w = 2 * inf_step + 1;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * w * w) return;
  const int i = (int)(t / (w * w)), r = (int)(t % (w * w));
  const int x = r / w - inf_step, y = r % w - inf_step;
  const float px = cloud[3 * i], py = cloud[3 * i + 1], pz = cloud[3 * i + 2];
  const double qx = px + x * res, qy = py + y * res;  // float + double, like `pt.x + x * mp_.resolution_`
  const int ix = (int)floor((qx - ox) * inv_res), iy = (int)floor((qy - oy) * inv_res);
  if (ix < 0 || iy < 0 || ix > nx - 1 || iy > ny - 1) return;
  for (int z = -1; z <= 1; z++) {  // inf_step_z = 1 is hard-coded in the reference (:736)
    const double qz = pz + z * res;
    const int iz = (int)floor((qz - oz) * inv_res);
    if (iz < 0 || iz > nz - 1) continue;
    occ[((size_t)ix * ny + iy) * nz + iz] = 1;
  }
}

// Eigen 3.3 VectorXd::dot == (a .* b).sum() with the linear vectorised reduction on 2-wide packets
// (Eigen/src/Core/Redux.h, redux_impl<..., LinearVectorizedTraversal, NoUnrolling>); "believed": Eigen is not available here.
__device__ double eigen_dot(const double* a, const double* b, int size) {
  const int aligned2 = (size / 4) * 4, aligned = (size / 2) * 2;
  double res;
  if (aligned > 0) {
    double r00 = a[0] * b[0], r01 = a[1] * b[1];
    if (aligned > 2) {
      double r10 = a[2] * b[2], r11 = a[3] * b[3];
      for (int k = 4; k < aligned2; k += 4) {
        r00 += a[k] * b[k]; r01 += a[k + 1] * b[k + 1];
        r10 += a[k + 2] * b[k + 2]; r11 += a[k + 3] * b[k + 3];
      }
      r00 = r00 + r10; r01 = r01 + r11;
      if (aligned > aligned2) { r00 += a[aligned2] * b[aligned2]; r01 += a[aligned2 + 1] * b[aligned2 + 1]; }
    }
    res = r00 + r01;
    for (int k = aligned; k < size; k++) res += a[k] * b[k];
  } else {
    res = a[0] * b[0];
    for (int k = 1; k < size; k++) res += a[k] * b[k];
  }
  return res;
}

__global__ void k_polytraj(int B, int nc, int S, const double* __restrict__ coef, const double* __restrict__ times, int n_t,
                           const double* __restrict__ ts, int deriv, double* __restrict__ out) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)B * n_t) return;
  const int b = (int)(id / n_t), k = (int)(id % n_t);
  const double* T = times + (size_t)b * S;
  double t = ts[k];
  // segment lookup (poly_traj.hpp:76-88).  The reference tests `t > times[idx] + 1e-4` BEFORE `idx < num_seg` and thereby reads
  // one past the end; here the bound is tested first — identical whenever the reference's read is defined.
  int idx = 0;
  while (idx < S && t > T[idx] + 1e-4) { t -= T[idx]; idx++; }
  if (idx == S) { idx--; t = T[idx]; }
  const int len = nc - deriv;
  double tv[16], cv[16];
  for (int ax = 0; ax < 3; ax++) {
    const double* c = coef + (((size_t)b * 3 + ax) * S + idx) * nc;
    for (int i = 0; i < len; i++) {
      tv[i] = (i == 0) ? 1.0 : tv[i - 1] * t;
      cv[i] = deriv == 0 ? c[i] : (deriv == 1 ? (double)(i + 1) * c[i + 1] : (double)((i + 2) * (i + 1)) * c[i + 2]);
    }
    out[((size_t)b * n_t + k) * 3 + ax] = eigen_dot(tv, cv, len);
  }
}

}  // namespace

extern "C" {

int uavmp_map_set_from_cloud(uavmp_ctx* ctx, const float* cloud_xyz, int n_cloud, int nx, int ny, int nz, const double origin[3],
                             const double map_size[3], double resolution, double obstacles_inflation) {
  if (!ctx || !cloud_xyz || n_cloud <= 0 || nx <= 0 || ny <= 0 || nz <= 0 || !(resolution > 0)) return UAVMP_EINVAL;
  if (nx >= (1 << 17) || ny >= (1 << 17) || nz >= (1 << 17)) return uavmp_fail(ctx, UAVMP_EINVAL, "grid dimension too large");
  {
    int r = uavmp_map_check_geometry(ctx, nx, ny, nz, origin, map_size, resolution);
    if (r) return r;
  }
  cudaSetDevice(ctx->device);
  drain_all(ctx);
  ctx->have_map = false;
  const size_t nvox = (size_t)nx * ny * nz;
  if (ctx->d_occ) { cudaFree(ctx->d_occ); ctx->d_occ = nullptr; }
  if (ctx->d_flags) { cudaFree(ctx->d_flags); ctx->d_flags = nullptr; }
  if (ctx->d_tmp) { cudaFree(ctx->d_tmp); ctx->d_tmp = nullptr; }
  if (ctx->d_cloud) { cudaFree(ctx->d_cloud); ctx->d_cloud = nullptr; }
  UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_occ, nvox));
  UAVMP_CUDA(ctx, cudaMalloc(&ctx->d_cloud, (size_t)n_cloud * 3 * sizeof(float)));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(ctx->d_cloud, cloud_xyz, (size_t)n_cloud * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  UAVMP_CUDA(ctx, cudaMemsetAsync(ctx->d_occ, 0, nvox, ctx->stream));
  const int inf_step = (int)ceil(obstacles_inflation / resolution);  // grid_map.cpp:735
  const long long work = (long long)n_cloud * (2 * inf_step + 1) * (2 * inf_step + 1);
  k_inflate<<<(unsigned)((work + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_cloud, n_cloud, ctx->d_occ, nx, ny, nz, origin[0], origin[1],
                                                                   origin[2], resolution, 1.0 / resolution, inf_step);
  UAVMP_CUDA(ctx, cudaGetLastError());
  return uavmp_map_commit(ctx, nx, ny, nz, origin, map_size, resolution, n_cloud);
}

int uavmp_map_get_occupancy(uavmp_ctx* ctx, int8_t* occ_inflate, long long cap) {
  if (!ctx || !occ_inflate) return UAVMP_EINVAL;
  if (!ctx->have_map) return uavmp_fail(ctx, UAVMP_ESTATE, "no map");
  const long long nvox = (long long)ctx->nx * ctx->ny * ctx->nz;
  if (cap < nvox) return uavmp_fail(ctx, UAVMP_ECAP, "occupancy buffer too small");
  cudaSetDevice(ctx->device);
  UAVMP_CUDA(ctx, cudaMemcpyAsync(occ_inflate, ctx->d_occ, (size_t)nvox, cudaMemcpyDeviceToHost, ctx->stream));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UAVMP_OK;
}

int uavmp_polytraj_eval_batch(uavmp_ctx* ctx, int B, int order, int S, const double* coef, const double* times, int n_t,
                              const double* t, int deriv, double* out) {
  if (!ctx || B <= 0 || S <= 0 || n_t <= 0 || !coef || !times || !t || !out) return UAVMP_EINVAL;
  if (order < 2 || order > 15 || deriv < 0 || deriv > 2) return uavmp_fail(ctx, UAVMP_EINVAL, "order must be 2..15, deriv 0..2");
  cudaSetDevice(ctx->device);
  const int nc = order + 1;
  const size_t nb_c = (size_t)B * 3 * S * nc * sizeof(double), nb_T = (size_t)B * S * sizeof(double), nb_t = (size_t)n_t * sizeof(double),
               nb_o = (size_t)B * n_t * 3 * sizeof(double);
  double *d_c, *d_T, *d_t, *d_o;
  UAVMP_CUDA(ctx, cudaMalloc(&d_c, nb_c + nb_T + nb_t + nb_o));
  d_T = d_c + (size_t)B * 3 * S * nc; d_t = d_T + (size_t)B * S; d_o = d_t + n_t;
  cudaStream_t st = ctx->stream;
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d_c, coef, nb_c, cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d_T, times, nb_T, cudaMemcpyHostToDevice, st));
  UAVMP_CUDA(ctx, cudaMemcpyAsync(d_t, t, nb_t, cudaMemcpyHostToDevice, st));
  const long long work = (long long)B * n_t;
  k_polytraj<<<(unsigned)((work + 127) / 128), 128, 0, st>>>(B, nc, S, d_c, d_T, n_t, d_t, deriv, d_o);
  UAVMP_CUDA(ctx, cudaGetLastError());
  UAVMP_CUDA(ctx, cudaMemcpyAsync(out, d_o, nb_o, cudaMemcpyDeviceToHost, st));
  UAVMP_CUDA(ctx, cudaStreamSynchronize(st));
  cudaFree(d_c);
  return UAVMP_OK;
}

}  // extern "C"
