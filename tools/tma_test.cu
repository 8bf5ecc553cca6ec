// standalone probe: 3-D uint8 TMA box load, descriptor in param space vs global memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <dlfcn.h>
#define TB 32
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, int mode, int z0, int y0, int x0, uint8_t* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned long long* mbar = (unsigned long long*)(sm + TB * TB * TB);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const void* d = (const void*)gmap;
    if (mode == 0) d = (const void*)&tmap;
    if (mode == 2) { asm volatile("prefetch.tensormap [%0];" :: "l"(gmap) : "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(TB * TB * TB) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(sm)), "l"(d), "r"(z0), "r"(y0), "r"(x0), "r"(smem_u32(mbar)) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(mbar)), "r"(0) : "memory");
  } while (!done);
  for (int i = threadIdx.x; i < TB * TB * TB; i += blockDim.x) out[i] = sm[i];
}
__global__ void probe2d(const __grid_constant__ CUtensorMap tmap, int c0, int c1, uint8_t* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned long long* mbar = (unsigned long long*)(sm + 32 * 32);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(32 * 32) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(sm)), "l"(&tmap), "r"(c0), "r"(c1), "r"(smem_u32(mbar)) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(mbar)), "r"(0) : "memory");
  } while (!done);
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
  int want_mode = argc > 1 ? atoi(argv[1]) : 0, want_promo = argc > 2 ? atoi(argv[2]) : 0;
  int nx = 200, ny = 200, nz = 50, nzp = 64;
  size_t n = (size_t)nx * ny * nzp;
  uint8_t* h = (uint8_t*)malloc(n);
  for (size_t i = 0; i < n; i++) h[i] = (uint8_t)((i * 2654435761u) >> 24);
  uint8_t *d, *dout; cudaMalloc(&d, n + 256); cudaMemcpy(d, h, n, cudaMemcpyHostToDevice); cudaMalloc(&dout, TB * TB * TB);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry point: %d %p %d\n", (int)e, fn, (int)q);
  if (argc > 3) {
    void* h2 = dlopen("libcuda.so.1", RTLD_NOW);
    void* f2 = h2 ? dlsym(h2, "cuTensorMapEncodeTiled") : nullptr;
    printf("dlsym: %p %p\n", h2, f2);
    if (f2) fn = f2;
  }
  if (want_mode == 3) {
    alignas(64) CUtensorMap tm;
    memset(&tm, 0xab, sizeof(tm));
    cuuint64_t dims[2] = {(cuuint64_t)nzp, (cuuint64_t)ny * nx};
    cuuint64_t strides[1] = {(cuuint64_t)nzp};
    cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("2d encode result %d\n", (int)r);
    for (int i = 0; i < 16; i++) printf("%016llx ", ((unsigned long long*)&tm)[i]);
    printf("\n");
    probe2d<<<1, 256, 32 * 32 + 64>>>(tm, argc > 4 ? atoi(argv[4]) : 8, 100, dout);
    cudaError_t le = cudaDeviceSynchronize();
    printf("2d: %s\n", cudaGetErrorString(le));
    return le != cudaSuccess;
  }
  for (int promo = want_promo; promo <= want_promo; promo++) {
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)nzp, (cuuint64_t)ny, (cuuint64_t)nx};
    cuuint64_t strides[2] = {(cuuint64_t)nzp, (cuuint64_t)nzp * ny};
    cuuint32_t box[3] = {TB, TB, TB}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("promo %d encode result %d\n", promo, (int)r);
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(tm)); cudaMemcpy(gm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, TB * TB * TB + 64);
    for (int mode = want_mode; mode <= want_mode; mode++) {
      int x0 = 37, y0 = 181, z0 = 41;
      probe<<<1, 256, TB * TB * TB + 64>>>(tm, gm, mode, z0, y0, x0, dout);
      cudaError_t le = cudaDeviceSynchronize();
      printf("promo %d mode %d (%s): %s\n", promo, mode, mode ? "global desc" : "param desc", cudaGetErrorString(le));
      if (le != cudaSuccess) return 1;
      uint8_t* ho = (uint8_t*)malloc(TB * TB * TB); cudaMemcpy(ho, dout, TB * TB * TB, cudaMemcpyDeviceToHost);
      long bad = 0;
      for (int x = 0; x < TB; x++) for (int y = 0; y < TB; y++) for (int z = 0; z < TB; z++) {
        int gx = x0 + x, gy = y0 + y, gz = z0 + z;
        uint8_t want = (gx < nx && gy < ny && gz < nzp) ? h[((size_t)gx * ny + gy) * nzp + gz] : 0;
        if (ho[(x * TB + y) * TB + z] != want) bad++;
      }
      printf("   mismatches %ld\n", bad);
    }
  }
  return 0;
}
