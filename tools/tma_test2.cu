#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait0(unsigned long long* mbar) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(mbar)), "r"(0) : "memory");
  } while (!done);
}
// variant 0: non-tensor bulk copy (UBLKCP); 1: tensor 2d with cache hint; 2: tensor 2d issued by an elected lane of a converged warp
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const uint8_t* src, int variant, uint8_t* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned long long* mbar = (unsigned long long*)(sm + 4096);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (variant == 0) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1024) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(sm)), "l"(src), "r"(1024), "r"(smem_u32(mbar)) : "memory");
    }
  } else if (variant == 1) {
    if (threadIdx.x == 0) {
      unsigned long long hint = 0x1000000000000000ull;  // EVICT_NORMAL
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1024) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                   ::"r"(smem_u32(sm)), "l"(&tmap), "r"(smem_u32(mbar)), "r"(0), "r"(0), "l"(hint) : "memory");
    }
  } else {
    if (threadIdx.x < 32) {
      uint32_t pred = 0;
      asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
      if (pred) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1024) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sm)), "l"(&tmap), "r"(smem_u32(mbar)), "r"(0), "r"(0) : "memory");
      }
    }
  }
  wait0(mbar);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0, cluster = argc > 2 ? atoi(argv[2]) : 0;
  size_t n = 1 << 20;
  uint8_t* h = (uint8_t*)malloc(n);
  for (size_t i = 0; i < n; i++) h[i] = (uint8_t)(i * 7 + 3);
  uint8_t *d, *dout; cudaMalloc(&d, n); cudaMemcpy(d, h, n, cudaMemcpyHostToDevice); cudaMalloc(&dout, 1024);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap tm;
  cuuint64_t dims[2] = {1024, 1024}; cuuint64_t strides[1] = {1024}; cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d cluster %d encode %d\n", variant, cluster, (int)r);
  if (cluster) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(1); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 8192;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, probe, tm, (const uint8_t*)d, variant, dout);
  } else {
    probe<<<1, 256, 8192>>>(tm, d, variant, dout);
  }
  cudaError_t le = cudaDeviceSynchronize();
  printf("  -> %s\n", cudaGetErrorString(le));
  if (le == cudaSuccess) {
    uint8_t ho[1024]; cudaMemcpy(ho, dout, 1024, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 1024; i++) { uint8_t want = variant == 0 ? h[i] : h[(i / 32) * 1024 + (i % 32)]; bad += ho[i] != want; }
    printf("  mismatches %d\n", bad);
  }
  return 0;
}
