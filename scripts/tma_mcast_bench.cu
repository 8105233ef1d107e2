// Microbenchmark (GPU box): how many bytes per second can one SM ingest through TMA, and does cluster multicast
// raise it?  Every CTA streams a [rows x 64] bf16 box per iteration through a 4-stage ring.
//   mode 0: unicast  - each CTA loads its own 128-row box (all CTAs read the same 2 tiles -> L2 hits)
//   mode 1: multicast- cluster of cs: each CTA loads 128/cs rows and multicasts them to all cs CTAs (each receives 128 rows)
//   mode 2: unicast, 64-row boxes (half the bytes) for reference
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_mcast_bench scripts/tma_mcast_bench.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t d;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(d) : "r"(smem_u32(b)), "r"(par) : "memory");
  return d != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) { while (!mbar_try(b, par)) {} }
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* b, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(b)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

constexpr int STAGES = 8;

__global__ void __launch_bounds__(64)
bench_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
             const __grid_constant__ CUtensorMap map32, const __grid_constant__ CUtensorMap map16, int iters, int mode,
             int num_kb, int cs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(base + STAGES * 16384);
  uint64_t* empty = full + STAGES;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], mode == 1 ? cs : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % STAGES;
      mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
      const int kb = i % num_kb;
      if (mode == 0) {
        mbar_expect(&full[s], 16384);
        tma_load(base + s * 16384, &map128, &full[s], kb * 64, rank * 128);
      } else if (mode == 2) {
        mbar_expect(&full[s], 8192);
        tma_load(base + s * 16384, &map64, &full[s], kb * 64, rank * 64);
      } else {
        mbar_expect(&full[s], 16384);          // this CTA receives all cs slices
        const int rows = 128 / cs;
        const CUtensorMap* mp = cs == 2 ? &map64 : (cs == 4 ? &map32 : &map16);
        tma_load_mc(base + s * 16384 + rank * rows * 128, mp, &full[s], kb * 64, rank * rows, (uint16_t)((1u << cs) - 1));
      }
    }
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % STAGES;
      mbar_wait(&full[s], (i / STAGES) & 1);
      if (mode == 1) { for (int c = 0; c < cs; ++c) mbar_arrive_cluster(&empty[s], c); }
      else mbar_arrive_cluster(&empty[s], rank);
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows = 256, K = 1536, num_kb = K / 64;
  void* A;
  CK(cudaMalloc(&A, (size_t)rows * K * 2));
  CK(cudaMemset(A, 0, (size_t)rows * K * 2));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  PFN enc = (PFN)fn;
  CUtensorMap m128, m64, m32, m16;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t estr[2] = {1, 1}, b128[2] = {64, 128}, b64[2] = {64, 64}, b32[2] = {64, 32}, b16[2] = {64, 16};
  enc(&m128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, strides, b128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&m64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, strides, b64, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&m32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, strides, b32, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&m16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, strides, b16, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int smem = STAGES * 16384 + 1024 + 256;
  CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2000;
  auto launch = [&](int grid, int cs, int it, int mode) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, bench_kernel, m128, m64, m32, m16, it, mode, num_kb, cs));
  };
  for (int grid : {8, 72, 144}) {
    for (int cs : {2, 4, 8}) {
      for (int mode : {0, 1, 2}) {
        if (mode != 1 && cs != 2) continue;
        launch(grid, cs, 100, mode);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        launch(grid, cs, iters, mode);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double recv = (mode == 2 ? 8192.0 : 16384.0) * iters;
        printf("grid %3d cluster %d mode %d (%s): %.1f us, received per CTA %.1f GB/s, chip %.2f TB/s\n", grid, cs, mode,
               mode == 0 ? "unicast 128 rows" : (mode == 1 ? "multicast 128/cs rows each" : "unicast 64 rows"), ms * 1e3,
               recv / (ms * 1e-3) * 1e-9, recv * grid / (ms * 1e-3) * 1e-12);
      }
    }
  }
  return 0;
}
