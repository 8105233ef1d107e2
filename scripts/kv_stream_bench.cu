// Microbenchmark (GPU box): what HBM read bandwidth does the decode attention's access pattern allow?
// Every CTA streams `keys` rows of `row_bytes` through a 4-stage ring of shared memory with cp.async.bulk and frees each
// stage as soon as it lands (no compute).  Patterns over a [L][B][T=64][D=1536] bf16 cache:
//   mode 0: (image, group of 6 heads) items - 768-byte pieces at a 3072-byte stride (today's cache layout), K then V
//   mode 1: same bytes per CTA, but each item's rows contiguous (a group-major cache layout)
//   mode 2: (image) items - 3072-byte rows, contiguous, 256 CTAs
//   mode 6: 768-byte pieces of a time-major cache [T][B][D] (a step's keys span t/64 of the pages)
//   mode 5: plain coalesced 16-byte loads of the same total bytes (grid-stride), as the read-only reference
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/kv_stream_bench.bin scripts/kv_stream_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t d;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(d) : "r"(smem_u32(b)), "r"(par) : "memory");
  return d != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) { while (!mbar_try(b, par)) {} }
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
constexpr int STAGES = 4;
// item: rows `keys` x row_bytes; piece r of pass p (0 = K, 1 = V) at base[p] + item_off + r * stride
__global__ void __launch_bounds__(64) stream_kernel(const uint8_t* K, const uint8_t* V, int keys, int row_bytes, size_t stride,
                                                    size_t item_stride, size_t group_stride, int groups, int ch, int one_copy) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = (uint64_t*)(smem + STAGES * ch * row_bytes);
  uint64_t* empty = full + STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const size_t off = (size_t)b * item_stride + (size_t)g * group_stride;
  const int nck = (keys + ch - 1) / ch;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * nck; ++i) {
      const int s = i % STAGES;
      mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
      const int ck = i < nck ? i : i - nck;
      const int rows = (keys - ck * ch) < ch ? (keys - ck * ch) : ch;
      mbar_expect(&full[s], rows * row_bytes);
      const uint8_t* src = (i < nck ? K : V) + off + (size_t)ck * ch * stride;
      if (one_copy) bulk(smem + s * ch * row_bytes, src, rows * row_bytes, &full[s]);
      else for (int r = 0; r < rows; ++r) bulk(smem + s * ch * row_bytes + r * row_bytes, src + r * stride, row_bytes, &full[s]);
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < 2 * nck; ++i) {
      const int s = i % STAGES;
      mbar_wait(&full[s], (i / STAGES) & 1);
      mbar_arrive(&empty[s]);
    }
  }
}
__global__ void ldg_kernel(const uint4* p, size_t n, uint4* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = __ldg(p + i);
    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  if (acc.x == 0x12345678 && acc.y == 0x9abcdef) *sink = acc;
}
int main() {
  const int L = 12, B = 256, T = 64, D = 1536;
  const size_t layer = (size_t)B * T * D * 2;
  uint8_t *K, *V; uint4* sink;
  CK(cudaMalloc(&K, layer * L)); CK(cudaMalloc(&V, layer * L)); CK(cudaMalloc(&sink, 16));
  CK(cudaMemset(K, 1, layer * L)); CK(cudaMemset(V, 2, layer * L));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int big = 0; big < 2; ++big)
  for (int keys : {16, 32, 64}) {
    const int Bq = big ? B * L : B;   // big: all 12 layer slabs as one launch (steady-state streaming rate)
    for (int mode = 0; mode < 7; ++mode) {
      float best = 1e9, sum = 0;
      const int iters = 24;
      for (int it = 0; it < iters + 3; ++it) {
        const uint8_t* k = K + (big ? 0 : (size_t)(it % L) * layer); const uint8_t* v = V + (big ? 0 : (size_t)(it % L) * layer);
        cudaEventRecord(e0);
        if (mode == 0) stream_kernel<<<Bq * 4, 64, STAGES * 8 * 768 + 64>>>(k, v, keys, 768, 3072, (size_t)T * 3072, 768, 4, 8, 0);
        else if (mode == 1) stream_kernel<<<Bq * 4, 64, STAGES * 8 * 768 + 64>>>(k, v, keys, 768, 768, (size_t)T * 3072, (size_t)T * 768, 4, 8, 1);
        else if (mode == 2) stream_kernel<<<Bq, 64, STAGES * 8 * 3072 + 64>>>(k, v, keys, 3072, 3072, (size_t)T * 3072, 0, 1, 8, 1);
        else if (mode == 3) stream_kernel<<<Bq * 4, 64, STAGES * 16 * 768 + 64>>>(k, v, keys, 768, 3072, (size_t)T * 3072, 768, 4, 16, 0);
        else if (mode == 4) stream_kernel<<<Bq * 2, 64, STAGES * 8 * 1536 + 64>>>(k, v, keys, 1536, 3072, (size_t)T * 3072, 1536, 2, 8, 0);
        else if (mode == 6) stream_kernel<<<Bq * 4, 64, STAGES * 8 * 768 + 64>>>(k, v, keys, 768, (size_t)Bq * 3072, 3072, 768, 4, 8, 0);
        else { ldg_kernel<<<148 * 8, 512>>>((const uint4*)k, (size_t)Bq * keys * D * 2 / 16, sink); ldg_kernel<<<148 * 8, 512>>>((const uint4*)v, (size_t)Bq * keys * D * 2 / 16, sink); }
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 3) { sum += ms; if (ms < best) best = ms; }
      }
      const double bytes = 2.0 * Bq * keys * D * 2;
      const char* names[7] = {"768B @3072 stride, 1024 CTAs, 8-key stages", "contiguous items (group-major), 1024 CTAs, 1 copy/stage", "3072B rows, 256 CTAs, 1 copy/stage",
                              "768B @3072 stride, 1024 CTAs, 16-key stages", "1536B @3072 stride, 512 CTAs", "coalesced LDG.128 x2 kernels", "768B pieces, TIME-major cache [T][B][D], 1024 CTAs"};
      printf("B %4d keys %2d mode %d (%s): mean %.2f us min %.2f us -> %.0f GB/s (min-time %.0f)\n", Bq, keys, mode, names[mode], sum / iters * 1e3, best * 1e3,
             bytes / (sum / iters * 1e-3) * 1e-9, bytes / (best * 1e-3) * 1e-9);
    }
  }
  return 0;
}
