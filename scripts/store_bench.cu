// Microbenchmark (GPU box): what does the GEMM epilogue's write-out cost?  Every CTA (8 warps, one CTA per SM) writes a
// [128 rows x cols] tile of a row-major [256, N] matrix the way epilogue_tile does - per warp and 32-column chunk, 8 store
// instructions of 4 rows x 8 lanes - and thread 0 of each warp clocks the ISSUE of those stores.  Variants: bf16 rows (64 B
// pieces), fp32 rows (128 B pieces), and the same bytes as ONE TMA tensor store per warp-chunk from shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/store_bench.bin scripts/store_bench.cu -lcuda
#include "../hqtransformer_b200/csrc/common.cuh"
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
using namespace hq;

// mode 0: st.global 8 B per lane (bf16 piece), 1: st.global 16 B per lane (fp32 piece), 2: TMA tensor store fp32 32x32 box
__global__ void __launch_bounds__(256, 1)
store_kernel(const __grid_constant__ CUtensorMap tmO, uint8_t* out, int N_bytes_per_row, int chunks, int mode, long long* stat) {
#if defined(__CUDA_ARCH__)
  __shared__ __align__(1024) uint8_t slab[8][4096];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int tile_col0 = blockIdx.x * chunks * 2;          // this CTA's first 32-column chunk (both halves: 2 * chunks)
  for (int i = lane; i < 1024; i += 32) reinterpret_cast<float*>(slab[warp])[i] = 1.0f;
  __syncthreads();
  const long long t0 = clock64();
  for (int c = 0; c < chunks; ++c) {
    const int chunk = tile_col0 + c * 2 + half;
    if (mode <= 1) {
      const int piece_bytes = mode == 0 ? 8 : 16;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = (blockIdx.y * 128) + quarter * 32 + it * 4 + (lane >> 3);
        uint8_t* dst = out + (size_t)row * N_bytes_per_row + (size_t)chunk * 32 * (mode == 0 ? 2 : 4) + (lane & 7) * piece_bytes;
        const float4 v = *reinterpret_cast<const float4*>(slab[warp] + (it * 4 + (lane >> 3)) * 128 + (lane & 7) * 16);
        if (mode == 0) *reinterpret_cast<float2*>(dst) = make_float2(v.x, v.y);
        else *reinterpret_cast<float4*>(dst) = v;
      }
    } else {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (elect_one()) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(smem_u32(slab[warp])), "r"(chunk * 32),
                       "r"((int)(blockIdx.y * 128) + quarter * 32)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
    }
  }
  const long long t1 = clock64();
  if (mode == 2 && elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  const long long t2 = clock64();
  if (lane == 0) { atomicAdd((unsigned long long*)&stat[0], (unsigned long long)(t1 - t0)); atomicAdd((unsigned long long*)&stat[1], (unsigned long long)(t2 - t0)); }
#endif
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int N = 6144, M = 256;
  uint8_t* out;
  CK(cudaMalloc(&out, (size_t)M * N * 4));
  long long* stat;
  CK(cudaMalloc(&stat, 16));
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  PFN_encodeTiled fn = reinterpret_cast<PFN_encodeTiled>(p);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)N * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t el[2] = {1, 1};
  if (fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("# grid 72 x 2 CTAs (one per SM), 8 warps, `chunks` 32-column chunks per warp; clk = mean per warp\n");
  printf("# %-22s %6s %12s %14s %10s\n", "mode", "chunks", "issue clk", "issue+drain clk", "kernel us");
  const char* names[3] = {"st.global 8 B (bf16)", "st.global 16 B (fp32)", "TMA store 32x32 fp32"};
  for (int mode = 0; mode < 3; ++mode)
    for (int chunks = 1; chunks <= 2; ++chunks) {
      const int cols_per_cta = chunks * 2 * 32;           // columns (elements) per CTA
      const int grid_x = 72 < N / cols_per_cta ? 72 : N / cols_per_cta;
      float best = 1e30f;
      long long h[2] = {0, 0};
      for (int rep = 0; rep < 5; ++rep) {
        CK(cudaMemset(stat, 0, 16));
        CK(cudaEventRecord(e0));
        store_kernel<<<dim3(grid_x, 2), 256>>>(tm, out, mode == 0 ? N * 2 : N * 4, chunks, mode, stat);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
        CK(cudaMemcpy(h, stat, 16, cudaMemcpyDeviceToHost));
      }
      CK(cudaGetLastError());
      const double nw = (double)grid_x * 2 * 8;
      printf("  %-22s %6d %12.0f %14.0f %10.2f\n", names[mode], chunks, h[0] / nw, h[1] / nw, best * 1e3);
    }
  return 0;
}
