// Microbenchmark (GPU box): how many bytes per second can ONE SM take in through TMA, and does a tile that arrives by
// cluster MULTICAST count against the same limit as one the SM requested itself?
//
// This is the question behind the M = 256 GEMMs of the sampling loop: every CTA of a pair must stage its
// 128 rows of the activation matrix A for the whole K range (16 KB per 64-wide k-block) whatever the tiling, all CTAs read
// the SAME rows, and the measured main-loop rate is ~78 GB/s per CTA.  If multicast delivery were not bound by that rate,
// sharing the A tile across a cluster of CTA pairs would shorten the main loop.
//
// Every CTA streams `iters` k-blocks of a [rows, K] bf16 matrix (L2 resident) through a ring of shared-memory stages with
// cp.async.bulk.tensor (box 64 columns x 128 rows, SWIZZLE_128B, exactly the GEMM's A tile) and frees a stage as soon as it
// has landed (no MMA).  Modes:
//   unicast   (cs = 1)          : each CTA requests its whole 16 KB tile
//   multicast (cs = 2 / 4 / 8)  : the CTAs of a cluster share the tile: CTA r requests rows [r * 128 / cs, (r + 1) * 128 / cs)
//                                 with .multicast::cluster to all cs CTAs, so each CTA RECEIVES 16 KB and REQUESTS 16 / cs KB
//   + w_rows                    : each CTA also streams its own private w_rows x 64 box per k-block (the GEMM's W half-tile)
// Reported: GB/s RECEIVED per CTA and summed over the grid.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/sm_ingest_bench.bin scripts/sm_ingest_bench.cu -lcuda
#include "../hqtransformer_b200/csrc/common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
using namespace hq;

constexpr int MAXSTAGES = 24;

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
#endif

// tmA: box {64, 128 / cs}; tmW: box {64, w_rows} (ignored when w_rows == 0)
__global__ void __launch_bounds__(64, 1)
ingest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int cs, int w_rows, int kblocks,
              int iters, int shared_rows, int m_tiles, int STAGES, int a_rows, long long* tstat) {
#if defined(__CUDA_ARCH__)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int A_BYTES = a_rows * 128;
  const int w_bytes = w_rows * 128;
  const int stage_bytes = A_BYTES + ((w_bytes + 1023) & ~1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* empty = full + MAXSTAGES;
  const uint32_t rank = cs > 1 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    if (w_rows) tma_prefetch_desc(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], cs);       // the consumer of every CTA of the cluster
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (cs > 1) cluster_sync_all();
  const int cluster_id = blockIdx.x / cs;
  const int piece_rows = a_rows / cs;
  // STAGES is a power of two here: slot / phase by mask and shift, k-block by a wrapping counter (no integer division in
  // the loops: the point is to time the barrier / TMA instructions themselves)
  const int smask = STAGES - 1, sshift = 31 - __clz(STAGES);
  if (threadIdx.x == 0) {
    // producer
    const int mt = shared_rows ? 0 : cluster_id % m_tiles;
    long long t_wait = 0, t_exp = 0, t_tma = 0;
    int kb = 0;
    for (int i = 0; i < iters; ++i) {
      const int s = i & smask;
      const long long c0 = clock64();
      mbar_wait(&empty[s], ((i >> sshift) & 1) ^ 1);
      const long long c1 = clock64();
      mbar_arrive_expect_tx(&full[s], A_BYTES + w_bytes);
      const long long c2 = clock64();
      uint8_t* dst = smem + s * stage_bytes;
      if (cs == 1) tma_load_2d(dst, &tmA, &full[s], kb * 64, mt * a_rows);
      else tma_load_2d_mc(dst + rank * piece_rows * 128, &tmA, &full[s], kb * 64, mt * a_rows + rank * piece_rows,
                          static_cast<uint16_t>((1u << cs) - 1u));
      if (w_rows) tma_load_2d(dst + A_BYTES, &tmW, &full[s], kb * 64, (blockIdx.x * w_rows) & 4095);
      const long long c3 = clock64();
      t_wait += c1 - c0; t_exp += c2 - c1; t_tma += c3 - c2;
      kb = kb + 1 == kblocks ? 0 : kb + 1;
    }
    if (blockIdx.x == 0 && tstat) { tstat[0] = t_wait; tstat[1] = t_exp; tstat[2] = t_tma; }
  } else if (threadIdx.x >= 32) {
    // consumer warp: frees a stage in every CTA of the cluster as soon as it has landed here (lane c signals CTA c)
    const int lane = threadIdx.x - 32;
    long long t_wait = 0, t_arr = 0;
    for (int i = 0; i < iters; ++i) {
      const int s = i & smask;
      const long long c0 = clock64();
      mbar_wait(&full[s], (i >> sshift) & 1);
      const long long c1 = clock64();
      if (cs == 1) { if (lane == 0) mbar_arrive(&empty[s]); }
      else if (lane < cs) mbar_arrive_cta(&empty[s], lane);
      __syncwarp();
      const long long c2 = clock64();
      t_wait += c1 - c0; t_arr += c2 - c1;
    }
    if (blockIdx.x == 0 && lane == 0 && tstat) { tstat[3] = t_wait; tstat[4] = t_arr; }
  }
  __syncthreads();
  if (cs > 1) cluster_sync_all();     // no CTA exits while a peer may still signal its barriers
#endif
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(PFN_encodeTiled fn, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t el[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
  return m;
}

static PFN_encodeTiled g_fn;
static void *g_A, *g_W;
static long long* g_t;
static cudaEvent_t e0, e1;
static void run(int cs, int grid, int w_rows, int shared_rows, int stages, int a_rows, int iters) {
  const int K = 1536, rows = 4096, kblocks = K / 64;
  CUtensorMap tmA = make_map(g_fn, g_A, rows, K, a_rows / cs);
  CUtensorMap tmW = make_map(g_fn, g_W, 4096, K, w_rows ? w_rows : 8);
  const int stage_bytes = a_rows * 128 + ((w_rows * 128 + 1023) & ~1023);
  const int smem = stages * stage_bytes + 1024 + 512;
  if (smem > 227 * 1024) return;
  CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, ingest_kernel, tmA, tmW, cs, w_rows, kblocks, iters, shared_rows, rows / a_rows, stages, a_rows, g_t));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double bytes = (double)iters * (a_rows * 128 + w_rows * 128);
  long long t[5];
  CK(cudaMemcpy(t, g_t, sizeof(t), cudaMemcpyDeviceToHost));
  printf("  %-10s %4d %6d %6d %6d %6d %5s  %9.3f %9.1f %12.1f %12.1f   | %6.0f %6.0f %6.0f | %6.0f %6.0f\n", cs == 1 ? "unicast" : "multicast", cs, grid, a_rows, w_rows,
         stages, shared_rows ? "same" : "own", best, best * 1e6 / iters, bytes / best / 1e6, bytes * grid / best / 1e6,
         (double)t[0] / iters, (double)t[1] / iters, (double)t[2] / iters, (double)t[3] / iters, (double)t[4] / iters);
  fflush(stdout);
}


// ---- weight-stream mode: every CTA streams ITS OWN tiles (64 rows x 2 k-blocks = 16 KB per instruction) of a 1.2 GB
//      [N, K = 1536] bf16 matrix once (DRAM, not L2), as the GEMMs read their W operand.  Two layouts of the same data:
//      row-major [N][K] (a tile = 64 pieces of 256 B at a 3 KB stride) and k-block-major [K/64][N][64] (a tile = 2
//      contiguous 8 KB pieces).
__global__ void __launch_bounds__(64, 1)
wstream_kernel(const __grid_constant__ CUtensorMap tmW, int STAGES, int row_blocks, int kpairs) {
#if defined(__CUDA_ARCH__)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * 16384);
  uint64_t* empty = full + MAXSTAGES;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int smask = STAGES - 1, sshift = 31 - __clz(STAGES);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  int i = 0;
  if (warp == 0) {
    for (int rb = blockIdx.x; rb < row_blocks; rb += gridDim.x)
      for (int kp = 0; kp < kpairs; ++kp, ++i) {
        const int s = i & smask;
        mbar_wait(&empty[s], ((i >> sshift) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], 16384);
          tma_load_3d(smem + s * 16384, &tmW, &full[s], 0, rb * 64, kp * 2);
        }
      }
  } else {
    for (int rb = blockIdx.x; rb < row_blocks; rb += gridDim.x)
      for (int kp = 0; kp < kpairs; ++kp, ++i) {
        const int s = i & smask;
        mbar_wait(&full[s], (i >> sshift) & 1);
        if (elect_one()) mbar_arrive(&empty[s]);
      }
  }
#endif
}

static void run_wstream(PFN_encodeTiled fn, void* buf, uint64_t N, uint64_t K, int packed, int stages, int grid) {
  CUtensorMap m;
  cuuint64_t dims[3] = {64, N, K / 64};
  cuuint64_t strides_rm[2] = {K * 2, 128};
  cuuint64_t strides_pk[2] = {128, N * 128};
  cuuint32_t box[3] = {64, 64, 2};
  cuuint32_t el[3] = {1, 1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, packed ? strides_pk : strides_rm, box, el,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode (wstream) failed %d\n", (int)r); return; }
  const int smem = stages * 16384 + 1024 + 512;
  CK(cudaFuncSetAttribute(wstream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    wstream_kernel<<<grid, 64, smem>>>(m, stages, (int)(N / 64), (int)(K / 128));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  printf("  wstream %-12s stages %2d ctas %3d: %8.3f ms  %8.1f GB/s\n", packed ? "kblock-major" : "row-major", stages, grid, best,
         (double)N * K * 2 / best / 1e6);
  fflush(stdout);
}
int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 4800;
  const int K = 1536, rows = 4096;
  CK(cudaMalloc(&g_A, (size_t)rows * K * 2));
  CK(cudaMalloc(&g_W, (size_t)4096 * K * 2));
  CK(cudaMemset(g_A, 0, (size_t)rows * K * 2));
  CK(cudaMemset(g_W, 0, (size_t)4096 * K * 2));
  CK(cudaMalloc(&g_t, 64));
  CK(cudaMemset(g_t, 0, 64));
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  g_fn = reinterpret_cast<PFN_encodeTiled>(p);
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  printf("# %d k-blocks per CTA; a_rows x 64 bf16 A box (+ private w_rows x 64 W box) per k-block; GB/s RECEIVED per CTA\n", iters);
  printf("# %-10s %4s %6s %6s %6s %6s %5s  %9s %9s %12s %12s\n", "mode", "cs", "ctas", "a_rows", "w_rows", "stages", "rows", "ms",
         "ns/kblock", "GB/s per CTA", "GB/s total   | producer clk/iter: wait(empty) expect_tx tma | consumer: wait(full) arrive");
  // 1. ring depth and box size, unicast, one CTA and a full grid
  const int grids[2] = {1, 144};
  for (int g = 0; g < 2; ++g)
    for (int a_rows = 64; a_rows <= 256; a_rows *= 2)
      for (int stages = 2; stages <= 16; stages *= 2) run(1, grids[g], 0, 1, stages, a_rows, iters);
  // 2. with the GEMM's private W box next to the shared A tile
  for (int w_rows = 32; w_rows <= 128; w_rows *= 2)
    for (int stages = 4; stages <= 8; stages *= 2) run(1, 144, w_rows, 1, stages, 128, iters);
  // 3. multicast of the shared A tile inside a cluster
  for (int cs = 2; cs <= 8; cs *= 2)
    for (int stages = 4; stages <= 8; stages *= 2) {
      run(cs, cs, 0, 1, stages, 128, iters);
      run(cs, 144, 0, 1, stages, 128, iters);
      run(cs, 144, 32, 1, stages, 128, iters);
    }
  // 4. private rows instead of shared ones
  run(1, 144, 0, 0, 8, 128, iters);
  // 5. the W operand from DRAM: row-major vs k-block-major tiles
  {
    const uint64_t N = 393216, K = 1536;
    void* big;
    CK(cudaMalloc(&big, N * K * 2));
    CK(cudaMemset(big, 0, N * K * 2));
    for (int packed = 0; packed < 2; ++packed)
      for (int stages = 2; stages <= 8; stages *= 2) {
        run_wstream(g_fn, big, N, K, packed, stages, 144);
        run_wstream(g_fn, big, N, K, packed, stages, 288);
      }
    CK(cudaFree(big));
  }
  return 0;
}
