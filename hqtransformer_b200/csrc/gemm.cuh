// GEMM family of libhqgraft: C[M,N] = A[M,K] * W[N,K]^T with fused epilogues.
//
//   gemm_tc_kernel    bf16 x bf16 -> fp32 on the 5th-gen tensor cores: TMA (SWIZZLE_128B) stages A and W
//                     tiles into a ring of shared-memory slots, one thread issues tcgen05.mma with the
//                     accumulator in TMEM, four warps drain TMEM with tcgen05.ld and apply the epilogue.
//   gemm_simt_kernel  fp32 on CUDA cores, for HQ_PREC_FP32 (the reference's use_fp16=False path).
//
// Epilogues replace the separate ATen ops of the reference:
//   EPI_QKV    + bias; q -> [M, D] activation buffer, k / v -> straight into their KV-cache slot
//              (layers.py:73, 84-96: three addmm + stack + cat)
//   EPI_RESID  x += acc + bias                (layers.py:190 + :326/:327 residual adds)
//   EPI_GELU   out = gelu_erf(acc + bias)     (layers.py:312-314)
//   EPI_F32    out = acc                      (head_top / head_bot, hierarchical_ar.py:695, 715)
#pragma once

#include "common.cuh"

namespace hq {

enum { EPI_QKV = 0, EPI_RESID = 1, EPI_GELU = 2, EPI_F32 = 3 };

template <typename AT>
struct EpiParams {
  const float* bias;  // [N] fp32 or nullptr
  // EPI_QKV: column n belongs to section sec0 + n / D (0 = q, 1 = k, 2 = v)
  AT* q;              // [M, D]
  AT* kdst;           // rows of D; row index = (m / rpb) * t_stride + t0 + (m % rpb)
  AT* vdst;
  AT* vdup;           // optional second copy of the v rows as [M, D] (depth pass 0: attention output == v)
  int D, sec0, rpb, t_stride, t0;
  // EPI_RESID
  float* x;           // [M, N] fp32, read-modify-write
  // EPI_GELU
  AT* out;            // [M, N]
  // EPI_F32
  float* outf;        // [M, ldo]
  int ldo;
};

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

// 8 consecutive columns n0..n0+7 of row m (n0 % 8 == 0, all < N)
template <int EPI, typename AT>
__device__ __forceinline__ void epi_store8(const EpiParams<AT>& ep, int m, int n0, int N, float (&v)[8]) {
  if (EPI != EPI_F32 && ep.bias != nullptr) {
    const float4 b0 = *reinterpret_cast<const float4*>(ep.bias + n0);
    const float4 b1 = *reinterpret_cast<const float4*>(ep.bias + n0 + 4);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (EPI == EPI_QKV) {
    const int sec = ep.sec0 + n0 / ep.D;
    const int c = n0 % ep.D;
    if (sec == 0) {
      store8(ep.q + static_cast<size_t>(m) * ep.D + c, v);
    } else {
      const size_t row = static_cast<size_t>(m / ep.rpb) * ep.t_stride + ep.t0 + (m % ep.rpb);
      AT* dst = (sec == 1 ? ep.kdst : ep.vdst) + row * ep.D + c;
      store8(dst, v);
      if (sec == 2 && ep.vdup != nullptr) store8(ep.vdup + static_cast<size_t>(m) * ep.D + c, v);
    }
  } else if (EPI == EPI_RESID) {
    float* p = ep.x + static_cast<size_t>(m) * N + n0;
    float4 a = *reinterpret_cast<float4*>(p);
    float4 b = *reinterpret_cast<float4*>(p + 4);
    a.x += v[0]; a.y += v[1]; a.z += v[2]; a.w += v[3];
    b.x += v[4]; b.y += v[5]; b.z += v[6]; b.w += v[7];
    *reinterpret_cast<float4*>(p) = a;
    *reinterpret_cast<float4*>(p + 4) = b;
  } else if (EPI == EPI_GELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
    store8(ep.out + static_cast<size_t>(m) * N + n0, v);
  } else {
    store8(ep.outf + static_cast<size_t>(m) * ep.ldo + n0, v);
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMA GEMM.  One CTA computes a 128 x BN tile; grid = (ceil(N/BN), ceil(M/128)).
//   warp 0: TMA producer (one lane)      warp 1: tcgen05.mma issuer (one lane)
//   warps 2-5: epilogue (TMEM lane quarter = warp % 4); warp 2 also owns the TMEM allocation
// tmA: A as a [rows_pad, K] bf16 tensor, box {64, 128}; tmW: W as [N_pad, K], box {64, 64}; both SWIZZLE_128B.
// ------------------------------------------------------------------------------------------------
template <int BN>
struct TcCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int EPI, typename AT>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int M, int N, int K,
               int w_row_off, EpiParams<AT> ep) {
#if defined(__CUDA_ARCH__)
  using C = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * C::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + C::STAGES * C::B_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * C::BM;
  const int n0 = blockIdx.x * BN;
  const int num_kb = K / C::BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer ----
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], C::A_BYTES + C::B_BYTES);
        tma_load_2d(sA + s * C::A_BYTES, &tmA, &full_bar[s], kb * C::BK, m0);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_2d(sB + s * C::B_BYTES + j * (64 * 128), &tmW, &full_bar[s], kb * C::BK, w_row_off + n0 + j * 64);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ----
      constexpr uint32_t idesc = umma_idesc_bf16(C::BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t da = umma_smem_desc_sw128(smem_u32(sA + s * C::A_BYTES));
        const uint64_t db = umma_smem_desc_sw128(smem_u32(sB + s * C::B_BYTES));
#pragma unroll
        for (int k = 0; k < C::BK / 16; ++k) {
          // advance 16 elements (32 bytes) along K inside the 128-byte swizzle atom: +2 in the >>4 address field
          umma_bf16(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // ---- epilogue: TMEM -> registers -> global ----
    const int quarter = warp & 3;
    const int m = m0 + quarter * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c * 32), r);
      tmem_ld_wait();
      if (m < M) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = n0 + c * 32 + g * 8;
          if (n < N) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
            epi_store8<EPI, AT>(ep, m, n, N, v);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// fp32 CUDA-core GEMM: 64 x 128 tile, BK = 16, 256 threads, 4 x 8 outputs per thread.
// A [M, K] and W [N, K] row-major fp32 (K % 16 == 0, N % 8 == 0).
// ------------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ W, int M, int N, int K, EpiParams<float> ep) {
  constexpr int BM = 64, BN = 128, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15;   // column group: cols tx*8 .. +8
  const int ty = tid >> 4;   // row group: rows ty*4 .. +4
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int a_row = tid >> 2, a_k = (tid & 3) * 4;  // 64 rows x 4 float4
  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + a_row < M) v = *reinterpret_cast<const float4*>(A + static_cast<size_t>(m0 + a_row) * K + k0 + a_k);
      As[a_k + 0][a_row] = v.x; As[a_k + 1][a_row] = v.y; As[a_k + 2][a_row] = v.z; As[a_k + 3][a_row] = v.w;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int w_row = (tid >> 2) + h * 64;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + w_row < N) v = *reinterpret_cast<const float4*>(W + static_cast<size_t>(n0 + w_row) * K + k0 + a_k);
      Ws[a_k + 0][w_row] = v.x; Ws[a_k + 1][w_row] = v.y; Ws[a_k + 2][w_row] = v.z; Ws[a_k + 3][w_row] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ws[k][tx * 8 + 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 8;
  if (n < N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m < M) epi_store8<EPI, float>(ep, m, n, N, acc[i]);
    }
  }
}

}  // namespace hq
