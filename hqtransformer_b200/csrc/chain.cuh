// Persistent op-chain kernel of libhqgraft (bf16 engine).
//
// A decode position is ~145 dependent kernels; every tcgen05 GEMM launch pays ~3 us of fixed cost that is not the
// kernel boundary itself (profiles/r2_grid_barrier_bench.txt: a grid barrier costs 1.2 us, a graph + PDL boundary
// 1.8 us) but what each launch repeats inside: barrier init + TMEM allocation + cluster sync, a cold first ring of
// weight tiles, teardown (profiles/r1_gemm_trace.txt).  chain_kernel runs a whole SEQUENCE of dependent ops - the
// depth transformer of a position (hierarchical_ar.py:667-789: 4-6 blocks x {LayerNorm, q/k/v GEMM, 4x5 attention,
// proj GEMM, LayerNorm, fc1+GELU GEMM, fc2 GEMM}, final LayerNorm, head GEMM) or the part of a spatial block between
// two cache attentions - in ONE launch of one CTA pair per SM pair:
//   * TMEM (two 256-column accumulators), the shared-memory ring and its mbarrier phases live for the whole chain;
//   * ops are separated by a hand-rolled grid barrier (red.release / ld.acquire on a global counter, split into
//     arrive and wait); the TMA producer of the NEXT GEMM requests its first ring of weight tiles BEFORE it waits -
//     weights never depend on the previous op - so the weight stream runs across op boundaries;
//   * LayerNorm (folding the split-K partial sums of the preceding residual GEMM), the depth attention and the depth
//     embedding run on the eight epilogue warps between GEMMs; the five tokens of a depth stack never leave L2.
// Arithmetic is bit-identical to the one-kernel-per-op path (same K order, same split-K slices, same reductions):
// tests compare the two paths exactly.
#pragma once

#include "gemm.cuh"
#include "kernels.cuh"

namespace hq {

enum { CH_OP_GEMM = 0, CH_OP_LN = 1, CH_OP_ATTN4 = 2, CH_OP_EMBED_DEPTH = 3 };
enum { CH_F_T0_RT = 1, CH_F_OUT_F32 = 2 };

struct ChainOp {
  int kind, flags;
  // ---- CH_OP_GEMM: C[M, N] = A[M, K] W[N, K]^T on CTA pairs, 256 x bn tiles, `splits` K slices
  int map_a, map_w;                 // indices into the ctx's tensor-map table (device memory)
  int M, N, K, bn, splits, epi, w_row_off, pad0;
  EpiParams<bf16> ep;
  // ---- CH_OP_LN: layernorm_kernel's arguments
  float* x;
  const float* gamma;
  const float* beta;
  const float* add;
  void* out;
  int rows, in_mul, in_off, n_fold;
  const float* fold;
  size_t fold_stride;
  const float* fold_bias;
  // ---- CH_OP_ATTN4 (attention_depth4_kernel's arguments) / CH_OP_EMBED_DEPTH (embed_depth_kernel's)
  const bf16* q;
  const bf16* kc;
  const bf16* vc;
  bf16* att;
  int B, n_heads, D, t_stride, n_keys, pad1;
  float* y;
  const float* E;
  const float* P;
  const int64_t* codes_top;
};

struct ChainRt {                    // per-launch values (everything else of a chain is position independent)
  int t0;                           // spatial cache slot of this position's token (GEMM ops flagged CH_F_T0_RT)
  int pos, S;                       // top position / row stride of the code arrays (CH_OP_EMBED_DEPTH)
  int trace_base;                   // >= 0: op i records its span in g_hq_trace[trace_base + i]
  int no_l2_prefetch;               // experiments: 1 = do not request the op's later W tiles into L2 ahead of the barrier
  int phase_op;                     // >= 0 (and g_hq_phase set): per-CTA %globaltimer stamps of that op (hq_debug_chain_phases)
  unsigned long long* bar;          // grid barrier counter (zeroed at the start of every run)
  unsigned long long bar_base;      // its value when this launch starts: arrivals of all earlier chain launches of the run
};

constexpr int CH_STAGES = 6;
constexpr int CH_A_BYTES = 128 * 64 * 2;              // this CTA's 128 rows of a 64-wide k-block of A
constexpr int CH_STAGE_BYTES = 2 * CH_A_BYTES;        // + up to 128 rows (bn / 2) of W
constexpr int CH_EPI_WARPS = 8;
constexpr int CH_THREADS = (2 + CH_EPI_WARPS) * 32;
constexpr int CH_SLAB_BYTES = CH_EPI_WARPS * 4096;
constexpr int CH_ACC_COLS = 256;
constexpr int CH_OPBUF_BYTES = 2 * ((static_cast<int>(sizeof(ChainOp)) + 15) / 16 * 16);   // two staged op descriptors
constexpr int CH_SMEM_BYTES = CH_STAGES * CH_STAGE_BYTES + CH_SLAB_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*bias*/ +
                              CH_OPBUF_BYTES;
static_assert(CH_SMEM_BYTES <= 227 * 1024, "chain kernel shared memory");
static_assert(sizeof(ChainOp) % 16 == 0, "ChainOp is staged with 16-byte copies");

// per-CTA phase stamps of one op (hq_debug_chain_phases); separate from g_hq_phase, which the attention kernel uses
__device__ unsigned long long* g_hq_chain_phase = nullptr;

#if defined(__CUDA_ARCH__)
// ---- grid barrier: a monotonically increasing arrival counter ----
__device__ __forceinline__ void grid_arrive(unsigned long long* bar) {
  // the release itself orders (cumulatively) everything the CTA's warps wrote before the named barrier that precedes it
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(1ull) : "memory");
}
// A wait that needs more than ~16M polls can only be a protocol bug (or a grid that is not co-resident): trap.
// The polls are RELAXED loads and one acquire fence follows the successful one: an acquire load is LDG.STRONG + CCTL.IVALL
// (invalidate the whole L1), and a thread spinning on it kept wiping the L1 under the warps of the same SM that were
// still working (their spilled registers and descriptor reads turned into L2 round trips: ops ran 2-3x slower).
__device__ __forceinline__ void grid_wait(const unsigned long long* bar, unsigned long long target) {
  uint32_t polls = 0;
  for (;;) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
    if (v >= target) break;
    if (++polls > (1u << 24)) __trap();
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// generic-proxy global writes -> async-proxy (TMA) reads of other CTAs, and back
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void chain_phase_mark(int p) {
  if (g_hq_chain_phase != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_hq_chain_phase[blockIdx.x * 8 + p] = t;
  }
}

// the next op's descriptor lines into L1 while the current op runs (each first touch would be a serialized L2 miss)
__device__ __forceinline__ void prefetch_op(const ChainOp* op) {
  const char* p = reinterpret_cast<const char*>(op);
#pragma unroll
  for (int o = 0; o < static_cast<int>(sizeof(ChainOp)); o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + o));
}
// one W box (the same tile a later tma_load_2d fetches) into L2
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
               "r"(c1)
               : "memory");
}

__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Epilogue of one 128 x bn accumulator (this CTA's rows of a pair tile): epilogue_tile with a run-time tile width,
// eight warps (two per TMEM lane quarter, alternating 32-column chunks).
template <int EPI>
__device__ __forceinline__ void chain_epilogue_tile(uint32_t tmem_base, uint8_t* slab_base, float* sbias, int warp, int lane,
                                                    int ew, int bn, int m0, int n0, int M, int N, const EpiParams<bf16>& ep,
                                                    uint64_t* tmem_full_bar, uint32_t full_parity) {
  const int quarter = warp & 3;
  const int half = ew >> 2;
  const int etid = ew * 32 + lane;
  const bool has_bias = (EPI != EPI_F32) && ep.bias != nullptr;
  if (has_bias) {
    named_bar(1, 256);                                         // previous tile's bias fully consumed
    for (int i = etid; i < bn; i += 256) sbias[i] = (n0 + i < N) ? ep.bias[n0 + i] : 0.f;
  }
  const int piece = lane & 7;
  size_t row_a[8], row_b[8];
  bool row_ok[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int m = m0 + quarter * 32 + it * 4 + (lane >> 3);
    row_ok[it] = m < M;
    if (EPI == EPI_QKV) {
      row_a[it] = static_cast<size_t>(m) * ep.D;
      row_b[it] = (static_cast<size_t>(m / ep.rpb) * ep.t_stride + ep.t0 + (m % ep.rpb)) * ep.D;
    } else if (EPI == EPI_F32) {
      row_a[it] = static_cast<size_t>(m) * ep.ldo;
      row_b[it] = 0;
    } else {
      row_a[it] = static_cast<size_t>(m) * N;
      row_b[it] = 0;
    }
  }
  named_bar(1, 256);                                           // sbias visible to all epilogue warps
  mbar_wait(tmem_full_bar, full_parity);
  tc_fence_after();

  uint8_t* slab = slab_base + ew * 4096;
  const uint32_t slab_u32 = smem_u32(slab);
#pragma unroll 1
  for (int c = half; c < bn / 32; c += 2) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c * 32), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t addr = slab_u32 + lane * 128 + ((j ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                   "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                   : "memory");
    }
    __syncwarp();
    const int nb = n0 + c * 32;
    const int n = nb + piece * 4;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_bias) b4 = *reinterpret_cast<const float4*>(sbias + c * 32 + piece * 4);
    int sec = 0, col = n;
    if (EPI == EPI_QKV) {
      sec = ep.sec0 + nb / ep.D;
      col = nb % ep.D + piece * 4;
    }
    if (n < N) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + (lane >> 3);
        float4 v;
        const uint32_t addr = slab_u32 + row * 128 + ((piece ^ (row & 7)) << 4);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
        v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
        if (!row_ok[it]) continue;
        if (EPI == EPI_F32) {
          *reinterpret_cast<float4*>(ep.outf + row_a[it] + col) = v;
        } else if (EPI == EPI_RESID) {
          float4* p = reinterpret_cast<float4*>(ep.x + row_a[it] + col);
          float4 a = *p;
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
          *p = a;
        } else {
          bf16* dst;
          bf16* dup = nullptr;
          if (EPI == EPI_GELU) {
            v.x = gelu_bf16out(v.x); v.y = gelu_bf16out(v.y); v.z = gelu_bf16out(v.z); v.w = gelu_bf16out(v.w);   // bf16 output, as the per-op path
            dst = ep.out + row_a[it] + col;
          } else if (sec == 0) {
            dst = ep.q + row_a[it] + col;
          } else {
            dst = (sec == 1 ? ep.kdst : ep.vdst) + row_b[it] + col;
            if (sec == 2 && ep.vdup != nullptr) dup = ep.vdup + row_a[it] + col;
          }
          uint2 u;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
          u.x = *reinterpret_cast<uint32_t*>(&h0);
          u.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(dst) = u;
          if (dup != nullptr) *reinterpret_cast<uint2*>(dup) = u;
        }
      }
    }
    __syncwarp();
  }
}

// LayerNorm rows on a team of 128 threads (four epilogue warps): layernorm_kernel's register path, same arithmetic.
__device__ __forceinline__ float team_sum_128(float v, float* red /*[4]*/, int wt, int lane, int bar_id) {
  v = warp_sum(v);
  named_bar(bar_id, 128);
  if (lane == 0) red[wt] = v;
  named_bar(bar_id, 128);
  return (red[0] + red[1]) + (red[2] + red[3]);
}

// The per-column parameters (gamma, beta, addend, bias of the folded GEMM) do not depend on the previous op: they are
// requested BEFORE the grid barrier (LnCols), the rows after it.
struct LnCols {
  float4 g[LN_MAXV], bt[LN_MAXV], ad[LN_MAXV], fb[LN_MAXV];
};
__device__ __forceinline__ void chain_ln_cols(const ChainOp* __restrict__ op, int tid, LnCols& c) {
  const int D = op->D;
  const float* __restrict__ gamma = op->gamma;
  const float* __restrict__ beta = op->beta;
  const float* __restrict__ add = op->add;
  const float* __restrict__ fold_bias = op->fold != nullptr ? op->fold_bias : nullptr;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < LN_MAXV; ++j) {
    const int i = (j * LN_THREADS + tid) * 4;
    const int ic = i < D ? i : 0;
    c.g[j] = *reinterpret_cast<const float4*>(gamma + ic);
    c.bt[j] = *reinterpret_cast<const float4*>(beta + ic);
    c.ad[j] = add != nullptr ? *reinterpret_cast<const float4*>(add + ic) : z4;
    c.fb[j] = fold_bias != nullptr ? *reinterpret_cast<const float4*>(fold_bias + ic) : z4;
  }
}

template <typename OutT>
__device__ __forceinline__ void chain_layernorm(const ChainOp* __restrict__ op, const LnCols& cols, int team_global, int n_teams,
                                                int tid, int wt, int lane, float* red, int bar_id) {
  const int D = op->D, rows = op->rows, in_mul = op->in_mul, in_off = op->in_off, n_fold = op->n_fold;
  const float* __restrict__ fold = op->fold;
  const size_t fold_stride = op->fold_stride;
  const float4 (&g)[LN_MAXV] = cols.g;
  const float4 (&bt)[LN_MAXV] = cols.bt;
  const float4 (&ad)[LN_MAXV] = cols.ad;
  const float4 (&fb)[LN_MAXV] = cols.fb;
  for (int r = team_global; r < rows; r += n_teams) {
    float* xr = op->x + (static_cast<size_t>(r) * in_mul + in_off) * D;
    OutT* o = static_cast<OutT*>(op->out) + static_cast<size_t>(r) * D;
    float4 v[LN_MAXV], f[LN_MAXFOLD][LN_MAXV];
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) {
      const int i = (j * LN_THREADS + tid) * 4;
      const int ic = i < D ? i : 0;
      v[j] = *reinterpret_cast<const float4*>(xr + ic);
#pragma unroll
      for (int sidx = 0; sidx < LN_MAXFOLD; ++sidx) f[sidx][j] = z4;
      if (fold != nullptr) {
        const float* fr = fold + (static_cast<size_t>(r) * in_mul + in_off) * D + ic;
#pragma unroll
        for (int sidx = 0; sidx < LN_MAXFOLD; ++sidx)
          if (sidx < n_fold) f[sidx][j] = *reinterpret_cast<const float4*>(fr + sidx * fold_stride);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) {
      const int i = (j * LN_THREADS + tid) * 4;
      if (fold != nullptr) {
        float4 a = f[0][j];
#pragma unroll
        for (int sidx = 1; sidx < LN_MAXFOLD; ++sidx) {
          a.x += f[sidx][j].x; a.y += f[sidx][j].y; a.z += f[sidx][j].z; a.w += f[sidx][j].w;
        }
        v[j].x += fb[j].x + a.x;
        v[j].y += fb[j].y + a.y;
        v[j].z += fb[j].z + a.z;
        v[j].w += fb[j].w + a.w;
        if (i < D) *reinterpret_cast<float4*>(xr + i) = v[j];
      }
      if (i < D) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = team_sum_128(s, red, wt, lane, bar_id) / static_cast<float>(D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) {
      if ((j * LN_THREADS + tid) * 4 < D) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = 1.0f / sqrtf(team_sum_128(q, red, wt, lane, bar_id) / static_cast<float>(D) + 1e-5f);
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) {
      const int i = (j * LN_THREADS + tid) * 4;
      if (i < D) {
        ln_store4<OutT>(o, i, (v[j].x - mean) * rstd * g[j].x + bt[j].x + ad[j].x,
                        (v[j].y - mean) * rstd * g[j].y + bt[j].y + ad[j].y,
                        (v[j].z - mean) * rstd * g[j].z + bt[j].z + ad[j].z,
                        (v[j].w - mean) * rstd * g[j].w + bt[j].w + ad[j].w);
      }
    }
  }
}

// depth pass 1 attention: attention_depth4_kernel's arithmetic, one warp per (image, head) item.  Only eight warps per SM
// run it here, so each warp keeps TWO items in flight (all 22 loads of a lane are issued before the first use).
struct Attn4Item {
  Raw8<bf16> qr, kr[ATT_DEPTH_KEYS], vr[ATT_DEPTH_KEYS];
  size_t qoff;
};
__device__ __forceinline__ void attn4_load(const ChainOp* __restrict__ op, int item, int lane, Attn4Item& a) {
  const int n_heads = op->n_heads, D = op->D, n_keys = op->n_keys;
  const int b = item / n_heads, h = item % n_heads;
  const int g = lane >> 3, c = lane & 7;
  a.qoff = (static_cast<size_t>(b) * 4 + g) * D + h * 64 + c * 8;
  const size_t kbase = static_cast<size_t>(b) * op->t_stride * D + h * 64 + c * 8;
  a.qr.load(op->q + a.qoff);
#pragma unroll
  for (int t = 0; t < ATT_DEPTH_KEYS; ++t) {
    const int tc = t < n_keys ? t : 0;
    a.kr[t].load(op->kc + kbase + static_cast<size_t>(tc) * D);
    a.vr[t].load(op->vc + kbase + static_cast<size_t>(tc) * D);
  }
}
__device__ __forceinline__ void attn4_finish(const ChainOp* __restrict__ op, const Attn4Item& a) {
  const int n_keys = op->n_keys;
  float qv[8];
  a.qr.get(qv);
  float s[ATT_DEPTH_KEYS];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < ATT_DEPTH_KEYS; ++t) {
    float kv[8];
    a.kr[t].get(kv);
    float d = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) d = fmaf(qv[e], kv[e] * 0.125f, d);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    s[t] = t < n_keys ? d : -INFINITY;
    mx = fmaxf(mx, s[t]);
  }
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < ATT_DEPTH_KEYS; ++t) {
    s[t] = t < n_keys ? expf(s[t] - mx) : 0.f;
    sum += s[t];
  }
  const float inv = 1.0f / sum;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
  for (int t = 0; t < ATT_DEPTH_KEYS; ++t) {
    float vv[8];
    a.vr[t].get(vv);
    const float p = s[t] * inv;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, vv[e], acc[e]);
  }
  store8(op->att + a.qoff, acc);
}
__device__ __forceinline__ void chain_attn4(const ChainOp* __restrict__ op, int warp_global, int n_warps, int lane) {
  const int n_items = op->B * op->n_heads;
  for (int item = warp_global; item < n_items; item += 2 * n_warps) {
    Attn4Item a0, a1;
    const bool two = item + n_warps < n_items;           // warp-uniform
    attn4_load(op, item, lane, a0);
    if (two) attn4_load(op, item + n_warps, lane, a1);
    attn4_finish(op, a0);
    if (two) attn4_finish(op, a1);
  }
}

// depth pass-1 inputs (embed_depth_kernel): y[b*4 + j] = E_top_depth[c_top[b]] + P_depth[j]; one 128-thread team per image
__device__ __forceinline__ void chain_embed_depth(const ChainOp* __restrict__ op, const ChainRt& rt, int team_global,
                                                  int n_teams, int tid) {
  const int D = op->D, n4 = D / 4;
  const float4* p = reinterpret_cast<const float4*>(op->P);
  for (int b = team_global; b < op->B; b += n_teams) {
    const int64_t ct = op->codes_top[static_cast<size_t>(b) * rt.S + rt.pos];
    const float4* e = reinterpret_cast<const float4*>(op->E + static_cast<size_t>(ct) * D);
    for (int i = tid; i < n4; i += 128) {
      const float4 a = e[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 c = p[j * n4 + i];
        reinterpret_cast<float4*>(op->y + (static_cast<size_t>(b) * 4 + j) * D)[i] =
            make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
      }
    }
  }
}
#endif  // __CUDA_ARCH__

// grid = 2 * P CTAs (P co-resident CTA pairs, one CTA per SM), cluster (2,1,1), CH_THREADS threads:
//   warp 0: TMA producer (one lane, both CTAs)   warp 1: tcgen05.mma issuer (one lane, leader CTA)
//   warps 2-9: GEMM epilogues and every non-GEMM op; thread etid == 0 arrives on the grid barrier after each op.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CH_THREADS, 1)
chain_kernel(int trace_id, const ChainOp* __restrict__ ops, int n_ops, const CUtensorMap* __restrict__ maps, ChainRt rt) {
#if defined(__CUDA_ARCH__)
  TraceScope trace_scope(trace_id);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* slab = smem + CH_STAGES * CH_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slab + CH_SLAB_BYTES);
  uint64_t* empty_bar = full_bar + CH_STAGES;
  uint64_t* tmem_full_bar = empty_bar + CH_STAGES;       // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2], used in the leader CTA only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* ln_red = reinterpret_cast<float*>(tmem_slot + 2);   // [2 teams][4]
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);
  // op descriptors of the epilogue warps, staged one op ahead: an acquire poll invalidates L1, so every descriptor field
  // read after a grid barrier would otherwise be a serialized L2 round trip on the critical path
  ChainOp* sop = reinterpret_cast<ChainOp*>(reinterpret_cast<uint8_t*>(sbias) + 1024);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const unsigned long long nctas = gridDim.x;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < CH_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 2 * CH_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, 2 * CH_ACC_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: own A rows + own half of the W tile of every k-block, credited to the leader's barrier.
      //      First ring of an op: W tiles are requested BEFORE the grid barrier / griddepcontrol.wait, A after it. ----
      uint32_t g = 0;
      bool pdl_done = false;
      for (int i = 0; i < n_ops; ++i) {
        const ChainOp* __restrict__ op = ops + i;
        if (op->kind != CH_OP_GEMM) continue;
        const int bn = op->bn, splits = op->splits;
        const int nt = op->N / bn, mt = (op->M + 255) / 256;
        const int total_tiles = nt * mt * splits;
        const int num_kb = (op->K / 64) / splits;
        const CUtensorMap* mA = maps + op->map_a;
        const CUtensorMap* mW = maps + op->map_w;
        const uint32_t stage_tx = 2u * static_cast<uint32_t>(CH_A_BYTES + bn * 64);
        const int w_row_off = op->w_row_off;
        bool synced = false;
        for (int tile = pair; tile < total_tiles; tile += npairs) {
          const int z = tile / (nt * mt), rem = tile % (nt * mt);
          const int m0 = (rem / nt) * 256 + static_cast<int>(rank) * 128;
          const int wrow = w_row_off + (rem % nt) * bn + static_cast<int>(rank) * (bn / 2);
          const int kb0 = z * num_kb;
          int kb = 0;
          if (!synced) {
            synced = true;
            tma_prefetch_desc(mA);
            tma_prefetch_desc(mW);
            const int pre = num_kb < CH_STAGES ? num_kb : CH_STAGES;
            // every W box this CTA will load in this op beyond the first ring -> L2 now, i.e. during the previous op's
            // tail and the grid barrier: when the stream gets there it runs at L2 latency, not HBM latency
            for (int t2 = tile; t2 < total_tiles && !rt.no_l2_prefetch; t2 += npairs) {
              const int z2 = t2 / (nt * mt), rem2 = t2 % (nt * mt);
              const int wrow2 = w_row_off + (rem2 % nt) * bn + static_cast<int>(rank) * (bn / 2);
              for (int k2 = (t2 == tile ? pre : 0); k2 < num_kb; ++k2) tma_prefetch_l2_2d(mW, (z2 * num_kb + k2) * 64, wrow2);
            }
            for (; kb < pre; ++kb) {
              const uint32_t gg = g + static_cast<uint32_t>(kb);
              const int s = static_cast<int>(gg % CH_STAGES);
              mbar_wait(&empty_bar[s], ((gg / CH_STAGES) & 1) ^ 1);
              if (leader) mbar_arrive_expect_tx(&full_bar[s], stage_tx);
              tma_load_2d_2sm(smem + s * CH_STAGE_BYTES + CH_A_BYTES, mW, &full_bar[s], (kb0 + kb) * 64, wrow);
            }
            if (!pdl_done) {
              pdl_wait();
              pdl_done = true;
            }
            if (i > 0) grid_wait(rt.bar, rt.bar_base + static_cast<unsigned long long>(i) * nctas);
            fence_proxy_async_global();
            for (int k2 = 0; k2 < pre; ++k2) {
              const int s = static_cast<int>((g + static_cast<uint32_t>(k2)) % CH_STAGES);
              tma_load_2d_2sm(smem + s * CH_STAGE_BYTES, mA, &full_bar[s], (kb0 + k2) * 64, m0);
            }
            g += static_cast<uint32_t>(pre);
          }
          for (; kb < num_kb; ++kb, ++g) {
            const int s = static_cast<int>(g % CH_STAGES);
            mbar_wait(&empty_bar[s], ((g / CH_STAGES) & 1) ^ 1);
            if (leader) mbar_arrive_expect_tx(&full_bar[s], stage_tx);
            tma_load_2d_2sm(smem + s * CH_STAGE_BYTES, mA, &full_bar[s], (kb0 + kb) * 64, m0);
            tma_load_2d_2sm(smem + s * CH_STAGE_BYTES + CH_A_BYTES, mW, &full_bar[s], (kb0 + kb) * 64, wrow);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // ---- MMA issuer (leader CTA) ----
      uint32_t g = 0, it = 0;
      for (int i = 0; i < n_ops; ++i) {
        const ChainOp* __restrict__ op = ops + i;
        if (op->kind != CH_OP_GEMM) continue;
        const int bn = op->bn, splits = op->splits;
        const int total_tiles = (op->N / bn) * ((op->M + 255) / 256) * splits;
        const int num_kb = (op->K / 64) / splits;
        const uint32_t idesc = umma_idesc_bf16(256, bn);
        for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
          const uint32_t buf = it & 1;
          mbar_wait(&tmem_empty_bar[buf], ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t acc = tmem_base + buf * CH_ACC_COLS;
          for (int kb = 0; kb < num_kb; ++kb, ++g) {
            const int s = static_cast<int>(g % CH_STAGES);
            mbar_wait(&full_bar[s], (g / CH_STAGES) & 1);
            tc_fence_after();
            const uint64_t da = umma_smem_desc_sw128(smem_u32(smem + s * CH_STAGE_BYTES));
            const uint64_t db = umma_smem_desc_sw128(smem_u32(smem + s * CH_STAGE_BYTES + CH_A_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_2sm(acc, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                            (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[s], 0x3);
          }
          umma_commit_2sm(&tmem_full_bar[buf], 0x3);
        }
      }
    }
  } else {
    // ---- epilogue warps: GEMM epilogues, LayerNorm, depth attention, depth embedding; grid barrier after every op ----
    const int ew = (warp & 3) + ((warp - 2) >> 2) * 4;          // 0..7; TMEM lane quarter = ew & 3 = warp & 3
    const int etid = ew * 32 + lane;
    const int team = ew >> 2, wt = ew & 3, ttid = wt * 32 + lane;
    const int n_teams = 2 * static_cast<int>(gridDim.x), team_global = static_cast<int>(blockIdx.x) * 2 + team;
    const int n_warps = CH_EPI_WARPS * static_cast<int>(gridDim.x), warp_global = static_cast<int>(blockIdx.x) * CH_EPI_WARPS + ew;
    uint32_t it = 0;
    constexpr int OP_VEC = static_cast<int>(sizeof(ChainOp) / 16);
    if (etid < OP_VEC) reinterpret_cast<uint4*>(&sop[0])[etid] = reinterpret_cast<const uint4*>(ops)[etid];
    for (int i = 0; i < n_ops; ++i) {
      // descriptor of op i + 1 -> the other buffer (its last readers finished op i - 1); visible after this op's barriers
      if (i + 1 < n_ops && etid < OP_VEC)
        reinterpret_cast<uint4*>(&sop[(i + 1) & 1])[etid] = reinterpret_cast<const uint4*>(ops + i + 1)[etid];
      if (i == 0) named_bar(2, 256);
      const ChainOp* op = &sop[i & 1];
      if (rt.trace_base >= 0 && etid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(&g_hq_trace[2 * (rt.trace_base + i)], t);
      }
      const bool ph = etid == 0 && i == rt.phase_op;
      if (ph) chain_phase_mark(0);                                     // op begins (previous op's arrival done)
      const int kind = op->kind;
      LnCols cols;
      if (kind == CH_OP_LN) chain_ln_cols(op, ttid, cols);       // nothing here depends on the previous op
      if (i == 0) {
        pdl_wait();
      } else {
        // ONE thread per CTA polls the barrier word (1184 pollers on one L2 line slowed every arrival down);
        // the other 255 wait for it at a hardware barrier
        if (etid == 0) grid_wait(rt.bar, rt.bar_base + static_cast<unsigned long long>(i) * nctas);
        if (ph) chain_phase_mark(1);                                   // barrier seen
        named_bar(2, 256);
      }
      if (ph) chain_phase_mark(2);
      if (kind == CH_OP_GEMM) {
        const int bn = op->bn, splits = op->splits, M = op->M, N = op->N;
        const int nt = N / bn, mt = (M + 255) / 256;
        const int total_tiles = nt * mt * splits;
        const int epi = op->epi;
        EpiParams<bf16> ep = op->ep;
        if (op->flags & CH_F_T0_RT) ep.t0 = rt.t0;
        float* const outf0 = ep.outf;
        for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
          const uint32_t buf = it & 1;
          const int z = tile / (nt * mt), rem = tile % (nt * mt);
          const int m0 = (rem / nt) * 256 + static_cast<int>(rank) * 128;
          const int n0 = (rem % nt) * bn;
          const uint32_t acc = tmem_base + buf * CH_ACC_COLS;
          const uint32_t par = (it >> 1) & 1;
          if (epi == EPI_F32) {
            ep.outf = outf0 + static_cast<size_t>(z) * ep.split_stride;
            chain_epilogue_tile<EPI_F32>(acc, slab, sbias, warp, lane, ew, bn, m0, n0, M, N, ep, &tmem_full_bar[buf], par);
          } else if (epi == EPI_QKV) {
            chain_epilogue_tile<EPI_QKV>(acc, slab, sbias, warp, lane, ew, bn, m0, n0, M, N, ep, &tmem_full_bar[buf], par);
          } else if (epi == EPI_GELU) {
            chain_epilogue_tile<EPI_GELU>(acc, slab, sbias, warp, lane, ew, bn, m0, n0, M, N, ep, &tmem_full_bar[buf], par);
          } else {
            chain_epilogue_tile<EPI_RESID>(acc, slab, sbias, warp, lane, ew, bn, m0, n0, M, N, ep, &tmem_full_bar[buf], par);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&tmem_empty_bar[buf], 0);
        }
      } else if (kind == CH_OP_LN) {
        if (op->flags & CH_F_OUT_F32) chain_layernorm<float>(op, cols, team_global, n_teams, ttid, wt, lane, ln_red + team * 4, 4 + team);
        else chain_layernorm<bf16>(op, cols, team_global, n_teams, ttid, wt, lane, ln_red + team * 4, 4 + team);
      } else if (kind == CH_OP_ATTN4) {
        chain_attn4(op, warp_global, n_warps, lane);
      } else {
        chain_embed_depth(op, rt, team_global, n_teams, ttid);
      }
      // publish this CTA's results of op i: every thread orders its own writes for TMA readers, then one arrival
      if (ph) chain_phase_mark(3);                                     // this warp's work done
      fence_proxy_async_global();
      if (ph) chain_phase_mark(4);
      named_bar(3, 256);
      if (ph) chain_phase_mark(5);                                     // every epilogue warp done
      if (etid == 0) {
        grid_arrive(rt.bar);
        if (ph) chain_phase_mark(6);
        if (rt.trace_base >= 0) {
          unsigned long long t;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
          atomicMax(&g_hq_trace[2 * (rt.trace_base + i) + 1], t);
        }
      }
    }
  }
  __syncwarp();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * CH_ACC_COLS);
  }
#endif
}

}  // namespace hq
