// Microbenchmark (GPU box): what does a grid-wide barrier cost on B200 next to the kernel boundary it would replace?
// Decides whether the depth transformer (and the whole position) should be ONE persistent kernel: a decode position is
// ~145 dependent phases, so the per-phase synchronisation cost is the design constant.
//   A  cooperative groups grid.sync()                                   (cudaLaunchCooperativeKernel)
//   B  hand-rolled: __syncthreads; thread 0: fence + atomicAdd (monotonic counter); spin ld.acquire; __syncthreads
//   C  hand-rolled: red.release (no return value) + spin ld.acquire on the same counter
//   D  as C, with a realistic phase between barriers: every CTA writes 8 KB and, after the barrier, reads the 8 KB
//      another CTA wrote (checks visibility; includes the store drain + L2 round trip a real phase pays)
//   E  kernel boundary: chain of dependent kernels of the same 8 KB write / read phase - plain stream launches,
//      programmatic dependent launch (griddepcontrol), and the same chains replayed from a CUDA graph
// grid = one CTA per SM (148) x 256 threads, and 2 CTAs per SM (296) for A-C.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/grid_barrier_bench.bin scripts/grid_barrier_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded spin: a protocol bug traps instead of hanging the box
__device__ __forceinline__ void spin_until(const unsigned* p, unsigned target) {
  unsigned polls = 0;
  while (static_cast<int>(ld_acquire(p) - target) < 0) {
    if (++polls > (1u << 26)) __trap();
  }
}

__global__ void __launch_bounds__(256) k_coop(int iters, unsigned* sink) {
  cg::grid_group g = cg::this_grid();
  for (int i = 0; i < iters; ++i) g.sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}

__global__ void __launch_bounds__(256) k_atomic(int iters, unsigned* counter) {
  for (int i = 0; i < iters; ++i) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
      spin_until(counter, static_cast<unsigned>(i + 1) * gridDim.x);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_red(int iters, unsigned* counter) {
  for (int i = 0; i < iters; ++i) {
    __syncthreads();
    if (threadIdx.x == 0) {
      red_release(counter, 1u);
      spin_until(counter, static_cast<unsigned>(i + 1) * gridDim.x);
    }
    __syncthreads();
  }
}

// phase: CTA b writes buf[parity][b][0..2047] (8 KB), barrier, reads what CTA (b + 37) % grid wrote and checks it
__device__ __forceinline__ void phase_write(float* buf, int it, int b, int nb) {
  float4* dst = reinterpret_cast<float4*>(buf + (static_cast<size_t>(it & 1) * nb + b) * 2048);
  const float v = static_cast<float>(it * 1000 + b);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) dst[i] = make_float4(v, v, v, v);
}
__device__ __forceinline__ int phase_check(const float* buf, int it, int b, int nb) {
  const int src = (b + 37) % nb;
  const float4* p = reinterpret_cast<const float4*>(buf + (static_cast<size_t>(it & 1) * nb + src) * 2048);
  const float want = static_cast<float>(it * 1000 + src);
  int bad = 0;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
    bad += (v.x != want) + (v.w != want);
  }
  return bad;
}

__global__ void __launch_bounds__(256) k_red_phase(int iters, unsigned* counter, float* buf, int* errors) {
  int bad = 0;
  for (int i = 0; i < iters; ++i) {
    phase_write(buf, i, blockIdx.x, gridDim.x);
    __syncthreads();
    if (threadIdx.x == 0) {
      red_release(counter, 1u);
      spin_until(counter, static_cast<unsigned>(i + 1) * gridDim.x);
    }
    __syncthreads();
    bad += phase_check(buf, i, blockIdx.x, gridDim.x);
  }
  if (bad) atomicAdd(errors, bad);
}

// the same phase as one kernel per iteration; every kernel checks the previous one's writes
__global__ void __launch_bounds__(256) k_phase_kernel(int it, float* buf, int* errors, int pdl) {
  if (pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  int bad = 0;
  if (it > 0) bad = phase_check(buf, it - 1, blockIdx.x, gridDim.x);
  phase_write(buf, it, blockIdx.x, gridDim.x);
  if (bad) atomicAdd(errors, bad);
}

static float time_ms(cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1) {
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

static void launch_chain(cudaStream_t st, int n, int grid, float* buf, int* errors, bool pdl) {
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k_phase_kernel, i, buf, errors, pdl ? 1 : 0));
  }
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  unsigned* counter;
  float* buf;
  int* errors;
  CK(cudaMalloc(&counter, 256));
  CK(cudaMalloc(&buf, static_cast<size_t>(2) * 2 * sms * 2048 * 4));
  CK(cudaMalloc(&errors, 4));
  CK(cudaMemset(errors, 0, 4));
  printf("grid barrier vs kernel boundary, %d SMs, %d iterations each, 256 threads per CTA\n", sms, iters);

  for (int per_sm = 1; per_sm <= 2; ++per_sm) {
    const int grid = sms * per_sm;
    // A: cooperative groups
    {
      int it = iters;
      unsigned* sink = counter + 32;
      void* args[] = {&it, &sink};
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0, st));
        CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_coop), dim3(grid), dim3(256), args, 0, st));
        CK(cudaEventRecord(e1, st));
        const float ms = time_ms(st, e0, e1);
        if (rep) printf("A cg::grid.sync          grid %3d : %.3f us per barrier\n", grid, ms * 1e3 / iters);
      }
    }
    // B / C
    for (int variant = 0; variant < 2; ++variant) {
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemsetAsync(counter, 0, 4, st));
        CK(cudaEventRecord(e0, st));
        int it = iters;
        void* args[] = {&it, &counter};
        // cooperative launch only to guarantee co-residency (a spin barrier deadlocks otherwise)
        CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(variant == 0 ? k_atomic : k_red), dim3(grid), dim3(256), args, 0, st));
        CK(cudaEventRecord(e1, st));
        const float ms = time_ms(st, e0, e1);
        if (rep) printf("%s grid %3d : %.3f us per barrier\n", variant == 0 ? "B fence+atomicAdd, spin   " : "C red.release, spin acquire",
                        grid, ms * 1e3 / iters);
      }
    }
  }
  // D: barrier + realistic phase
  {
    const int grid = sms;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaMemsetAsync(counter, 0, 4, st));
      CK(cudaEventRecord(e0, st));
      int it = iters;
      void* args[] = {&it, &counter, &buf, &errors};
      CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_red_phase), dim3(grid), dim3(256), args, 0, st));
      CK(cudaEventRecord(e1, st));
      const float ms = time_ms(st, e0, e1);
      if (rep) printf("D persistent: 8 KB write -> barrier -> 8 KB read   grid %3d : %.3f us per phase\n", grid, ms * 1e3 / iters);
    }
  }
  // E: the same phase as a chain of kernels
  for (int pdl = 0; pdl < 2; ++pdl) {
    const int grid = sms;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0, st));
      launch_chain(st, iters, grid, buf, errors, pdl != 0);
      CK(cudaEventRecord(e1, st));
      const float ms = time_ms(st, e0, e1);
      if (rep) printf("E kernel per phase, stream launches, %s : %.3f us per phase\n", pdl ? "PDL  " : "plain", ms * 1e3 / iters);
    }
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    launch_chain(st, iters, grid, buf, errors, pdl != 0);
    CK(cudaStreamEndCapture(st, &graph));
    CK(cudaGraphInstantiate(&exec, graph, 0));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0, st));
      CK(cudaGraphLaunch(exec, st));
      CK(cudaEventRecord(e1, st));
      const float ms = time_ms(st, e0, e1);
      if (rep) printf("E kernel per phase, CUDA graph replay,  %s : %.3f us per phase\n", pdl ? "PDL  " : "plain", ms * 1e3 / iters);
    }
    CK(cudaGraphExecDestroy(exec));
    CK(cudaGraphDestroy(graph));
  }
  int herr = 0;
  CK(cudaMemcpy(&herr, errors, 4, cudaMemcpyDeviceToHost));
  printf("visibility errors: %d\n", herr);
  return herr != 0;
}
