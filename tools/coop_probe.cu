// coop_probe.cu — what the platform allows for a persistent multi-phase decode kernel:
//  (1) latency of a software grid barrier across one CTA per SM,
//  (2) whether a cooperative launch can be combined with programmatic dependent launch,
//  (3) whether such launches survive stream capture into a CUDA graph,
//  (4) whether two such grids on different streams deadlock each other (bounded spins).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/coop_probe tools/coop_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ unsigned g_cnt[4], g_gen[4];
__device__ unsigned g_timeouts;

__device__ __forceinline__ uint64_t gtime() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void grid_barrier(int which, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned gen = *(volatile unsigned*)&g_gen[which];
    if (atomicAdd(&g_cnt[which], 1u) == n - 1) {
      g_cnt[which] = 0;
      __threadfence();
      atomicAdd(&g_gen[which], 1u);
    } else {
      const uint64_t t0 = gtime();
      while (*(volatile unsigned*)&g_gen[which] == gen) {
        if (gtime() - t0 > 2000000000ull) { atomicAdd(&g_timeouts, 1u); break; }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(288, 1) barrier_kernel(int which, int iters, int pdl, float* sink) {
  extern __shared__ uint8_t smem[];
  if (pdl) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    grid_barrier(which, gridDim.x);
    acc += smem[(threadIdx.x + i) & 1023];
  }
  if (acc == 12345.f) sink[0] = acc;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("  %s -> %s\n", #x, cudaGetErrorString(e_)); } } while (0)

static cudaError_t launch(cudaStream_t st, int grid, bool coop, bool pdl, int which, int iters, float* sink) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(288);
  cfg.dynamicSmemBytes = 200 * 1024;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (coop) { at[n].id = cudaLaunchAttributeCooperative; at[n].val.cooperative = 1; ++n; }
  if (pdl) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, barrier_kernel, which, iters, pdl ? 1 : 0, sink);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, cooperativeLaunch=%d\n", prop.name, sms, prop.cooperativeLaunch);
  CK(cudaFuncSetAttribute(barrier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  float* sink;
  cudaMalloc(&sink, 4);
  cudaStream_t s1, s2;
  cudaStreamCreate(&s1);
  cudaStreamCreate(&s2);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float ms;
  for (int mode = 0; mode < 4; ++mode) {
    const bool coop = mode & 1, pdl = mode & 2;
    cudaError_t e = launch(s1, sms, coop, pdl, 0, 10, sink);
    cudaError_t e2 = cudaStreamSynchronize(s1);
    printf("eager launch coop=%d pdl=%d: launch=%s sync=%s\n", coop, pdl, cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaGetLastError();
    if (e != cudaSuccess || e2 != cudaSuccess) continue;
    cudaEventRecord(a, s1);
    launch(s1, sms, coop, pdl, 0, 2000, sink);
    cudaEventRecord(b, s1);
    cudaStreamSynchronize(s1);
    cudaEventElapsedTime(&ms, a, b);
    printf("   grid barrier: %.3f us each (2000 barriers, %d CTAs x 288 threads)\n", ms * 1000.f / 2000.f, sms);
    // graph capture of 8 back-to-back launches
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaError_t c0 = cudaStreamBeginCapture(s1, cudaStreamCaptureModeThreadLocal);
    cudaError_t cl = cudaSuccess;
    for (int i = 0; i < 8 && cl == cudaSuccess; ++i) cl = launch(s1, sms, coop, pdl, 0, 20, sink);
    cudaError_t c1 = cudaStreamEndCapture(s1, &g);
    printf("   capture: begin=%s launch=%s end=%s\n", cudaGetErrorString(c0), cudaGetErrorString(cl), cudaGetErrorString(c1));
    if (c1 == cudaSuccess && cl == cudaSuccess) {
      cudaError_t ci = cudaGraphInstantiate(&ge, g, 0);
      printf("   instantiate=%s\n", cudaGetErrorString(ci));
      if (ci == cudaSuccess) {
        cudaGraphLaunch(ge, s1);
        cudaStreamSynchronize(s1);
        cudaEventRecord(a, s1);
        for (int r = 0; r < 20; ++r) cudaGraphLaunch(ge, s1);
        cudaEventRecord(b, s1);
        cudaError_t cs = cudaStreamSynchronize(s1);
        cudaEventElapsedTime(&ms, a, b);
        printf("   graph replay: %s, %.2f us per kernel of 20 barriers\n", cudaGetErrorString(cs), ms * 1000.f / (20 * 8));
        cudaGraphExecDestroy(ge);
      }
      cudaGraphDestroy(g);
    }
    cudaGetLastError();
  }
  // two grids on two streams at once
  for (int coop = 0; coop < 2; ++coop) {
    unsigned zero = 0;
    cudaMemcpyToSymbol(g_timeouts, &zero, 4);
    cudaEventRecord(a, s1);
    for (int r = 0; r < 20; ++r) {
      launch(s1, sms, coop, false, 1, 200, sink);
      launch(s2, sms, coop, false, 2, 200, sink);
    }
    cudaError_t e1 = cudaStreamSynchronize(s1), e2 = cudaStreamSynchronize(s2);
    cudaEventRecord(b, s1);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    unsigned to = 0;
    cudaMemcpyFromSymbol(&to, g_timeouts, 4);
    printf("two streams, coop=%d: %s %s, %.1f ms total, barrier timeouts=%u\n", coop, cudaGetErrorString(e1),
           cudaGetErrorString(e2), ms, to);
  }
  return 0;
}
