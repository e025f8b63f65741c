// Micro-benchmark of the weight-streaming GEMV variants on matrices >> L2.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iadversarial-spec_b200/csrc -o tools/gemv_bench tools/gemv_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../adversarial-spec_b200/csrc/gemv_mma.cuh"
#include "ab_kernels/gemv_stream.cuh"
#include "ab_kernels/gemv_cpasync.cuh"
#include "ab_kernels/gemv_rmma.cuh"
using namespace advspec;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int KC, int RT = 16>
float run_mma(const GemvParams& p0, const std::vector<__nv_bfloat16*>& Ws, int stages, int reps) {
  constexpr int B = 3;
  auto kern = gemv_mma_kernel<B, KC, RT>;
  const size_t xb = ((size_t)B * ((size_t)p0.K * 2 + 16) + 127) / 128 * 128;
  const size_t dyn = (size_t)stages * GmCfg<KC, RT>::kStageBytes + xb;
  if (dyn > 216 * 1024) return -1.f;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  cudaEvent_t a, z; cudaEventCreate(&a); cudaEventCreate(&z);
  float best = 1e9;
  for (int rep = 0; rep < reps; ++rep) {
    cudaEventRecord(a);
    for (auto* W : Ws) { GemvParams p = p0; p.W = W; kern<<<148, kGmThreads, dyn>>>(p, stages, 0); }
    cudaEventRecord(z); CK(cudaEventSynchronize(z));
    float ms; cudaEventElapsedTime(&ms, a, z); best = ms < best ? ms : best;
  }
  CK(cudaGetLastError());
  return best / Ws.size() * 1e3f;
}
template <int NST>
float run_cpa(const GemvParams& p0, const std::vector<__nv_bfloat16*>& Ws, int reps, int grid = 148) {
  constexpr int B = 3;
  auto kern = gemv_cpasync_kernel<B, NST>;
  const size_t xb = ((size_t)B * ((size_t)p0.K * 2 + 16) + 127) / 128 * 128;
  const bool xin = p0.in_mode == 1 || (8 * NST * kGcWarpStage + xb <= 216 * 1024);
  const size_t dyn = (size_t)8 * NST * kGcWarpStage + (xin ? xb : 0);
  if (dyn > 216 * 1024) return -1.f;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  cudaEvent_t a, z; cudaEventCreate(&a); cudaEventCreate(&z);
  float best = 1e9;
  for (int rep = 0; rep < reps; ++rep) {
    cudaEventRecord(a);
    for (auto* W : Ws) { GemvParams p = p0; p.W = W; kern<<<grid, kGcThreads, dyn>>>(p, (p0.in_mode != 1 && xin) ? 1 : 0); }
    cudaEventRecord(z); CK(cudaEventSynchronize(z));
    float ms; cudaEventElapsedTime(&ms, a, z); best = ms < best ? ms : best;
  }
  CK(cudaGetLastError());
  return best / Ws.size() * 1e3f;
}
float run_rmma(const GemvParams& p0, const std::vector<__nv_bfloat16*>& Ws, int reps, bool pdl) {
  constexpr int B = 3;
  auto kern = gemv_rmma_kernel<B>;
  const size_t xb = ((size_t)B * ((size_t)p0.K * 2 + 64) + 127) / 128 * 128;
  const size_t dyn = xb + (size_t)kGrWarps * kGrRBG * kGrRT * B * 4;
  if (dyn > 110 * 1024) return -1.f;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  cudaEvent_t a, z; cudaEventCreate(&a); cudaEventCreate(&z);
  float best = 1e9;
  for (int rep = 0; rep < reps; ++rep) {
    cudaEventRecord(a);
    for (auto* W : Ws) {
      GemvParams p = p0; p.W = W;
      cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(296); cfg.blockDim = dim3(kGrThreads); cfg.dynamicSmemBytes = dyn; cfg.stream = 0;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
      CK(cudaLaunchKernelEx(&cfg, kern, p));
    }
    cudaEventRecord(z); CK(cudaEventSynchronize(z));
    float ms; cudaEventElapsedTime(&ms, a, z); best = ms < best ? ms : best;
  }
  CK(cudaGetLastError());
  return best / Ws.size() * 1e3f;
}
float run_v1(const GemvParams& p0, const std::vector<__nv_bfloat16*>& Ws, int reps) {
  cudaEvent_t a, z; cudaEventCreate(&a); cudaEventCreate(&z);
  float best = 1e9;
  for (int rep = 0; rep < reps; ++rep) {
    cudaEventRecord(a);
    for (auto* W : Ws) { GemvParams p = p0; p.W = W; gemv_kernel<3, 2><<<296, 256>>>(p); }
    cudaEventRecord(z); CK(cudaEventSynchronize(z));
    float ms; cudaEventElapsedTime(&ms, a, z); best = ms < best ? ms : best;
  }
  CK(cudaGetLastError());
  return best / Ws.size() * 1e3f;
}
__global__ void read_only_kernel(const uint4* __restrict__ p, size_t n, uint4* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = ldg_stream(p + i); acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  if (acc.x == 0x12345678u) *sink = acc;
}
int main() {
  const int shapes[5][2] = {{6144, 4096}, {4096, 4096}, {28672, 4096}, {4096, 14336}, {128256, 4096}};
  const char* names[5] = {"qkv", "o", "gate_up", "down", "lm_head"};
  float *x; __nv_bfloat16* xb; float* nw; void* y;
  CK(cudaMalloc(&x, 3 * 14336 * 4)); CK(cudaMalloc(&xb, 3 * 14336 * 2)); CK(cudaMalloc(&nw, 14336 * 4)); CK(cudaMalloc(&y, 3 * 128256 * 4));
  cudaMemset(x, 0, 3 * 14336 * 4); cudaMemset(xb, 0, 3 * 14336 * 2); cudaMemset(nw, 0, 14336 * 4); cudaMemset(y, 0, 3 * 128256 * 4);
  // pure read bandwidth reference
  {
    size_t bytes = (size_t)4 << 30; uint4* buf; uint4* sink; CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 16)); cudaMemset(buf, 1, bytes);
    cudaEvent_t a, z; cudaEventCreate(&a); cudaEventCreate(&z);
    for (int g : {148 * 4, 148 * 8, 148 * 16}) {
      float best = 1e9;
      for (int r = 0; r < 3; ++r) { cudaEventRecord(a); read_only_kernel<<<g, 512>>>(buf, bytes / 16, sink); cudaEventRecord(z); cudaEventSynchronize(z); float ms; cudaEventElapsedTime(&ms, a, z); best = ms < best ? ms : best; }
      printf("read-only stream grid=%d: %.1f GB/s\n", g, bytes / best / 1e6);
    }
    cudaFree(buf);
  }
  {
    cudaEvent_t a, z; cudaEventCreate(&a); cudaEventCreate(&z);
    cudaEventRecord(a); for (int i = 0; i < 200; ++i) read_only_kernel<<<148, 256>>>(nullptr, 0, nullptr); cudaEventRecord(z); cudaEventSynchronize(z);
    float ms; cudaEventElapsedTime(&ms, a, z); printf("empty kernel back-to-back: %.2f us per launch\n", ms / 200 * 1e3);
  }
  for (int si = 0; si < 5; ++si) {
    const int N = shapes[si][0], K = shapes[si][1];
    const size_t wbytes = (size_t)N * K * 2;
    const int nW = si == 4 ? 6 : 24;
    std::vector<__nv_bfloat16*> Ws(nW);
    for (auto& w : Ws) { CK(cudaMalloc(&w, wbytes)); cudaMemset(w, 0, wbytes); }
    const bool norm = (si == 0 || si == 2 || si == 4);
    GemvParams p{nullptr, norm ? (void*)x : (void*)xb, norm ? nw : nullptr, nullptr, y, N, K, norm ? 1 : 0,
                 si == 2 ? EPI_GATED_BF16 : (si == 4 ? EPI_F32 : (norm ? EPI_BF16 : EPI_RESADD_F32)), 0, 1e-5f};
    const double ideal = wbytes / 6489.9e9 * 1e6;
    printf("%-8s N=%d K=%d  %.1f MB  ideal@6490GB/s %.2f us\n", names[si], N, K, wbytes / 1e6, ideal);
    float t1 = run_v1(p, Ws, 3);
    printf("   v1 (LDG regs)            : %7.2f us  %.0f GB/s\n", t1, wbytes / t1 / 1e3);
    for (int st : {4}) { float t = run_mma<1024>(p, Ws, st, 3); if (t > 0) printf("   mma KC=1024 stages=%d     : %7.2f us  %.0f GB/s\n", st, t, wbytes / t / 1e3); }
    for (int st : {2, 3}) { float t = run_mma<2048>(p, Ws, st, 3); if (t > 0) printf("   mma KC=2048 stages=%d     : %7.2f us  %.0f GB/s\n", st, t, wbytes / t / 1e3); }
    for (int st : {4, 5, 6}) { float t = run_mma<2048, 8>(p, Ws, st, 3); if (t > 0) printf("   mma KC=2048 RT=8 st=%d    : %7.2f us  %.0f GB/s\n", st, t, wbytes / t / 1e3); }
    for (int st : {2, 3}) { float t = run_mma<4096, 8>(p, Ws, st, 3); if (t > 0) printf("   mma KC=4096 RT=8 st=%d    : %7.2f us  %.0f GB/s\n", st, t, wbytes / t / 1e3); }
    { float t = run_rmma(p, Ws, 3, false); if (t > 0) printf("   rmma (LDG->mma) no PDL   : %7.2f us  %.0f GB/s\n", t, wbytes / t / 1e3); }
    { float t = run_rmma(p, Ws, 3, true); if (t > 0) printf("   rmma (LDG->mma) PDL      : %7.2f us  %.0f GB/s\n", t, wbytes / t / 1e3); }


    for (auto w : Ws) cudaFree(w);
  }
  return 0;
}
