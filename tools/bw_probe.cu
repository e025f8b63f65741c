// How fast can ANY kernel stream a small buffer?  Read-only grid-stride kernels over distinct
// buffers (total >> L2), back to back, with and without programmatic dependent launch.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}
template <int UNROLL, bool PDL>
__global__ void __launch_bounds__(512) reader(const uint4* __restrict__ p, size_t n, uint4* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = ldg_stream(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
  }
  for (; i < n; i += stride) { uint4 v = ldg_stream(p + i); acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; }
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (acc.x == 0x12345678u) *sink = acc;
  if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <int UNROLL, bool PDL>
float run(const std::vector<uint4*>& bufs, size_t bytes, int grid, int block, uint4* sink) {
  cudaEvent_t a, z; cudaEventCreate(&a); cudaEventCreate(&z);
  float best = 1e9;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    for (auto* b : bufs) {
      if (PDL) {
        cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = 0;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const uint4* pb = b; size_t n = bytes / 16;
        CK(cudaLaunchKernelEx(&cfg, reader<UNROLL, PDL>, pb, n, sink));
      } else {
        reader<UNROLL, PDL><<<grid, block>>>(b, bytes / 16, sink);
      }
    }
    cudaEventRecord(z); CK(cudaEventSynchronize(z));
    float ms; cudaEventElapsedTime(&ms, a, z); best = ms < best ? ms : best;
  }
  return best / bufs.size() * 1e3f;
}
int main() {
  uint4* sink; CK(cudaMalloc(&sink, 16));
  const double sizes_mb[] = {33.55, 50.33, 117.44, 234.88, 1050.7};
  for (double mb : sizes_mb) {
    size_t bytes = (size_t)(mb * 1e6) / 4096 * 4096;
    int nb = (int)std::max(4.0, std::min(48.0, 6e9 / bytes));
    std::vector<uint4*> bufs(nb);
    for (auto& b : bufs) { CK(cudaMalloc(&b, bytes)); cudaMemset(b, 1, bytes); }
    double ideal = bytes / 6489.9e9 * 1e6;
    printf("%.1f MB (ideal %.2f us @6490):\n", bytes / 1e6, ideal);
    struct { const char* n; float t; } rows[] = {
      {"grid 148x4  blk 512 unroll 1      ", run<1, false>(bufs, bytes, 592, 512, sink)},
      {"grid 148x4  blk 512 unroll 4      ", run<4, false>(bufs, bytes, 592, 512, sink)},
      {"grid 148x2  blk 256 unroll 8      ", run<8, false>(bufs, bytes, 296, 256, sink)},
      {"grid 148x1  blk 256 unroll 8      ", run<8, false>(bufs, bytes, 148, 256, sink)},
      {"grid 148x4  blk 512 unroll 4  PDL ", run<4, true>(bufs, bytes, 592, 512, sink)},
      {"grid 148x2  blk 256 unroll 8  PDL ", run<8, true>(bufs, bytes, 296, 256, sink)},
    };
    for (auto& r : rows) printf("   %s: %7.2f us  %5.0f GB/s\n", r.n, r.t, bytes / r.t / 1e3);
    for (auto b : bufs) cudaFree(b);
  }
  return 0;
}
