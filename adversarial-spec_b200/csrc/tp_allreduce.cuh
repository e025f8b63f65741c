// tp_allreduce.cuh — the decode step's tensor-parallel exchange over NVLink peer memory.
//
// What is exchanged: the b x d_model fp32 residual partials after the o-proj and the down-proj
// (16-256 KB), 2 x n_layers times per token.  At that size a library all-reduce is pure latency
// (measured: ~40 us each through NCCL inside the captured step, 160 of them per Llama-3-70B token).
// Here every rank PUSHES its partial into a slot of every peer's buffer (peer memory mapped through
// CUDA IPC; NVSwitch gives all peers full bandwidth at once), raises a flag there, waits for the tp
// flags in its own buffer and sums the tp slots in rank order — so every rank computes bit-identical
// sums.  One launch of G CTAs; CTA i moves and reduces slice i only, so there is no grid-wide step.
// (tp_allreduce_kernel below is this flag protocol, kept for A/B; the engine runs tp_allreduce_ll_kernel,
// which carries the flag inside the data.)
//
// Layout of a rank's region:  flags u32 [2 sets][tp][kArCtas] (first kArFlagBytes), then
// slots f32 [2 sets][tp][max_elems].  Two sets alternate by generation: a rank can be one exchange
// ahead of a slow peer (never two: exchange g+1 cannot finish without the slow peer's push, which
// comes after it has read g), so g and g+1 must not share slots.
#pragma once

#include "common.cuh"

namespace advspec {

constexpr int kArCtas = 8;
constexpr int kArThreads = 512;
constexpr int kArFlagBytes = 1024;
constexpr int kArMaxRanks = 8;

struct ArParams {
  uint8_t* peer[kArMaxRanks];  // every rank's region as mapped in THIS process (peer[rank] is local)
  float* data;                 // [n] in: this rank's partial; out: the sum over ranks
  int n;                       // multiple of 4
  int tp, rank;
  int64_t max_elems;
  unsigned int* gen;           // [kArCtas] generations completed by each CTA (device state: graph replays advance it)
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kArThreads) tp_allreduce_kernel(ArParams p) {
  __shared__ unsigned int s_gen;
  const int tid = threadIdx.x, i = blockIdx.x;
  pdl_launch_dependents();  // the next kernel may stage its weights; its own wait covers this one
  pdl_wait();               // the partial comes from the GEMV before us
  if (tid == 0) s_gen = p.gen[i] + 1;
  __syncthreads();
  const unsigned int gen = s_gen;
  const int set = (int)(gen & 1u);
  const int n4 = p.n >> 2;
  const int lo = (int)(((int64_t)n4 * i) / gridDim.x), hi = (int)(((int64_t)n4 * (i + 1)) / gridDim.x);
  auto slot = [&](int owner, int src) {
    return reinterpret_cast<float4*>(p.peer[owner] + kArFlagBytes) + ((int64_t)(set * p.tp + src) * p.max_elems >> 2);
  };
  auto flag = [&](int owner, int src) {
    return reinterpret_cast<unsigned int*>(p.peer[owner]) + (set * p.tp + src) * kArCtas + i;
  };
  // 1. push this rank's slice to every rank (its own included)
  const float4* mine = reinterpret_cast<const float4*>(p.data);
  for (int idx = lo + tid; idx < hi; idx += kArThreads) {
    const float4 v = __ldcg(mine + idx);
#pragma unroll 1
    for (int r = 0; r < p.tp; ++r) slot(r, p.rank)[idx] = v;
  }
  __syncthreads();
  if (tid < p.tp) {
    __threadfence_system();  // the CTA's stores (ordered before this thread by the barrier) before the flag
    st_release_sys(flag(tid, p.rank), gen);
  }
  // 2. wait for every rank's slice in the local region
  if (tid < p.tp) {
    const unsigned int* f = flag(p.rank, tid);
    const uint64_t t0 = global_timer_ns();
    uint32_t spins = 0;
    while (ld_acquire_sys(f) != gen) {
      if ((++spins & 63u) == 0) {
        if (*(volatile unsigned int*)&g_watchdog_code != 0) break;
        if (global_timer_ns() - t0 > 2000000000ull) {
          atomicCAS(&g_watchdog_code, 0u, 0x80000A00u | (unsigned)tid);
          break;
        }
      }
    }
  }
  __syncthreads();
  // 3. sum in rank order (identical on every rank), bypassing L1: the slots are written by peers
  float4* out = reinterpret_cast<float4*>(p.data);
  for (int idx = lo + tid; idx < hi; idx += kArThreads) {
    float4 a = __ldcg(slot(p.rank, 0) + idx);
    for (int r = 1; r < p.tp; ++r) {
      const float4 v = __ldcg(slot(p.rank, r) + idx);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    out[idx] = a;
  }
  if (tid == 0) p.gen[i] = gen;
}

// ---------------------------------------------------------------------------------------------------
// The same exchange with the flag INSIDE the data ("low-latency" packets): every float travels as an 8-byte
// {bits, generation} pair, so a receiver polls the data words themselves — no system-scope fence waiting for
// the remote writes to be acknowledged, no separate flag store behind it.  Costs 2x the bytes (64 KB instead
// of 32 KB per exchange at d_model 8192), saves the fence + flag round trip (several us at 8 ranks).
// Slot layout of the LL region: uint2 [2 sets][tp][max_elems] after the same kArFlagBytes header.
// 16-byte stores carry two packets; each 8-byte packet is valid on its own, so a torn 16-byte store is harmless.
__global__ void __launch_bounds__(kArThreads) tp_allreduce_ll_kernel(ArParams p) {
  __shared__ unsigned int s_gen;
  const int tid = threadIdx.x, i = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (tid == 0) s_gen = p.gen[i] + 1;
  __syncthreads();
  const unsigned int gen = s_gen;
  const int set = (int)(gen & 1u);
  const int n4 = p.n >> 2;
  const int lo = (int)(((int64_t)n4 * i) / gridDim.x), hi = (int)(((int64_t)n4 * (i + 1)) / gridDim.x);
  auto slot = [&](int owner, int src) {  // uint4 = two packets = two floats
    return reinterpret_cast<uint4*>(p.peer[owner] + kArFlagBytes) + ((int64_t)(set * p.tp + src) * p.max_elems >> 1);
  };
  const float4* mine = reinterpret_cast<const float4*>(p.data);
  float4* out = reinterpret_cast<float4*>(p.data);
  for (int idx = lo + tid; idx < hi; idx += kArThreads) {
    const float4 v = __ldcg(mine + idx);
    const uint4 a = make_uint4(__float_as_uint(v.x), gen, __float_as_uint(v.y), gen);
    const uint4 b = make_uint4(__float_as_uint(v.z), gen, __float_as_uint(v.w), gen);
#pragma unroll 1
    for (int r = 0; r < p.tp; ++r) {
      uint4* dst = slot(r, p.rank) + 2 * idx;
      asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
      asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 1), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
    }
  }
  // wait for and sum every rank's packets of this CTA's slice, in rank order (identical sums on every rank)
  for (int idx = lo + tid; idx < hi; idx += kArThreads) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < p.tp; ++r) {
      const uint4* src = slot(p.rank, r) + 2 * idx;
      uint4 a, b;
      const uint64_t t0 = global_timer_ns();
      uint32_t spins = 0;
      while (true) {
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(src) : "memory");
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(src + 1) : "memory");
        if (a.y == gen && a.w == gen && b.y == gen && b.w == gen) break;
        if ((++spins & 63u) == 0) {
          if (*(volatile unsigned int*)&g_watchdog_code != 0) break;
          if (global_timer_ns() - t0 > 2000000000ull) {
            atomicCAS(&g_watchdog_code, 0u, 0x80000B00u | (unsigned)r);
            break;
          }
        }
      }
      acc.x += __uint_as_float(a.x);
      acc.y += __uint_as_float(a.z);
      acc.z += __uint_as_float(b.x);
      acc.w += __uint_as_float(b.z);
    }
    out[idx] = acc;
  }
  __syncthreads();
  if (tid == 0) p.gen[i] = gen;
}

}  // namespace advspec
