// common.cuh — shared device helpers for the sm_100a kernels: bf16 packing,
// mbarrier / TMA / tcgen05 inline PTX, the counter-based RNG of the sampler.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace advspec {

constexpr int kWarp = 32;

// ---------------------------------------------------------------------------
// bf16 helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float round_bf16(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

// 16-byte streaming load that does not allocate in L1 (weights are read once).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// One MUFU.EX2 (exp2f() adds a range test and two scalings per element for denormal results, which a
// probability that is about to be rounded to bf16 does not need; -inf -> 0).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Programmatic dependent launch: a kernel launched with the PDL attribute may
// start while its predecessor drains; everything that reads the predecessor's
// output must come after pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---------------------------------------------------------------------------
// Sampler RNG: splitmix64-style mix of (seed, step, vocab index) -> u in (0,1).
// oracle/sampling_ref.py restates this bit for bit in numpy.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ float uniform01(uint64_t seed, uint32_t step, uint32_t idx) {
  uint64_t h = mix64(seed ^ mix64(((uint64_t)step << 32) | (uint64_t)idx));
  uint32_t m = (uint32_t)(h >> 40);                  // 24 random bits
  return ((float)m + 0.5f) * (1.0f / 16777216.0f);   // (0,1), never 0 or 1
}

// ---------------------------------------------------------------------------
// In-situ kernel timeline: when enabled, block 0 of every decode-path kernel stamps
// %globaltimer on entry.  Kernels of a step are serialised by data dependencies, so the
// difference of consecutive stamps is each kernel's real cost (run time + launch gap)
// inside the CUDA-graph replay — something neither ncu (serialised, cold cache) nor
// host-side events (CPU-bound for 5 us kernels) can see.
// ---------------------------------------------------------------------------
constexpr int kTraceCap = 8192;
__device__ unsigned long long g_ktrace[kTraceCap];
__device__ unsigned int g_ktrace_n = 0;
__device__ int g_ktrace_on = 0;
__device__ int g_l2_evict_first = 1;  // ADVSPEC_L2_EVICT_FIRST=0 turns the weight stream's L2 evict_first hint off (A/B)
enum TraceKind : int { TK_GEMV = 1, TK_ATTN = 2, TK_SAMPLE = 3, TK_ROPE = 4, TK_COMBINE = 5, TK_SAMPLE_SCAN = 6 };
__device__ __forceinline__ void ktrace_mark(int kind) {
  if (g_ktrace_on && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned int i = atomicAdd(&g_ktrace_n, 1u);
    if (i < (unsigned int)kTraceCap) g_ktrace[i] = (t << 4) | (unsigned long long)kind;
  }
}

// Phase stamps of ONE CTA (blockIdx.x == g_phase_cta) of the most recent launch of an instrumented
// kernel: where inside the kernel the time goes (diagnostic; only written while tracing is on).
__device__ unsigned long long g_phase[16];
__device__ int g_phase_cta = 1;
__device__ __forceinline__ void phase_mark(int i) {
  if (g_ktrace_on && blockIdx.x == (unsigned)g_phase_cta && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_phase[i] = t;
  }
}

// ---------------------------------------------------------------------------
// shared-memory addresses, mbarrier
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Device-side watchdog word: kernels that spin on an mbarrier give up after a
// bounded number of polls, record where, and trap — a wrong barrier protocol
// must surface as an error code, never as a hung GPU.
__device__ unsigned int g_watchdog_code = 0;

__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Returns false (and latches g_watchdog_code) if the phase did not complete
// within ~2 s; once the code is latched every later wait gives up at once, so
// a broken pipeline drains in milliseconds and the host reports
// ADVSPEC_ERR_KERNEL instead of the box hanging.
__device__ __noinline__ bool mbar_wait_slow(uint64_t* bar, uint32_t parity, uint32_t site) {
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      if (*(volatile unsigned int*)&g_watchdog_code != 0) return false;
      if (global_timer_ns() - t0 > 2000000000ull) {
        atomicCAS(&g_watchdog_code, 0u, 0x80000000u | site);
        return false;
      }
    }
  }
  return true;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, uint32_t site) {
  if (mbar_try_wait(bar, parity)) return true;
  return mbar_wait_slow(bar, parity, site);
}

// ---------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — SASS: UTMALDG
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// ---------------------------------------------------------------------------
// tcgen05 / TMEM — SASS: UTCHMMA, LDTM, UTCBAR
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread retire.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i gets TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 tile stored as rows of
// 128 bytes (64 elements) with the 128-byte swizzle TMA applies:
//   start address >> 4 in bits [0,14); LBO (unused for swizzled K-major) = 0;
//   SBO = 1024 B (8 rows x 128 B) >> 4 in bits [32,46); descriptor version 1
//   in bits [46,48); layout type SWIZZLE_128B (= 2) in bits [61,64).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16: fp32 accumulator (bits [4,6) = 1), bf16
// A and B (bits [7,10) and [10,13) = 1), both K-major (bits 15,16 = 0),
// N >> 3 in bits [17,23), M >> 4 in bits [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

}  // namespace advspec

namespace advspec {
// 1-D bulk async copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :
      : "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// Same copy with an L2 eviction policy (createpolicy): a 16 GB/step weight stream marked evict_first
// leaves the small data every layer re-reads (activations, norm weights, tensor maps, RoPE rows,
// attention partials) resident in the 126 MB L2 instead of flushing it once per layer.
__device__ __forceinline__ void bulk_load_1d_hint(void* smem_dst, const void* gsrc, uint32_t bytes,
                                                  uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      :
      : "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// L2 prefetch hint for the 128-byte line holding gptr (fire and forget)
__device__ __forceinline__ void prefetch_l2_line(const void* gptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(gptr) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Reduce up to 32 independent values across the 32 lanes of a warp with 31 shuffles
// (recursive halving); on return lane L holds the warp-wide sum of v[L].
__device__ __forceinline__ float warp_reduce_32vals(float (&v)[32], int lane) {
#pragma unroll
  for (int n = 32, o = 16; o >= 1; n >>= 1, o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = upper ? v[i + n / 2] : v[i];
      const float send = upper ? v[i] : v[i + n / 2];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}
}  // namespace advspec
