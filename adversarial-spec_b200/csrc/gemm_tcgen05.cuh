// gemm_tcgen05.cuh — the prefill GEMM: C[M,N] = A[M,K] * B[N,K]^T, bf16 in,
// fp32 accumulate in TMEM, fused epilogues.  A = activations of the shared
// prompt (staged ONCE via TMA whatever the number of opponents), B = weights.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0 (1 lane)  TMA producer: cp.async.bulk.tensor 128B-swizzled tiles -> smem ring
//   warp 1 (1 lane)  tcgen05.mma issuer: 128 x BN x 16 UMMA, accumulators in TMEM
//   warp 2           TMEM allocate / free
//   warps 4..7       epilogue: tcgen05.ld -> registers -> fused op -> global
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of
// tile i+1.  Roofline: tensor pipe (2*M*N*K flops).
// Tile order (GemmSched): M-fastest inside BANDS of m-tiles sized so that a band of A stays in L2 while
// the weight tiles stream past it once per band (the 5,068 x 14,336 down-proj input is 147 MB: walked
// M-fastest over all 40 m-tiles it was re-read from DRAM every wave — 963 MB of reads for 345 MB).
// fp32 epilogues (residual add / plain store) may split the tiles of the last, partial wave along K
// among the SMs that would otherwise idle; the splits of a tile add into C in split order (a per-tile
// semaphore), so the result does not depend on timing.
#pragma once

#include "common.cuh"

namespace advspec {

enum GemmEpilogue : int {
  EPI_BF16 = 0,        // C bf16 [M,N] = acc (+ bias[n])
  EPI_RESADD_F32 = 1,  // C f32  [M,N] += acc            (residual stream)
  EPI_GATED_BF16 = 2,  // C bf16 [M,N/2] = act(acc[2i]) * acc[2i+1]
  EPI_F32 = 3          // C f32  [M,N] = acc
};
enum Activation : int { ACT_SILU = 0, ACT_GELU_TANH = 1 };

__device__ __forceinline__ float act_silu(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float act_gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f;  // sqrt(2/pi)
  return 0.5f * x * (1.0f + tanhf(k0 * (x + 0.044715f * x * x * x)));
}
__device__ __forceinline__ float apply_act(float x, int act) {
  return act == ACT_GELU_TANH ? act_gelu_tanh(x) : act_silu(x);
}

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kGemmThreads = 256;

struct GemmParams {
  void* C;
  int64_t ldc;        // elements of C's row pitch
  const float* bias;  // EPI_BF16 only, may be null
  int M, N, K;
  int act;
  // schedule (filled by the host launcher, see gemm_schedule in engine.cu)
  int band_m;         // m-tiles per raster band (>= 1)
  int full_items;     // work items [0, full_items) are whole tiles
  int split;          // tiles past full_items are cut into `split` K-ranges each (1 = none)
  int total_items;    // full_items + (tiles - full_items) * split
  unsigned int* sem;  // [tiles - full_items] zero-initialised ordering words (split > 1 only)
};

// One work item of the persistent loop: an output tile and the K-blocks [kb0, kb1) it accumulates.
struct GemmWork {
  int m0, n0, kb0, kb1, q, tail_idx;
};
template <int BN>
__device__ __forceinline__ GemmWork gemm_work(const GemmParams& p, int w, int num_m, int num_n, int num_kb) {
  GemmWork r;
  int t;
  if (w < p.full_items) {
    t = w;
    r.q = 0;
    r.kb0 = 0;
    r.kb1 = num_kb;
    r.tail_idx = -1;
  } else {
    const int rr = w - p.full_items;
    r.tail_idx = rr / p.split;
    r.q = rr % p.split;
    t = p.full_items + r.tail_idx;
    r.kb0 = (int)((int64_t)num_kb * r.q / p.split);
    r.kb1 = (int)((int64_t)num_kb * (r.q + 1) / p.split);
  }
  const int per_band = p.band_m * num_n;
  const int b = t / per_band, within = t - b * per_band;
  const int h = min(p.band_m, num_m - b * p.band_m);
  r.m0 = (b * p.band_m + within % h) * kGemmBM;
  r.n0 = (within / h) * BN;
  return r;
}

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024: manual alignment
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                       // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;            // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;        // [2]        MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kStages + 2;   // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + kGemmBM - 1) / kGemmBM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + kGemmBK - 1) / kGemmBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------ TMA producer ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
      const GemmWork wk = gemm_work<BN>(p, w, num_m, num_n, num_kb);
      for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u, 0x100u + stage);
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
        tma_load_2d(smemA + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * kGemmBK, wk.m0);
        tma_load_2d(smemB + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * kGemmBK, wk.n0);
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = make_idesc_bf16(kGemmBM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++iter) {
      const GemmWork wk = gemm_work<BN>(p, w, num_m, num_n, num_kb);
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1u, 0x200u + acc);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase, 0x300u + stage);
        tc_fence_after();
        const uint64_t da = make_smem_desc_sw128(smem_u32(smemA + stage * Cfg::kABytes));
        const uint64_t db = make_smem_desc_sw128(smem_u32(smemB + stage * Cfg::kBBytes));
#pragma unroll
        for (int k = 0; k < kGemmBK / 16; ++k) {
          // advance 16 bf16 = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
          tc_mma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                     (kb != wk.kb0 || k != 0) ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
        if (kb == wk.kb1 - 1) tc_commit(&tfull_bar[acc]);
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ----------------------------------
    const int quad = warp - 4;  // == warp % 4: the TMEM lane quadrant this warp may read
    int iter = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++iter) {
      const GemmWork wk = gemm_work<BN>(p, w, num_m, num_n, num_kb);
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const int m0 = wk.m0, n0 = wk.n0;
      mbar_wait(&tfull_bar[acc], acc_phase, 0x400u + acc);
      tc_fence_after();
      // K-split tail tile: the splits add into C in split order (deterministic sums)
      bool add_into = (EPI == EPI_RESADD_F32);
      bool from_split = false;  // C holds another CTA's partial sums of this tile: read it past L1
      if constexpr (EPI == EPI_RESADD_F32 || EPI == EPI_F32) {
        if (wk.tail_idx >= 0 && p.split > 1) {
          if (wk.q > 0) {
            add_into = true;
            from_split = true;
            if (lane == 0) {
              const uint64_t t0 = global_timer_ns();
              while (*(volatile unsigned int*)&p.sem[wk.tail_idx] != (unsigned int)wk.q) {
                if (*(volatile unsigned int*)&g_watchdog_code != 0) break;
                if (global_timer_ns() - t0 > 2000000000ull) {
                  atomicCAS(&g_watchdog_code, 0u, 0x80000000u | 0x500u);
                  break;
                }
              }
            }
            __syncwarp();
            __threadfence();  // acquire the earlier splits' stores
          }
        }
      }
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the per-lane stores
        tmem_ld_32x32(taddr0 + (uint32_t)c0, v);
        tmem_ld_wait();
        const int n = n0 + c0;
        if (!row_ok || n >= p.N) continue;
        const bool full = (n + 32 <= p.N);
        if constexpr (EPI == EPI_BF16) {
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)row * p.ldc + n;
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < p.N) v[j] = __float_as_uint(__uint_as_float(v[j]) + p.bias[n + j]);
          }
          if (full && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 o;
              o.x = pack_bf16(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]));
              o.y = pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              o.z = pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
              o.w = pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
              *reinterpret_cast<uint4*>(out + j) = o;
            }
          } else {
            for (int j = 0; j < 32 && n + j < p.N; ++j)
              out[j] = __float2bfloat16_rn(__uint_as_float(v[j]));
          }
        } else if constexpr (EPI == EPI_RESADD_F32 || EPI == EPI_F32) {
          float* out = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + n;
          // (thread-per-row 16-byte accesses: each thread walks its own 128-byte line, which L1 serves; routing the
          // chunk through shared memory for fully coalesced rows measured 4 ms SLOWER per prefill)
          if (full && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o;
              if (add_into) {
                o = from_split ? __ldcg(reinterpret_cast<const float4*>(out + j)) : *reinterpret_cast<const float4*>(out + j);
              } else {
                o = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              o.x += __uint_as_float(v[j + 0]);
              o.y += __uint_as_float(v[j + 1]);
              o.z += __uint_as_float(v[j + 2]);
              o.w += __uint_as_float(v[j + 3]);
              *reinterpret_cast<float4*>(out + j) = o;
            }
          } else {
            for (int j = 0; j < 32 && n + j < p.N; ++j) {
              const float base = add_into ? __ldcg(out + j) : 0.f;
              out[j] = base + __uint_as_float(v[j]);
            }
          }
        } else {  // EPI_GATED_BF16: columns (2i, 2i+1) = (gate_i, up_i)
          __nv_bfloat16* out =
              reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)row * p.ldc + (n >> 1);
          float h[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            h[j] = apply_act(__uint_as_float(v[2 * j]), p.act) * __uint_as_float(v[2 * j + 1]);
          if (full && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              uint4 o;
              o.x = pack_bf16(h[j + 0], h[j + 1]);
              o.y = pack_bf16(h[j + 2], h[j + 3]);
              o.z = pack_bf16(h[j + 4], h[j + 5]);
              o.w = pack_bf16(h[j + 6], h[j + 7]);
              *reinterpret_cast<uint4*>(out + j) = o;
            }
          } else {
            for (int j = 0; j < 16 && n + 2 * j + 1 < p.N; ++j) out[j] = __float2bfloat16_rn(h[j]);
          }
        }
      }
      // all of this thread's TMEM reads for the tile are complete (tmem_ld_wait above)
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if constexpr (EPI == EPI_RESADD_F32 || EPI == EPI_F32) {
        if (wk.tail_idx >= 0 && p.split > 1) {
          // the four epilogue warps own disjoint rows of the tile: each publishes its rows, the last of
          // the four (named barrier) hands the tile to the next split, the final split re-zeroes the word
          __threadfence();
          named_bar_sync(1, 128);
          if (warp == 4 && lane == 0) {
            __threadfence();
            *(volatile unsigned int*)&p.sem[wk.tail_idx] = (wk.q + 1 == p.split) ? 0u : (unsigned int)(wk.q + 1);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Plain CUDA-core GEMM with the same epilogues: the on-device check for the
// tcgen05 kernel (tests/ only; never launched by the engine's product path).
template <int EPI>
__global__ void gemm_check_kernel(const __nv_bfloat16* __restrict__ A, int64_t lda,
                                  const __nv_bfloat16* __restrict__ B, int64_t ldb,
                                  GemmParams p) {
  const int ncols = (EPI == EPI_GATED_BF16) ? p.N / 2 : p.N;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)p.M * ncols) return;
  const int m = (int)(idx / ncols);
  const int c = (int)(idx % ncols);
  auto dot = [&](int n) {
    float acc = 0.f;
    const __nv_bfloat16* a = A + (int64_t)m * lda;
    const __nv_bfloat16* b = B + (int64_t)n * ldb;
    for (int k = 0; k < p.K; ++k) acc = fmaf(__bfloat162float(a[k]), __bfloat162float(b[k]), acc);
    return acc;
  };
  if constexpr (EPI == EPI_BF16) {
    float v = dot(c);
    if (p.bias) v += p.bias[c];
    reinterpret_cast<__nv_bfloat16*>(p.C)[(int64_t)m * p.ldc + c] = __float2bfloat16_rn(v);
  } else if constexpr (EPI == EPI_RESADD_F32) {
    reinterpret_cast<float*>(p.C)[(int64_t)m * p.ldc + c] += dot(c);
  } else if constexpr (EPI == EPI_F32) {
    reinterpret_cast<float*>(p.C)[(int64_t)m * p.ldc + c] = dot(c);
  } else {
    const float g = dot(2 * c), u = dot(2 * c + 1);
    reinterpret_cast<__nv_bfloat16*>(p.C)[(int64_t)m * p.ldc + c] =
        __float2bfloat16_rn(apply_act(g, p.act) * u);
  }
}

}  // namespace advspec
