// gemv_chain.cuh — several dependent decode GEMVs in ONE persistent launch.
//
// A decode step is a chain of small (N x K) x (K x b) products; as separate kernels each pays its
// own launch, pipeline fill and drain (in-situ: 2.7-6.3 us per kernel on top of the HBM time, with
// HBM idle in between).  Here the chain o-proj -> gate/up -> down -> next layer's qkv (or lm_head)
// runs as phases of one kernel, one CTA per SM, same tile machinery as gemv_mma.cuh:
//   * the producer warp never stops at a phase boundary: weights are constants, so it keeps
//     filling the ring with the NEXT phase's first tiles while the consumers finish the current
//     phase, cross the grid barrier and stage their new input — HBM keeps streaming;
//   * the consumer warps separate phases with a software grid barrier (1.9 us measured across
//     148 CTAs, tools/coop_probe.cu); the launch is cooperative so all CTAs are co-resident.
// Inputs produced inside the launch by other CTAs are read with ld.global.cg (L2): this SM's L1
// may hold a copy from an earlier phase.
#pragma once

#include "../../adversarial-spec_b200/csrc/gemv_mma.cuh"

namespace advspec {

constexpr int kChainMaxPhases = 4;

struct ChainPhase {
  const __nv_bfloat16* W;
  const void* x;        // in_mode 0: bf16 [b][K]; in_mode 1: f32 [b][K] (fused RMSNorm)
  const float* norm_w;  // in_mode 1
  const float* bias;    // EPI_BF16 only, may be null
  void* y;
  int N, K;
  int in_mode, epilogue;
  int x_in_smem;        // in_mode 0: stage x in shared memory (else B fragments are read through L1)
};

struct ChainParams {
  ChainPhase ph[kChainMaxPhases];
  int n_ph;
  int act;
  float eps;
  unsigned int* bar;  // [0] arrivals, [1] generation; zero-initialised once, self-resetting
};

__device__ __forceinline__ float4 ldcg_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// Barrier across the consumer warps of every CTA of the grid (the producer warps do not take part).
__device__ __forceinline__ void chain_grid_barrier(unsigned int* bar, int tid) {
  named_bar_sync(1, kGmConsumers);
  if (tid == 0) {
    __threadfence();
    const unsigned int gen = *(volatile unsigned int*)&bar[1];
    if (atomicAdd(&bar[0], 1u) == gridDim.x - 1) {
      bar[0] = 0;
      __threadfence();
      atomicAdd(&bar[1], 1u);
    } else {
      const uint64_t t0 = global_timer_ns();
      uint32_t spins = 0;
      while (*(volatile unsigned int*)&bar[1] == gen) {
        if ((++spins & 255u) == 0) {
          if (*(volatile unsigned int*)&g_watchdog_code != 0) break;
          if (global_timer_ns() - t0 > 2000000000ull) {
            atomicCAS(&g_watchdog_code, 0u, 0x80000900u);
            break;
          }
        }
      }
    }
    __threadfence();
  }
  named_bar_sync(1, kGmConsumers);
}

template <int B>
__global__ void __launch_bounds__(kGmThreads, 1) gemv_chain_kernel(ChainParams p, int n_stages) {
  constexpr int KC = kGmKCDefault, RT = kGmRTDefault;
  constexpr int kRowPitch = GmCfg<KC, RT>::kRowPitch;
  constexpr int kStage = GmCfg<KC, RT>::kStageBytes;
  constexpr int kWC = GmCfg<KC, RT>::kWarpCols;
  constexpr int kSteps = GmCfg<KC, RT>::kSteps;
  extern __shared__ __align__(128) uint8_t gc_smem[];
  uint8_t* ring = gc_smem;
  uint8_t* xs_raw = gc_smem + (size_t)n_stages * kStage;
  __shared__ uint64_t full_bar[kGmMaxStages], empty_bar[kGmMaxStages];
  __shared__ float s_part[2][8][RT][8];
  __shared__ float s_red[8][B];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  ktrace_mark(TK_GEMV);

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (!g_ktrace_on) pdl_launch_dependents();

  auto rows_of = [&](const ChainPhase& q, int& rb0, int& re0) {
    const int pairs = (q.N + 1) / 2;
    rb0 = (int)(((int64_t)pairs * blockIdx.x) / gridDim.x) * 2;
    re0 = min(q.N, (int)(((int64_t)pairs * (blockIdx.x + 1)) / gridDim.x) * 2);
  };

  if (warp == 8) {
    // ------------------------------ producer: one uninterrupted stream over all phases
    int tg = 0;  // tiles issued so far (ring position)
    for (int pi = 0; pi < p.n_ph; ++pi) {
      const ChainPhase& q = p.ph[pi];
      int row_begin, row_end;
      rows_of(q, row_begin, row_end);
      const int n_rb = (row_end - row_begin + RT - 1) / RT;
      const int n_kc = (q.K + KC - 1) / KC;
      for (int t = 0; t < n_rb * n_kc; ++t, ++tg) {
        const int s = tg % n_stages;
        const uint32_t ph = (uint32_t)(tg / n_stages) & 1u;
        const int rb = row_begin + (t / n_kc) * RT;
        const int kc = (t % n_kc) * KC;
        const int rows = min(RT, row_end - rb);
        const uint32_t cbytes = (uint32_t)min(KC, q.K - kc) * 2u;
        if (lane == 0) {
          mbar_wait(&empty_bar[s], ph ^ 1u, 0x500u + s);
          mbar_arrive_expect_tx(&full_bar[s], cbytes * (uint32_t)rows);
        }
        __syncwarp();
        if (lane < rows)
          bulk_load_1d(ring + (size_t)s * kStage + (size_t)lane * kRowPitch,
                       q.W + (int64_t)(rb + lane) * q.K + kc, cbytes, &full_bar[s]);
      }
    }
    pdl_wait();
  } else {
    // ------------------------------ consumers
    phase_mark(0);
    pdl_wait();
    phase_mark(1);
    const int n_opp = lane >> 2, t4 = lane & 3;
    const bool opp_ok = n_opp < B;
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int a_col = (lane >> 4) * 8;
    int tg = 0;
    for (int pi = 0; pi < p.n_ph; ++pi) {
      const ChainPhase& q = p.ph[pi];
      constexpr int kMaxV = 4;
      const bool in_regs = q.K <= kMaxV * kGmConsumers * 4;
      float4 nw[kMaxV];  // RMSNorm weights: constants, fetched BEFORE the barrier (cold DRAM reads queue
                         // behind the weight stream for ~2.4 us)
      if (q.in_mode == 1 && in_regs) {
#pragma unroll
        for (int i = 0; i < kMaxV; ++i) {
          const int k = (tid + i * kGmConsumers) * 4;
          nw[i] = (k < q.K) ? *reinterpret_cast<const float4*>(q.norm_w + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (pi > 0) {
        chain_grid_barrier(p.bar, tid);
        phase_mark(1 + 3 * pi);
        ktrace_mark(TK_GEMV);  // block 0 stamps the phase start: timelines read like the unchained step
      }
      int row_begin, row_end;
      rows_of(q, row_begin, row_end);
      const int n_rb = (row_end - row_begin + RT - 1) / RT;
      const int n_kc = (q.K + KC - 1) / KC;
      const int xpitch = q.K * 2 + 16;

      // ---- stage this phase's input
      const uint8_t* xbase;
      int xstride;
      if (q.in_mode == 1) {
        const float* xf = reinterpret_cast<const float*>(q.x);
        float4 xv[B][kMaxV];
        float ss[B];
#pragma unroll
        for (int b = 0; b < B; ++b) ss[b] = 0.f;
        if (in_regs) {
#pragma unroll
          for (int i = 0; i < kMaxV; ++i) {
            const int k = (tid + i * kGmConsumers) * 4;
#pragma unroll
            for (int b = 0; b < B; ++b) {
              xv[b][i] = (k < q.K) ? ldcg_f4(xf + (int64_t)b * q.K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
              ss[b] += xv[b][i].x * xv[b][i].x + xv[b][i].y * xv[b][i].y + xv[b][i].z * xv[b][i].z +
                       xv[b][i].w * xv[b][i].w;
            }
          }
        } else {
          for (int k = tid * 4; k < q.K; k += kGmConsumers * 4) {
#pragma unroll
            for (int b = 0; b < B; ++b) {
              const float4 v = ldcg_f4(xf + (int64_t)b * q.K + k);
              ss[b] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
          }
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const float t = warp_sum(ss[b]);
          if (lane == 0) s_red[warp][b] = t;
        }
        named_bar_sync(1, kGmConsumers);
        float inv[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) t += s_red[w][b];
          inv[b] = rsqrtf(t / (float)q.K + p.eps);
        }
        if (in_regs) {
#pragma unroll
          for (int i = 0; i < kMaxV; ++i) {
            const int k = (tid + i * kGmConsumers) * 4;
            if (k < q.K) {
              const float4 w4 = nw[i];
#pragma unroll
              for (int b = 0; b < B; ++b) {
                uint2 o;
                o.x = pack_bf16(xv[b][i].x * inv[b] * w4.x, xv[b][i].y * inv[b] * w4.y);
                o.y = pack_bf16(xv[b][i].z * inv[b] * w4.z, xv[b][i].w * inv[b] * w4.w);
                *reinterpret_cast<uint2*>(xs_raw + (size_t)b * xpitch + (size_t)k * 2) = o;
              }
            }
          }
        } else {
          for (int k = tid * 4; k < q.K; k += kGmConsumers * 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(q.norm_w + k);
#pragma unroll
            for (int b = 0; b < B; ++b) {
              const float4 v = ldcg_f4(xf + (int64_t)b * q.K + k);
              uint2 o;
              o.x = pack_bf16(v.x * inv[b] * w4.x, v.y * inv[b] * w4.y);
              o.y = pack_bf16(v.z * inv[b] * w4.z, v.w * inv[b] * w4.w);
              *reinterpret_cast<uint2*>(xs_raw + (size_t)b * xpitch + (size_t)k * 2) = o;
            }
          }
        }
        named_bar_sync(1, kGmConsumers);
        xbase = xs_raw;
        xstride = xpitch;
      } else if (q.x_in_smem) {
        const __nv_bfloat16* xg = reinterpret_cast<const __nv_bfloat16*>(q.x);
        for (int i = tid * 8; i < B * q.K; i += kGmConsumers * 8) {
          const int b = i / q.K, k = i % q.K;
          *reinterpret_cast<uint4*>(xs_raw + (size_t)b * xpitch + (size_t)k * 2) =
              __ldcg(reinterpret_cast<const uint4*>(xg + i));
        }
        named_bar_sync(1, kGmConsumers);
        xbase = xs_raw;
        xstride = xpitch;
      } else {
        // read through L1: safe because no phase of this launch has touched these addresses before
        // (a chain never reads the same global input in two phases)
        xbase = reinterpret_cast<const uint8_t*>(q.x);
        xstride = q.K * 2;
      }
      const uint8_t* xrow = xbase + (size_t)(opp_ok ? n_opp : 0) * xstride;
      phase_mark(2 + 3 * pi);

      // ---- the phase's tiles
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int t = 0; t < n_rb * n_kc; ++t, ++tg) {
        const int s = tg % n_stages;
        const uint32_t ph = (uint32_t)(tg / n_stages) & 1u;
        const int rbi = t / n_kc;
        const int kci = t % n_kc;
        const int col0 = kci * KC + warp * kWC;
        uint32_t bf[kSteps][2];
#pragma unroll
        for (int ks = 0; ks < kSteps; ++ks) {
          const int k = col0 + ks * 16 + 2 * t4;
          const bool ok = opp_ok && (col0 + ks * 16) < q.K;
          bf[ks][0] = ok ? *reinterpret_cast<const uint32_t*>(xrow + (size_t)k * 2) : 0u;
          bf[ks][1] = ok ? *reinterpret_cast<const uint32_t*>(xrow + (size_t)(k + 8) * 2) : 0u;
        }
        mbar_wait(&full_bar[s], ph, 0x600u + s);
        const uint8_t* tile = ring + (size_t)s * kStage + (size_t)a_row * kRowPitch + (size_t)(warp * kWC + a_col) * 2;
#pragma unroll
        for (int ks = 0; ks < kSteps; ++ks) {
          if (col0 + ks * 16 < q.K) {
            uint32_t a[4];
            ldmatrix_x4(a, tile + ks * 32);
            mma_bf16_16816(acc, a, bf[ks][0], bf[ks][1]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);

        if (kci == n_kc - 1) {
          const int rb = row_begin + rbi * RT;
          const int rows = min(RT, row_end - rb);
          const int buf = rbi & 1;
          const int g = lane >> 2;
          *reinterpret_cast<float2*>(&s_part[buf][warp][g][2 * t4]) = make_float2(acc[0], acc[1]);
          *reinterpret_cast<float2*>(&s_part[buf][warp][g + 8][2 * t4]) = make_float2(acc[2], acc[3]);
          acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
          named_bar_sync(1, kGmConsumers);
          if (q.epilogue == EPI_GATED_BF16) {
            if (tid < (rows / 2) * B) {
              const int pr = tid / B, b = tid % B;
              float gs = 0.f, us = 0.f;
#pragma unroll
              for (int w = 0; w < 8; ++w) {
                gs += s_part[buf][w][2 * pr][b];
                us += s_part[buf][w][2 * pr + 1][b];
              }
              reinterpret_cast<__nv_bfloat16*>(q.y)[(int64_t)b * (q.N / 2) + (rb >> 1) + pr] =
                  __float2bfloat16_rn(apply_act(gs, p.act) * us);
            }
          } else if (tid < rows * B) {
            const int r = tid / B, b = tid % B;
            float tsum = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) tsum += s_part[buf][w][r][b];
            const int n = rb + r;
            if (q.epilogue == EPI_BF16) {
              if (q.bias) tsum += q.bias[n];
              reinterpret_cast<__nv_bfloat16*>(q.y)[(int64_t)b * q.N + n] = __float2bfloat16_rn(tsum);
            } else if (q.epilogue == EPI_RESADD_F32) {
              float* yp = reinterpret_cast<float*>(q.y) + (int64_t)b * q.N + n;
              *yp = __ldcg(yp) + tsum;
            } else {
              reinterpret_cast<float*>(q.y)[(int64_t)b * q.N + n] = tsum;
            }
          }
        }
      }
      phase_mark(3 + 3 * pi);
    }
  }
  pdl_launch_dependents();
}

}  // namespace advspec
