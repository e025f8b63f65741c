// gemv_mma.cuh — the decode GEMV, third design: a bulk-async weight stream consumed by
// tensor cores.  With b <= 8 opponents the step's matmuls are (N x K) x (K x b) GEMMs; on
// CUDA cores they cost ~33 issue slots per 16 bytes of W (unpack + b FMAs per element),
// which starves the SM's 4 schedulers long before HBM is saturated.  Here the b opponents
// are the N=8 dimension of mma.sync.m16n8k16 (bf16 in, fp32 accumulate): one ldmatrix.x4
// + one MMA per 512 bytes of W.
//
// One persistent CTA per SM, 288 threads.  Warp 8 is the producer: tiles of 16 rows x 2048
// columns, one 4 KB 1-D bulk async copy per row (marked L2 evict_first) into a shared-memory
// ring whose rows are padded by 16 bytes (conflict-free ldmatrix); the ring is primed before
// griddepcontrol.wait, and so are the constants the consumers need after it (norm weights, bias).
// Warps 0..7 consume: warp w owns columns [w*256, w*256+256) of each tile, accumulates the
// 16 x 8 tile of outputs over all column chunks of a row block, then the 8 warps' partials
// are summed in shared memory and the fused epilogue (bias / residual add / gated
// activation / fp32 store) writes the rows.  Algorithmic bytes per launch: N*K*2.
#pragma once

#include "attn.cuh"  // ldmatrix / mma wrappers
#include "common.cuh"
#include "decode_kernels.cuh"

namespace advspec {

constexpr int kGmRTDefault = 16;    // rows per tile (8 = upper half of the MMA tile zero; measured slower)
constexpr int kGmKCDefault = 2048;  // columns per tile: one 4 KB bulk copy per row
constexpr int kGmMaxStages = 8;
template <int KC, int RT = kGmRTDefault>
struct GmCfg {
  static constexpr int kRowPitch = KC * 2 + 16;  // bytes; +16 staggers rows across banks
  static constexpr int kStageBytes = RT * kRowPitch;
  static constexpr int kWarpCols = KC / 8;       // columns of a tile owned by one consumer warp
  static constexpr int kSteps = kWarpCols / 16;  // MMA k-steps per warp per tile
};
constexpr int kGmStageBytes = GmCfg<kGmKCDefault, kGmRTDefault>::kStageBytes;  // 65,792
constexpr int kGmConsumers = 256;
constexpr int kGmThreads = 288;

template <int B, int KC = kGmKCDefault, int RT = kGmRTDefault>
__global__ void __launch_bounds__(kGmThreads, 1) gemv_mma_kernel(GemvParams p, int n_stages, int x_in_smem) {
  static_assert(RT == 8 || RT == 16, "a tile is half or all of the 16-row MMA M dimension");
  constexpr int kGmKC = KC;
  constexpr int kGmRT = RT;
  constexpr int kGmRowPitch = GmCfg<KC, RT>::kRowPitch;
  constexpr int kGmStageBytes = GmCfg<KC, RT>::kStageBytes;
  constexpr int kWC = GmCfg<KC, RT>::kWarpCols;
  constexpr int kSteps = GmCfg<KC, RT>::kSteps;
  extern __shared__ __align__(128) uint8_t gm_smem[];
  uint8_t* ring = gm_smem;
  uint8_t* xs_raw = gm_smem + (size_t)n_stages * kGmStageBytes;  // bf16 [B][K] with pitch K*2+16
  const int xpitch = p.K * 2 + 16;
  __shared__ uint64_t full_bar[kGmMaxStages], empty_bar[kGmMaxStages];
  __shared__ float s_part[2][8][RT][8];
  __shared__ float s_red[8][B];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  ktrace_mark(TK_GEMV);
  const int pairs = (p.N + 1) / 2;
  const int row_begin = (int)(((int64_t)pairs * blockIdx.x) / gridDim.x) * 2;
  const int row_end = min(p.N, (int)(((int64_t)pairs * (blockIdx.x + 1)) / gridDim.x) * 2);
  const int n_rb = (row_end - row_begin + kGmRT - 1) / kGmRT;
  const int n_kc = (p.K + kGmKC - 1) / kGmKC;
  const int n_tiles = n_rb * n_kc;

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    fence_mbar_init();
  }
  __syncthreads();
  // Let the next kernel's CTAs be scheduled as soon as SMs free up: everything they do before their
  // own griddepcontrol.wait (priming their weight ring) is independent of this kernel's output.
  if (!g_ktrace_on) pdl_launch_dependents();  // while tracing, kernels trigger at their end so that
                                              // start-to-start stamps are clean per-kernel costs
  if (warp == 8) {
    // ------------------------------ producer ------------------------------
    const uint64_t pol = l2_policy_evict_first();
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % n_stages;
      const uint32_t ph = (uint32_t)(t / n_stages) & 1u;
      if (t == n_stages) pdl_wait();  // the ring is primed; from here on we depend on the consumers anyway
      const int rb = row_begin + (t / n_kc) * kGmRT;
      const int kc = (t % n_kc) * kGmKC;
      const int rows = min(kGmRT, row_end - rb);
      const uint32_t cbytes = (uint32_t)min(kGmKC, p.K - kc) * 2u;
      if (lane == 0) {
        mbar_wait(&empty_bar[s], ph ^ 1u, 0x500u + s);
        mbar_arrive_expect_tx(&full_bar[s], cbytes * (uint32_t)rows);
      }
      __syncwarp();
      if (lane < rows) {  // rows <= RT
        if (g_l2_evict_first)
          bulk_load_1d_hint(ring + (size_t)s * kGmStageBytes + (size_t)lane * kGmRowPitch,
                            p.W + (int64_t)(rb + lane) * p.K + kc, cbytes, &full_bar[s], pol);
        else
          bulk_load_1d(ring + (size_t)s * kGmStageBytes + (size_t)lane * kGmRowPitch,
                       p.W + (int64_t)(rb + lane) * p.K + kc, cbytes, &full_bar[s]);
      }
    }
    pdl_wait();
  } else {
    // ------------------------------ consumers -----------------------------
    // Constants this kernel needs AFTER the dependency wait are fetched BEFORE it: a cold DRAM read
    // issued behind a saturated weight stream waits ~2.4 us in the HBM queue (measured), and the
    // RMSNorm weights (and the bias) are cold at every layer.
    constexpr int kMaxV = 4;  // float4 per opponent per thread held in registers: K <= 4096
    const bool in_regs = p.K <= kMaxV * kGmConsumers * 4;
    float4 nw[kMaxV];
    if (p.in_mode == 1) {
      if (in_regs) {
#pragma unroll
        for (int i = 0; i < kMaxV; ++i) {
          const int k = (tid + i * kGmConsumers) * 4;
          nw[i] = (k < p.K) ? *reinterpret_cast<const float4*>(p.norm_w + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        for (int k = tid * 32; k < p.K; k += kGmConsumers * 32) prefetch_l2_line(p.norm_w + k);
      }
    }
    if (p.bias != nullptr && tid * 32 < row_end - row_begin) prefetch_l2_line(p.bias + row_begin + tid * 32);
    pdl_wait();
    const uint8_t* xbase;  // bf16 rows of x, pitch xstride bytes
    int xstride;
    if (p.in_mode == 1) {
      // Fused RMSNorm of the fp32 residual stream.  When b*K is small (every model's d_model at b <= 4)
      // the row slice of each thread stays in registers between the sum of squares and the scaling.
      const float* xf = reinterpret_cast<const float*>(p.x);
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
            xv[b][i] = (k < p.K) ? *reinterpret_cast<const float4*>(xf + (int64_t)b * p.K + k)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            ss[b] += xv[b][i].x * xv[b][i].x + xv[b][i].y * xv[b][i].y + xv[b][i].z * xv[b][i].z +
                     xv[b][i].w * xv[b][i].w;
          }
        }
      } else {
        for (int k = tid * 4; k < p.K; k += kGmConsumers * 4) {
#pragma unroll
          for (int b = 0; b < B; ++b) {
            const float4 v = *reinterpret_cast<const float4*>(xf + (int64_t)b * p.K + k);
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
        for (int w = 0; w < 8; ++w) t += s_red[w][b];  // same fixed order in every thread
        inv[b] = rsqrtf(t / (float)p.K + p.eps);
      }
      if (in_regs) {
#pragma unroll
        for (int i = 0; i < kMaxV; ++i) {
          const int k = (tid + i * kGmConsumers) * 4;
          if (k < p.K) {
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
        for (int k = tid * 4; k < p.K; k += kGmConsumers * 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(p.norm_w + k);
#pragma unroll
          for (int b = 0; b < B; ++b) {
            const float4 v = *reinterpret_cast<const float4*>(xf + (int64_t)b * p.K + k);
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
    } else if (x_in_smem) {
      const __nv_bfloat16* xg = reinterpret_cast<const __nv_bfloat16*>(p.x);
      for (int i = tid * 8; i < B * p.K; i += kGmConsumers * 8) {
        const int b = i / p.K, k = i % p.K;  // K % 8 == 0: a 16-byte chunk never straddles rows
        *reinterpret_cast<uint4*>(xs_raw + (size_t)b * xpitch + (size_t)k * 2) =
            *reinterpret_cast<const uint4*>(xg + i);
      }
      named_bar_sync(1, kGmConsumers);
      xbase = xs_raw;
      xstride = xpitch;
    } else {
      xbase = reinterpret_cast<const uint8_t*>(p.x);
      xstride = p.K * 2;
    }

    const int n_opp = lane >> 2, t4 = lane & 3;  // this lane's column (opponent) of the B fragment
    const bool opp_ok = n_opp < B;
    const uint8_t* xrow = xbase + (size_t)(opp_ok ? n_opp : 0) * xstride;
    // ldmatrix source row/column of this lane inside a 16 x 16 A block
    // ldmatrix source of this lane inside a 16 x 16 A block (RT == 8: a 8 x 16 block, x2 load)
    const int a_row = (RT == 16) ? (lane & 7) + ((lane >> 3) & 1) * 8 : (lane & 7);
    const int a_col = (RT == 16) ? (lane >> 4) * 8 : ((lane >> 3) & 1) * 8;

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % n_stages;
      const uint32_t ph = (uint32_t)(t / n_stages) & 1u;
      const int rbi = t / n_kc;
      const int kci = t % n_kc;
      const int col0 = kci * kGmKC + warp * kWC;  // first column of this warp's slice
      // B fragments (x) for the 8 k-steps of the slice
      uint32_t bf[kSteps][2];
#pragma unroll
      for (int ks = 0; ks < kSteps; ++ks) {
        const int k = col0 + ks * 16 + 2 * t4;
        const bool ok = opp_ok && (col0 + ks * 16) < p.K;
        bf[ks][0] = ok ? *reinterpret_cast<const uint32_t*>(xrow + (size_t)k * 2) : 0u;
        bf[ks][1] = ok ? *reinterpret_cast<const uint32_t*>(xrow + (size_t)(k + 8) * 2) : 0u;
      }
      mbar_wait(&full_bar[s], ph, 0x600u + s);
      const uint8_t* tile = ring + (size_t)s * kGmStageBytes + (size_t)a_row * kGmRowPitch +
                            (size_t)(warp * kWC + a_col) * 2;
#pragma unroll
      for (int ks = 0; ks < kSteps; ++ks) {
        if (col0 + ks * 16 < p.K) {  // warp-uniform
          uint32_t a[4];
          if constexpr (RT == 16) {
            ldmatrix_x4(a, tile + ks * 32);
          } else {
            uint32_t lo, hi;
            asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                         : "=r"(lo), "=r"(hi)
                         : "r"(smem_u32(tile + ks * 32)));
            a[0] = lo; a[1] = 0u; a[2] = hi; a[3] = 0u;  // rows 8..15 of the MMA tile are zero
          }
          mma_bf16_16816(acc, a, bf[ks][0], bf[ks][1]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);

      if (kci == n_kc - 1) {
        const int rb = row_begin + rbi * kGmRT;
        const int rows = min(kGmRT, row_end - rb);
        const int buf = rbi & 1;
        const int g = lane >> 2;
        // accumulator layout: acc[0],acc[1] = (row g, opponents 2*t4, 2*t4+1); acc[2],acc[3] = row g+8
        *reinterpret_cast<float2*>(&s_part[buf][warp][g][2 * t4]) = make_float2(acc[0], acc[1]);
        if constexpr (RT == 16)
          *reinterpret_cast<float2*>(&s_part[buf][warp][g + 8][2 * t4]) = make_float2(acc[2], acc[3]);
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        named_bar_sync(1, kGmConsumers);
        if (p.epilogue == EPI_GATED_BF16) {
          if (tid < (rows / 2) * B) {
            const int pr = tid / B, b = tid % B;
            float gs = 0.f, us = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
              gs += s_part[buf][w][2 * pr][b];
              us += s_part[buf][w][2 * pr + 1][b];
            }
            reinterpret_cast<__nv_bfloat16*>(p.y)[(int64_t)b * (p.N / 2) + (rb >> 1) + pr] =
                __float2bfloat16_rn(apply_act(gs, p.act) * us);
          }
        } else if (tid < rows * B) {
          const int r = tid / B, b = tid % B;
          float tsum = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) tsum += s_part[buf][w][r][b];
          const int n = rb + r;
          if (p.epilogue == EPI_BF16) {
            if (p.bias) tsum += p.bias[n];
            reinterpret_cast<__nv_bfloat16*>(p.y)[(int64_t)b * p.N + n] = __float2bfloat16_rn(tsum);
          } else if (p.epilogue == EPI_RESADD_F32) {
            reinterpret_cast<float*>(p.y)[(int64_t)b * p.N + n] += tsum;
          } else {
            reinterpret_cast<float*>(p.y)[(int64_t)b * p.N + n] = tsum;
          }
        }
      }
    }
  }
  pdl_launch_dependents();
}

}  // namespace advspec
