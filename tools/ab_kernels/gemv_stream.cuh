// gemv_stream.cuh — the decode GEMV as a producer/consumer weight stream.
//
// One persistent CTA per SM.  Warp 8 (one lane) is the producer: it walks this
// CTA's rows in tiles of RT rows x KC columns and pulls each row segment into a
// shared-memory ring with 1-D bulk async copies (cp.async.bulk, completion on an
// mbarrier) — up to ~190 KB in flight per SM, independent of register pressure,
// and it starts BEFORE griddepcontrol.wait because weights never depend on the
// previous kernel.  Warps 0..7 are consumers: warp w owns columns
// [w*256, w*256+256) of every tile (lane = 16 bytes), keeps the matching slice of
// x for all b opponents in registers, accumulates RT x b partial sums while the
// tile sequence stays on the same rows, then reduces them with a 31-shuffle
// multi-value butterfly.  Algorithmic bytes per launch: N*K*2, independent of b.
#pragma once

#include "../../adversarial-spec_b200/csrc/common.cuh"
#include "../../adversarial-spec_b200/csrc/decode_kernels.cuh"

namespace advspec {

constexpr int kGsRT = 8;                         // rows per tile
constexpr int kGsKC = 2048;                      // columns per tile
constexpr int kGsStageBytes = kGsRT * kGsKC * 2;  // 32 KB
constexpr int kGsMaxStages = 6;
constexpr int kGsConsumers = 256;
constexpr int kGsThreads = 288;

template <int B>
__global__ void __launch_bounds__(kGsThreads, 1) gemv_stream_kernel(GemvParams p, int n_stages,
                                                                    int x_in_smem) {
  extern __shared__ __align__(128) uint8_t gs_smem[];
  uint8_t* ring = gs_smem;
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(gs_smem + (size_t)n_stages * kGsStageBytes);
  __shared__ uint64_t full_bar[kGsMaxStages], empty_bar[kGsMaxStages];
  __shared__ float s_part[2][8][kGsRT][B];
  __shared__ float s_red[8][B];
  __shared__ float s_inv[B];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  ktrace_mark(TK_GEMV);
  const int pairs = (p.N + 1) / 2;
  const int row_begin = (int)(((int64_t)pairs * blockIdx.x) / gridDim.x) * 2;
  const int row_end = min(p.N, (int)(((int64_t)pairs * (blockIdx.x + 1)) / gridDim.x) * 2);
  const int n_rb = (row_end - row_begin + kGsRT - 1) / kGsRT;
  const int n_kc = (p.K + kGsKC - 1) / kGsKC;
  const int n_tiles = n_rb * n_kc;

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == 8) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % n_stages;
        const uint32_t ph = (uint32_t)(t / n_stages) & 1u;
        if (t == n_stages) pdl_wait();  // ring is primed; nothing else to do before the dependency
        mbar_wait(&empty_bar[s], ph ^ 1u, 0x500u + s);
        const int rb = row_begin + (t / n_kc) * kGsRT;
        const int kc = (t % n_kc) * kGsKC;
        const int rows = min(kGsRT, row_end - rb);
        const uint32_t cbytes = (uint32_t)min(kGsKC, p.K - kc) * 2u;
        mbar_arrive_expect_tx(&full_bar[s], cbytes * (uint32_t)rows);
        uint8_t* dst = ring + (size_t)s * kGsStageBytes;
        const __nv_bfloat16* src = p.W + (int64_t)rb * p.K + kc;
        for (int r = 0; r < rows; ++r)
          bulk_load_1d(dst + (size_t)r * (kGsKC * 2), src + (int64_t)r * p.K, cbytes, &full_bar[s]);
      }
    }
    pdl_wait();  // every thread observes the dependency before the kernel can complete
  } else {
    // ------------------------------ consumers -----------------------------
    pdl_wait();
    const __nv_bfloat16* xsrc;  // bf16 [B][K]: shared copy, or the caller's global buffer
    if (p.in_mode == 1) {
      const float* xf = reinterpret_cast<const float*>(p.x);
      float ss[B];
#pragma unroll
      for (int b = 0; b < B; ++b) ss[b] = 0.f;
      for (int k = tid * 4; k < p.K; k += kGsConsumers * 4) {
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const float4 v = *reinterpret_cast<const float4*>(xf + (int64_t)b * p.K + k);
          ss[b] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
      }
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const float t = warp_sum(ss[b]);
        if (lane == 0) s_red[warp][b] = t;
      }
      named_bar_sync(1, kGsConsumers);
      if (tid < B) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += s_red[w][tid];
        s_inv[tid] = rsqrtf(t / (float)p.K + p.eps);
      }
      named_bar_sync(1, kGsConsumers);
      for (int k = tid * 4; k < p.K; k += kGsConsumers * 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(p.norm_w + k);
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const float4 v = *reinterpret_cast<const float4*>(xf + (int64_t)b * p.K + k);
          const float inv = s_inv[b];
          uint2 o;
          o.x = pack_bf16(v.x * inv * w4.x, v.y * inv * w4.y);
          o.y = pack_bf16(v.z * inv * w4.z, v.w * inv * w4.w);
          *reinterpret_cast<uint2*>(xs + (int64_t)b * p.K + k) = o;
        }
      }
      named_bar_sync(1, kGsConsumers);
      xsrc = xs;
    } else if (x_in_smem) {
      const __nv_bfloat16* xg = reinterpret_cast<const __nv_bfloat16*>(p.x);
      for (int i = tid * 8; i < B * p.K; i += kGsConsumers * 8)
        *reinterpret_cast<uint4*>(xs + i) = *reinterpret_cast<const uint4*>(xg + i);
      named_bar_sync(1, kGsConsumers);
      xsrc = xs;
    } else {
      xsrc = reinterpret_cast<const __nv_bfloat16*>(p.x);
    }

    float acc[kGsRT][B];
#pragma unroll
    for (int r = 0; r < kGsRT; ++r)
#pragma unroll
      for (int b = 0; b < B; ++b) acc[r][b] = 0.f;

    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % n_stages;
      const uint32_t ph = (uint32_t)(t / n_stages) & 1u;
      const int rbi = t / n_kc;
      const int kci = t % n_kc;
      const int col = kci * kGsKC + warp * 256 + lane * 8;
      const bool col_ok = col < p.K;
      float xr[B][8];
#pragma unroll
      for (int b = 0; b < B; ++b) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (col_ok) v = *reinterpret_cast<const uint4*>(xsrc + (int64_t)b * p.K + col);
        xr[b][0] = bf16lo(v.x); xr[b][1] = bf16hi(v.x);
        xr[b][2] = bf16lo(v.y); xr[b][3] = bf16hi(v.y);
        xr[b][4] = bf16lo(v.z); xr[b][5] = bf16hi(v.z);
        xr[b][6] = bf16lo(v.w); xr[b][7] = bf16hi(v.w);
      }
      mbar_wait(&full_bar[s], ph, 0x600u + s);
      if (col_ok) {
        const uint8_t* tile = ring + (size_t)s * kGsStageBytes + warp * 512 + lane * 16;
#pragma unroll
        for (int r = 0; r < kGsRT; ++r) {
          const uint4 w = *reinterpret_cast<const uint4*>(tile + (size_t)r * (kGsKC * 2));
          const float w0 = bf16lo(w.x), w1 = bf16hi(w.x), w2 = bf16lo(w.y), w3 = bf16hi(w.y);
          const float w4 = bf16lo(w.z), w5 = bf16hi(w.z), w6 = bf16lo(w.w), w7 = bf16hi(w.w);
#pragma unroll
          for (int b = 0; b < B; ++b) {
            float a = acc[r][b];
            a = fmaf(w0, xr[b][0], a); a = fmaf(w1, xr[b][1], a);
            a = fmaf(w2, xr[b][2], a); a = fmaf(w3, xr[b][3], a);
            a = fmaf(w4, xr[b][4], a); a = fmaf(w5, xr[b][5], a);
            a = fmaf(w6, xr[b][6], a); a = fmaf(w7, xr[b][7], a);
            acc[r][b] = a;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);  // this warp is done reading the stage

      if (kci == n_kc - 1) {
        // rows [rb, rb+RT) are complete for this warp's columns: reduce and publish
        const int rb = row_begin + rbi * kGsRT;
        const int rows = min(kGsRT, row_end - rb);
        const int buf = rbi & 1;
        constexpr int RPR = (kGsRT * B <= 32) ? kGsRT : kGsRT / 2;  // rows per reduction round
#pragma unroll
        for (int r0 = 0; r0 < kGsRT; r0 += RPR) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
          for (int r = 0; r < RPR; ++r)
#pragma unroll
            for (int b = 0; b < B; ++b) v[r * B + b] = acc[r0 + r][b];
          const float tot = warp_reduce_32vals(v, lane);
          if (lane < RPR * B) s_part[buf][warp][r0 + lane / B][lane % B] = tot;
        }
#pragma unroll
        for (int r = 0; r < kGsRT; ++r)
#pragma unroll
          for (int b = 0; b < B; ++b) acc[r][b] = 0.f;
        named_bar_sync(1, kGsConsumers);  // all 8 warps' partials for this row block are visible
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
        // s_part is double-buffered by row block: the next block's writers cannot race these readers
        // (they pass another named barrier before buffer `buf` is written again)
      }
    }
  }
  pdl_launch_dependents();
}

}  // namespace advspec
