// gemv_cpasync.cuh — decode GEMV, fourth design: every warp streams its own slice of W with
// cp.async (16 bytes per lane per instruction) into a PRIVATE shared-memory ring and feeds it to
// mma.sync (opponents = the N=8 dimension), so
//   * loads are issued by all 256 threads (the 1-D bulk/TMA path tops out at one copy per ~60 ns
//     per SM, i.e. 4.5 TB/s with 2 KB row segments — measured, tools/gemv_bench.cu),
//   * nothing but cp.async.wait_group + __syncwarp orders producer and consumer (same warp),
//   * up to (NST-1) x 8 KB per warp = 128-190 KB per SM are in flight, issued BEFORE
//     griddepcontrol.wait because weights never depend on the previous kernel.
// A CTA tile is 16 rows x 2048 columns; warp w owns columns [w*256, w*256+256) of it.  Per row
// block the 8 warps' 16 x 8 partial outputs are summed in shared memory and the fused epilogue
// (bias / residual add / gated activation / fp32 store) writes the rows.
#pragma once

#include "../../adversarial-spec_b200/csrc/attn.cuh"  // cp_async16 / ldmatrix / mma wrappers
#include "../../adversarial-spec_b200/csrc/common.cuh"
#include "../../adversarial-spec_b200/csrc/decode_kernels.cuh"

namespace advspec {

constexpr int kGcRT = 16;                          // rows per tile (MMA M)
constexpr int kGcWC = 256;                         // columns per warp per tile
constexpr int kGcKC = 8 * kGcWC;                   // columns per CTA tile
constexpr int kGcRowPitch = kGcWC * 2 + 16;        // 528 B: staggers the 16 rows across banks
constexpr int kGcWarpStage = kGcRT * kGcRowPitch;  // 8,448 B per warp per stage
constexpr int kGcThreads = 256;

template <int B, int NST>
__global__ void __launch_bounds__(kGcThreads, 1) gemv_cpasync_kernel(GemvParams p, int x_in_smem) {
  extern __shared__ __align__(128) uint8_t gc_smem[];
  uint8_t* xs_raw = gc_smem + (size_t)8 * NST * kGcWarpStage;  // bf16 [B][K], pitch K*2+16
  const int xpitch = p.K * 2 + 16;
  __shared__ float s_part[2][8][kGcRT][8];
  __shared__ float s_red[8][B];
  __shared__ float s_inv[B];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  ktrace_mark(TK_GEMV);
  const int pairs = (p.N + 1) / 2;
  const int row_begin = (int)(((int64_t)pairs * blockIdx.x) / gridDim.x) * 2;
  const int row_end = min(p.N, (int)(((int64_t)pairs * (blockIdx.x + 1)) / gridDim.x) * 2);
  const int n_rb = (row_end - row_begin + kGcRT - 1) / kGcRT;
  const int n_kc = (p.K + kGcKC - 1) / kGcKC;
  const int n_tiles = n_rb * n_kc;
  uint8_t* wring = gc_smem + (size_t)warp * NST * kGcWarpStage;

  // lane -> (row parity, 16-byte chunk) of the warp's 16 x 256 slice: 2 rows per instruction
  const int l_row = lane >> 4, l_chunk = lane & 15;
  auto issue_tile = [&](int t) {
    const int rb = row_begin + (t / n_kc) * kGcRT;
    const int col = (t % n_kc) * kGcKC + warp * kGcWC + l_chunk * 8;
    uint8_t* dst = wring + (size_t)(t % NST) * kGcWarpStage + l_chunk * 16;
    const bool col_ok = col < p.K;
#pragma unroll
    for (int r2 = 0; r2 < kGcRT / 2; ++r2) {
      const int r = r2 * 2 + l_row;
      const bool ok = col_ok && (rb + r) < row_end;
      const __nv_bfloat16* src = p.W + (int64_t)(ok ? rb + r : row_begin) * p.K + (ok ? col : 0);
      cp_async16(dst + (size_t)r * kGcRowPitch, src, ok);
    }
  };
#pragma unroll
  for (int s = 0; s < NST - 1; ++s) {
    if (s < n_tiles) issue_tile(s);
    cp_async_commit();
  }
  pdl_wait();

  const uint8_t* xbase;
  int xstride;
  if (p.in_mode == 1) {
    const float* xf = reinterpret_cast<const float*>(p.x);
    float ss[B];
#pragma unroll
    for (int b = 0; b < B; ++b) ss[b] = 0.f;
    for (int k = tid * 4; k < p.K; k += kGcThreads * 4) {
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
    __syncthreads();
    if (tid < B) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += s_red[w][tid];
      s_inv[tid] = rsqrtf(t / (float)p.K + p.eps);
    }
    __syncthreads();
    for (int k = tid * 4; k < p.K; k += kGcThreads * 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(p.norm_w + k);
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const float4 v = *reinterpret_cast<const float4*>(xf + (int64_t)b * p.K + k);
        const float inv = s_inv[b];
        uint2 o;
        o.x = pack_bf16(v.x * inv * w4.x, v.y * inv * w4.y);
        o.y = pack_bf16(v.z * inv * w4.z, v.w * inv * w4.w);
        *reinterpret_cast<uint2*>(xs_raw + (size_t)b * xpitch + (size_t)k * 2) = o;
      }
    }
    __syncthreads();
    xbase = xs_raw;
    xstride = xpitch;
  } else if (x_in_smem) {
    const __nv_bfloat16* xg = reinterpret_cast<const __nv_bfloat16*>(p.x);
    for (int i = tid * 8; i < B * p.K; i += kGcThreads * 8) {
      const int b = i / p.K, k = i % p.K;
      *reinterpret_cast<uint4*>(xs_raw + (size_t)b * xpitch + (size_t)k * 2) =
          *reinterpret_cast<const uint4*>(xg + i);
    }
    __syncthreads();
    xbase = xs_raw;
    xstride = xpitch;
  } else {
    xbase = reinterpret_cast<const uint8_t*>(p.x);
    xstride = p.K * 2;
  }

  const int n_opp = lane >> 2, t4 = lane & 3;
  const bool opp_ok = n_opp < B;
  const uint8_t* xrow = xbase + (size_t)(opp_ok ? n_opp : 0) * xstride;
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int a_col = (lane >> 4) * 8;

  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t = 0; t < n_tiles; ++t) {
    if (t + NST - 1 < n_tiles) issue_tile(t + NST - 1);
    cp_async_commit();
    const int rbi = t / n_kc, kci = t % n_kc;
    const int col0 = kci * kGcKC + warp * kGcWC;
    cp_async_wait<NST - 1>();
    __syncwarp();  // every lane's copies of tile t have landed
    const uint8_t* tile = wring + (size_t)(t % NST) * kGcWarpStage + (size_t)a_row * kGcRowPitch + a_col * 2;
#pragma unroll
    for (int ks = 0; ks < kGcWC / 16; ++ks) {
      const int kcol = col0 + ks * 16;
      if (kcol < p.K) {  // warp-uniform (K % 16 == 0)
        const int k = kcol + 2 * t4;
        const uint32_t b0 = opp_ok ? *reinterpret_cast<const uint32_t*>(xrow + (size_t)k * 2) : 0u;
        const uint32_t b1 = opp_ok ? *reinterpret_cast<const uint32_t*>(xrow + (size_t)(k + 8) * 2) : 0u;
        uint32_t a[4];
        ldmatrix_x4(a, tile + ks * 32);
        mma_bf16_16816(acc, a, b0, b1);
      }
    }
    __syncwarp();  // all lanes are done with this stage before a later issue_tile overwrites it

    if (kci == n_kc - 1) {
      const int rb = row_begin + rbi * kGcRT;
      const int rows = min(kGcRT, row_end - rb);
      const int buf = rbi & 1;
      const int g = lane >> 2;
      *reinterpret_cast<float2*>(&s_part[buf][warp][g][2 * t4]) = make_float2(acc[0], acc[1]);
      *reinterpret_cast<float2*>(&s_part[buf][warp][g + 8][2 * t4]) = make_float2(acc[2], acc[3]);
      acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
      __syncthreads();
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
  cp_async_wait<0>();
  pdl_launch_dependents();
}

}  // namespace advspec
