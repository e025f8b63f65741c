// gemv_rmma.cuh — decode GEMV, fifth design (the one the engine runs): weights go
// HBM -> registers -> tensor cores, with 32 warps per SM keeping >= 128 KB in flight.
//
// Measured on this B200 (tools/bw_probe.cu, tools/gemv_bench.cu):
//   * a read-only stream needs ~128 KB in flight per SM (32+ warps) to reach 6.8 TB/s;
//     8 warps/SM cap at 3.8 TB/s, 16 warps/SM at 6.2 TB/s;
//   * 1-D bulk async copies cost ~60 ns each per SM, so a ring fed with 2 KB row segments
//     tops out at 4.5 TB/s; per-lane cp.async is slower still;
//   * with b = 3 opponents a CUDA-core FMA consumer needs ~33 issue slots per 16 B of W.
// So: plain 16-byte LDGs (L1 no-allocate) issued by 1024 threads per SM, no shared-memory
// staging of W at all, and mma.sync.m16n8k16 as the consumer with the opponents as N = 8.
//
// Fragment trick: an MMA sums over k, so any permutation of k applied to A and B alike is
// free.  Lane (g = lane/4, t = lane%4) loads 16 bytes = W[row g][c + 8t .. c + 8t + 7] and
// W[row g+8][same]; the first four elements of each load are its (k = 2t, 2t+1, 2t+8, 2t+9)
// slots of one MMA, the last four of a second MMA — registers of the load ARE the A
// fragments.  The matching B fragment is x[opponent g][c + 8t .. c + 8t + 7], one 16-byte
// shared-memory load feeding both MMAs.  One warp step = 16 rows x 32 columns = 1 KB of W
// for 2 LDG + 1 LDS + 2 MMA.
//
// CTA = 16 warps, 2 CTAs per SM, each CTA owns a contiguous row range; a row block is 16
// rows; its K is cut into 128-column chunks dealt round-robin to the 16 warps; per-warp
// 16 x b partial sums go to shared memory and are reduced once per group of 8 row blocks.
// The first chunk's loads are issued before griddepcontrol.wait.
#pragma once

#include "../../adversarial-spec_b200/csrc/attn.cuh"  // mma wrapper
#include "../../adversarial-spec_b200/csrc/common.cuh"
#include "../../adversarial-spec_b200/csrc/decode_kernels.cuh"

namespace advspec {

constexpr int kGrWarps = 16;
constexpr int kGrThreads = kGrWarps * 32;
constexpr int kGrRT = 16;       // rows per block (MMA M)
constexpr int kGrSteps = 4;     // warp steps (32 columns each) per chunk
constexpr int kGrChunk = 32 * kGrSteps;  // 128 columns
constexpr int kGrRBG = 8;       // row blocks reduced together

template <int B>
__global__ void __launch_bounds__(kGrThreads, 2) gemv_rmma_kernel(GemvParams p) {
  extern __shared__ __align__(128) uint8_t gr_smem[];
  // bf16 x [B][K] with a 64-byte skew per opponent (conflict-free 16-byte fragment loads)
  const int xpitch = p.K * 2 + 64;
  uint8_t* xs = gr_smem;
  float* s_part = reinterpret_cast<float*>(gr_smem + (((size_t)B * xpitch + 127) / 128) * 128);  // [16][RBG][16][B]
  __shared__ float s_red[kGrWarps][B];
  __shared__ float s_inv[B];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  ktrace_mark(TK_GEMV);
  const int pairs = (p.N + 1) / 2;
  const int row_begin = (int)(((int64_t)pairs * blockIdx.x) / gridDim.x) * 2;
  const int row_end = min(p.N, (int)(((int64_t)pairs * (blockIdx.x + 1)) / gridDim.x) * 2);
  const int n_rb = (row_end - row_begin + kGrRT - 1) / kGrRT;
  const int n_ch = (p.K + kGrChunk - 1) / kGrChunk;  // chunks per row block
  const int my_ch = (n_ch - warp + kGrWarps - 1) / kGrWarps;  // chunks of a row block dealt to this warp
  const int n_it = n_rb * my_ch;                     // this warp's iterations

  uint4 wq[kGrSteps][2];  // [step][row g | row g+8]
  // (row block, chunk) -> 16-byte loads of rows g and g+8; rows past the range are clamped
  // (their results are discarded), columns past K load zeros
  auto issue = [&](int rbi, int ci) {
    const int rb = row_begin + rbi * kGrRT;
    const int col = (warp + ci * kGrWarps) * kGrChunk + 8 * t4;
    const __nv_bfloat16* p0 = p.W + (int64_t)min(rb + g, row_end - 1) * p.K + col;
    const __nv_bfloat16* p1 = p.W + (int64_t)min(rb + g + 8, row_end - 1) * p.K + col;
#pragma unroll
    for (int s = 0; s < kGrSteps; ++s) {
      const bool ok = col + 32 * s < p.K;  // K % 32 == 0 is required by the host wrapper
      wq[s][0] = ok ? ldg_stream(p0 + 32 * s) : make_uint4(0u, 0u, 0u, 0u);
      wq[s][1] = ok ? ldg_stream(p1 + 32 * s) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  if (n_it > 0) issue(0, 0);  // weights do not depend on the previous kernel
  pdl_wait();

  // ---- stage x (bf16) in shared memory; in_mode 1 fuses the RMSNorm of the fp32 residual stream
  if (p.in_mode == 1) {
    const float* xf = reinterpret_cast<const float*>(p.x);
    float ss[B];
#pragma unroll
    for (int b = 0; b < B; ++b) ss[b] = 0.f;
    for (int k = tid * 4; k < p.K; k += kGrThreads * 4) {
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
      for (int w = 0; w < kGrWarps; ++w) t += s_red[w][tid];
      s_inv[tid] = rsqrtf(t / (float)p.K + p.eps);
    }
    __syncthreads();
    for (int k = tid * 4; k < p.K; k += kGrThreads * 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(p.norm_w + k);
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const float4 v = *reinterpret_cast<const float4*>(xf + (int64_t)b * p.K + k);
        const float inv = s_inv[b];
        uint2 o;
        o.x = pack_bf16(v.x * inv * w4.x, v.y * inv * w4.y);
        o.y = pack_bf16(v.z * inv * w4.z, v.w * inv * w4.w);
        *reinterpret_cast<uint2*>(xs + (size_t)b * xpitch + (size_t)k * 2) = o;
      }
    }
  } else {
    const __nv_bfloat16* xg = reinterpret_cast<const __nv_bfloat16*>(p.x);
    for (int i = tid * 8; i < B * p.K; i += kGrThreads * 8) {
      const int b = i / p.K, k = i % p.K;
      *reinterpret_cast<uint4*>(xs + (size_t)b * xpitch + (size_t)k * 2) = *reinterpret_cast<const uint4*>(xg + i);
    }
  }
  __syncthreads();

  const bool opp_ok = g < B;
  const uint8_t* xrow = xs + (size_t)(opp_ok ? g : 0) * xpitch + 16 * t4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};

  for (int rb0 = 0; rb0 < n_rb; rb0 += kGrRBG) {
    const int rbs = min(kGrRBG, n_rb - rb0);
    for (int rbl = 0; rbl < rbs; ++rbl) {
      for (int ci = 0; ci < my_ch; ++ci) {
        const int col0 = (warp + ci * kGrWarps) * kGrChunk;
#pragma unroll
        for (int s = 0; s < kGrSteps; ++s) {
          if (col0 + 32 * s < p.K) {  // warp-uniform
            const uint4 xb = opp_ok ? *reinterpret_cast<const uint4*>(xrow + (size_t)(col0 + 32 * s) * 2)
                                    : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t a1[4] = {wq[s][0].x, wq[s][1].x, wq[s][0].y, wq[s][1].y};
            const uint32_t a2[4] = {wq[s][0].z, wq[s][1].z, wq[s][0].w, wq[s][1].w};
            mma_bf16_16816(acc, a1, xb.x, xb.y);
            mma_bf16_16816(acc, a2, xb.z, xb.w);
          }
        }
        // refill the same registers with this warp's next chunk; the other 31 warps of the SM cover
        // the latency (1024 threads x 8 x 16 B = 128 KB in flight per SM)
        if (ci + 1 < my_ch) issue(rb0 + rbl, ci + 1);
        else if (rb0 + rbl + 1 < n_rb) issue(rb0 + rbl + 1, 0);
      }
      if (my_ch > 0) {
        // this warp's share of the row block is complete: park it (opponents < B only)
        float* dst = s_part + ((size_t)(warp * kGrRBG + rbl) * kGrRT) * B;
        if (2 * t4 < B) dst[g * B + 2 * t4] = acc[0];
        if (2 * t4 + 1 < B) dst[g * B + 2 * t4 + 1] = acc[1];
        if (2 * t4 < B) dst[(g + 8) * B + 2 * t4] = acc[2];
        if (2 * t4 + 1 < B) dst[(g + 8) * B + 2 * t4 + 1] = acc[3];
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
      }
    }
    if (my_ch == 0) {  // warps beyond the chunk count contribute zeros
      for (int i = lane; i < rbs * kGrRT * B; i += 32) s_part[(size_t)warp * kGrRBG * kGrRT * B + i] = 0.f;
    }
    __syncthreads();
    // ---- reduce the 16 warps' partials and apply the fused epilogue
    const int rb_first = row_begin + rb0 * kGrRT;
    const int rows = min(rbs * kGrRT, row_end - rb_first);
    if (p.epilogue == EPI_GATED_BF16) {
      for (int i = tid; i < (rows / 2) * B; i += kGrThreads) {
        const int pr = i / B, b = i % B;
        float gs = 0.f, us = 0.f;
#pragma unroll
        for (int w = 0; w < kGrWarps; ++w) {
          const float* src = s_part + (size_t)w * kGrRBG * kGrRT * B;
          gs += src[(2 * pr) * B + b];
          us += src[(2 * pr + 1) * B + b];
        }
        reinterpret_cast<__nv_bfloat16*>(p.y)[(int64_t)b * (p.N / 2) + (rb_first >> 1) + pr] =
            __float2bfloat16_rn(apply_act(gs, p.act) * us);
      }
    } else {
      for (int i = tid; i < rows * B; i += kGrThreads) {
        const int r = i / B, b = i % B;
        float tsum = 0.f;
#pragma unroll
        for (int w = 0; w < kGrWarps; ++w) tsum += s_part[(size_t)w * kGrRBG * kGrRT * B + r * B + b];
        const int n = rb_first + r;
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
    if (rb0 + kGrRBG < n_rb) __syncthreads();  // s_part is rewritten by the next group
  }
  pdl_launch_dependents();
}

}  // namespace advspec
