// decode_kernels.cuh — the HBM-bound half of the path: one new token per
// opponent per step.  The dominant kernel is gemv_kernel: every weight matrix
// is streamed from HBM exactly once per step for ALL b opponents (b <= 8), so
// its algorithmic bytes are N*K*2 per launch whatever b is.
#pragma once

#include "common.cuh"
#include "gemm_tcgen05.cuh"  // epilogue enums, activations

namespace advspec {

// ---------------------------------------------------------------------------
// gemv_kernel: y[b][n] = epilogue( sum_k W[n][k] * x[b][k] )
//   W bf16 [N][K] row-major, 16-byte streaming loads (L1 no-allocate).
//   The CTA's 8 warps split K: in one "pass" warp w owns CH chunks of 256
//   elements, keeps its slice of x for all b opponents in REGISTERS, and
//   streams rows of W against it (4 rows in flight per warp = 8 x 16 B per
//   lane).  Partial sums are warp-reduced and accumulated per warp in shared
//   memory; rows are finished after the last pass.
//   in_mode 1 fuses RMSNorm: x is the fp32 residual stream, every CTA derives
//   1/rms itself (b*K fp32 from L2) and applies norm_w, rounding to bf16 exactly
//   like the prefill path does before its GEMM.
// ---------------------------------------------------------------------------
struct GemvParams {
  const __nv_bfloat16* W;
  const void* x;        // in_mode 0: bf16 [b][K]; in_mode 1: f32 [b][K]
  const float* norm_w;  // in_mode 1
  const float* bias;    // EPI_BF16 only, may be null
  void* y;
  int N, K;
  int in_mode, epilogue, act;
  float eps;
};

constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = 8;
constexpr int kGemvRB = 32;  // rows finished per shared-memory round
constexpr int kGemvRU = 4;   // rows in flight per warp

template <int B, int CH>
__global__ void __launch_bounds__(kGemvThreads, 2) gemv_kernel(GemvParams p) {
  constexpr int KP = kGemvWarps * CH * 256;  // K elements covered by one pass
  __shared__ float s_part[kGemvWarps][kGemvRB][B];
  __shared__ float s_red[kGemvWarps][B];
  __shared__ float s_inv[B];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  ktrace_mark(TK_GEMV);
  // contiguous row range of this CTA, in units of 2 rows (gated pairs stay together)
  const int pairs = (p.N + 1) / 2;
  const int row_begin = (int)(((int64_t)pairs * blockIdx.x) / gridDim.x) * 2;
  const int row_end = min(p.N, (int)(((int64_t)pairs * (blockIdx.x + 1)) / gridDim.x) * 2);
  const int n_pass = (p.K + KP - 1) / KP;

  // Weights do not depend on the previous kernel: start pulling the first rows
  // before waiting on it (the lines land in L2 while the predecessor drains).
  if (row_begin < row_end) {
    const int k = warp * CH * 256 + lane * 8;
    if (k < p.K) {
      const __nv_bfloat16* w0 = p.W + (int64_t)row_begin * p.K + k;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(w0));
    }
  }
  pdl_wait();

  if (p.in_mode == 1) {
    // 1/rms per opponent over the full row
    const float* xf = reinterpret_cast<const float*>(p.x);
    float ss[B];
#pragma unroll
    for (int b = 0; b < B; ++b) ss[b] = 0.f;
    for (int k = tid * 4; k < p.K; k += kGemvThreads * 4) {
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
      for (int w = 0; w < kGemvWarps; ++w) t += s_red[w][tid];
      s_inv[tid] = rsqrtf(t / (float)p.K + p.eps);
    }
    __syncthreads();
  }

  for (int rb = row_begin; rb < row_end; rb += kGemvRB) {
    const int rows_here = min(kGemvRB, row_end - rb);
    for (int i = lane; i < kGemvRB * B; i += 32) (&s_part[warp][0][0])[i] = 0.f;
    __syncwarp();

    for (int pass = 0; pass < n_pass; ++pass) {
      const int kw = pass * KP + warp * CH * 256;  // this warp's first column
      float xr[CH][B][8];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int k = kw + c * 256 + lane * 8;
        const bool ok = k < p.K;  // K % 8 == 0 is required by the host wrapper
#pragma unroll
        for (int b = 0; b < B; ++b) {
          if (!ok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) xr[c][b][e] = 0.f;
          } else if (p.in_mode == 0) {
            const uint4 v = *reinterpret_cast<const uint4*>(
                reinterpret_cast<const __nv_bfloat16*>(p.x) + (int64_t)b * p.K + k);
            xr[c][b][0] = bf16lo(v.x); xr[c][b][1] = bf16hi(v.x);
            xr[c][b][2] = bf16lo(v.y); xr[c][b][3] = bf16hi(v.y);
            xr[c][b][4] = bf16lo(v.z); xr[c][b][5] = bf16hi(v.z);
            xr[c][b][6] = bf16lo(v.w); xr[c][b][7] = bf16hi(v.w);
          } else {
            const float* xf = reinterpret_cast<const float*>(p.x) + (int64_t)b * p.K + k;
            const float4 a0 = *reinterpret_cast<const float4*>(xf);
            const float4 a1 = *reinterpret_cast<const float4*>(xf + 4);
            const float4 w0 = *reinterpret_cast<const float4*>(p.norm_w + k);
            const float4 w1 = *reinterpret_cast<const float4*>(p.norm_w + k + 4);
            const float inv = s_inv[b];
            xr[c][b][0] = round_bf16(a0.x * inv * w0.x);
            xr[c][b][1] = round_bf16(a0.y * inv * w0.y);
            xr[c][b][2] = round_bf16(a0.z * inv * w0.z);
            xr[c][b][3] = round_bf16(a0.w * inv * w0.w);
            xr[c][b][4] = round_bf16(a1.x * inv * w1.x);
            xr[c][b][5] = round_bf16(a1.y * inv * w1.y);
            xr[c][b][6] = round_bf16(a1.z * inv * w1.z);
            xr[c][b][7] = round_bf16(a1.w * inv * w1.w);
          }
        }
      }
      if (kw >= p.K) continue;  // warp-uniform: this warp has no columns in this pass

      for (int r0 = 0; r0 < rows_here; r0 += kGemvRU) {
        uint4 wv[kGemvRU][CH];
#pragma unroll
        for (int u = 0; u < kGemvRU; ++u) {
          const int r = min(rb + r0 + u, row_end - 1);  // clamp: tail rows are discarded below
          const __nv_bfloat16* wrow = p.W + (int64_t)r * p.K + kw + lane * 8;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const bool ok = (kw + c * 256 + lane * 8) < p.K;
            wv[u][c] = ok ? ldg_stream(wrow + c * 256) : make_uint4(0u, 0u, 0u, 0u);
          }
        }
        float acc[kGemvRU][B];
#pragma unroll
        for (int u = 0; u < kGemvRU; ++u) {
#pragma unroll
          for (int b = 0; b < B; ++b) acc[u][b] = 0.f;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const float w0 = bf16lo(wv[u][c].x), w1 = bf16hi(wv[u][c].x);
            const float w2 = bf16lo(wv[u][c].y), w3 = bf16hi(wv[u][c].y);
            const float w4 = bf16lo(wv[u][c].z), w5 = bf16hi(wv[u][c].z);
            const float w6 = bf16lo(wv[u][c].w), w7 = bf16hi(wv[u][c].w);
#pragma unroll
            for (int b = 0; b < B; ++b) {
              float a = acc[u][b];
              a = fmaf(w0, xr[c][b][0], a); a = fmaf(w1, xr[c][b][1], a);
              a = fmaf(w2, xr[c][b][2], a); a = fmaf(w3, xr[c][b][3], a);
              a = fmaf(w4, xr[c][b][4], a); a = fmaf(w5, xr[c][b][5], a);
              a = fmaf(w6, xr[c][b][6], a); a = fmaf(w7, xr[c][b][7], a);
              acc[u][b] = a;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kGemvRU; ++u) {
#pragma unroll
          for (int b = 0; b < B; ++b) {
            const float t = warp_sum(acc[u][b]);
            if (lane == 0 && r0 + u < rows_here) s_part[warp][r0 + u][b] += t;
          }
        }
      }
    }
    __syncthreads();

    // finish rows: fixed-order sum over the 8 warps, then the fused epilogue
    if (p.epilogue == EPI_GATED_BF16) {
      for (int i = tid; i < (rows_here / 2) * B; i += kGemvThreads) {
        const int pr = i / B, b = i % B;
        float gsum = 0.f, usum = 0.f;
#pragma unroll
        for (int w = 0; w < kGemvWarps; ++w) {
          gsum += s_part[w][2 * pr][b];
          usum += s_part[w][2 * pr + 1][b];
        }
        const int col = (rb >> 1) + pr;
        reinterpret_cast<__nv_bfloat16*>(p.y)[(int64_t)b * (p.N / 2) + col] =
            __float2bfloat16_rn(apply_act(gsum, p.act) * usum);
      }
    } else {
      for (int i = tid; i < rows_here * B; i += kGemvThreads) {
        const int r = i / B, b = i % B;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kGemvWarps; ++w) t += s_part[w][r][b];
        const int n = rb + r;
        if (p.epilogue == EPI_BF16) {
          if (p.bias) t += p.bias[n];
          reinterpret_cast<__nv_bfloat16*>(p.y)[(int64_t)b * p.N + n] = __float2bfloat16_rn(t);
        } else if (p.epilogue == EPI_RESADD_F32) {
          reinterpret_cast<float*>(p.y)[(int64_t)b * p.N + n] += t;
        } else {
          reinterpret_cast<float*>(p.y)[(int64_t)b * p.N + n] = t;
        }
      }
    }
    __syncthreads();
  }
  pdl_launch_dependents();
}

// ---------------------------------------------------------------------------
// prefill-side elementwise kernels
// ---------------------------------------------------------------------------
// x f32 [n][d] <- embed[tok] * scale
__global__ void embed_kernel(const int* __restrict__ tokens, const __nv_bfloat16* __restrict__ embed,
                             float* __restrict__ x, int d, float scale) {
  const int row = blockIdx.x;
  const __nv_bfloat16* e = embed + (int64_t)tokens[row] * d;
  // HF multiplies in the activation dtype; the oracle runs fp32 activations on
  // bf16-valued weights, so scale is applied in fp32 here as well.
  for (int k = threadIdx.x; k < d; k += blockDim.x)
    x[(int64_t)row * d + k] = __bfloat162float(e[k]) * scale;
}

// xn bf16 [n][d] = round_bf16( x * rsqrt(mean(x^2) + eps) * w )   (one CTA per row)
__global__ void rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                               __nv_bfloat16* __restrict__ xn, int d, float eps) {
  __shared__ float s_red[32];
  const int row = blockIdx.x;
  const float* xr = x + (int64_t)row * d;
  float ss = 0.f;
  for (int k = threadIdx.x; k < d; k += blockDim.x) ss += xr[k] * xr[k];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) s_red[0] = rsqrtf(t / (float)d + eps);
  }
  __syncthreads();
  const float inv = s_red[0];
  for (int k = threadIdx.x; k < d; k += blockDim.x)
    xn[(int64_t)row * d + k] = __float2bfloat16_rn(xr[k] * inv * w[k]);
}

// Prefill RMSNorm, bandwidth version (d % 4 == 0, d <= 8192): one CTA per row keeps the whole row in
// registers as 16-byte vectors — the row is read ONCE (fp32, 16 bytes per load), normalised and written as
// packed bf16 (8 bytes per store).  The scalar kernel above re-read the row and ran at ~2 TB/s on the
// 5,068 x 4,096 residual stream (60 us x 64 launches per prefill).
template <int VPT>
__global__ void __launch_bounds__(256) rmsnorm_vec_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          __nv_bfloat16* __restrict__ xn, int d, float eps) {
  __shared__ float s_red[8];
  __shared__ float s_inv;
  const int row = blockIdx.x, tid = threadIdx.x, n4 = d >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * d);
  float4 v[VPT];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int k = tid + i * 256;
    v[i] = k < n4 ? xr[k] : make_float4(0.f, 0.f, 0.f, 0.f);
    ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) s_red[tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_red[i];
    s_inv = rsqrtf(t / (float)d + eps);
  }
  __syncthreads();
  const float inv = s_inv;
  const float4* w4 = reinterpret_cast<const float4*>(w);
  uint2* out = reinterpret_cast<uint2*>(xn + (int64_t)row * d);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int k = tid + i * 256;
    if (k < n4) {
      const float4 g = w4[k];
      out[k] = make_uint2(pack_bf16(v[i].x * inv * g.x, v[i].y * inv * g.y),
                          pack_bf16(v[i].z * inv * g.z, v[i].w * inv * g.w));
    }
  }
}

inline cudaError_t launch_rmsnorm(const float* x, const float* w, __nv_bfloat16* xn, int rows, int d, float eps,
                                  cudaStream_t st) {
  const int n4 = d / 4;
  if (d % 4 == 0 && n4 <= 8 * 256) {
    if (n4 <= 256) rmsnorm_vec_kernel<1><<<rows, 256, 0, st>>>(x, w, xn, d, eps);
    else if (n4 <= 512) rmsnorm_vec_kernel<2><<<rows, 256, 0, st>>>(x, w, xn, d, eps);
    else if (n4 <= 1024) rmsnorm_vec_kernel<4><<<rows, 256, 0, st>>>(x, w, xn, d, eps);
    else rmsnorm_vec_kernel<8><<<rows, 256, 0, st>>>(x, w, xn, d, eps);
  } else {
    rmsnorm_kernel<<<rows, 256, 0, st>>>(x, w, xn, d, eps);
  }
  return cudaGetLastError();
}

// cos/sin table [max_pos][DH/2]: angle = (float)pos * inv_freq[i] in fp32, as
// HF's rotary embedding does (modeling_llama.py rotary forward), then cosf/sinf.
__global__ void rope_table_kernel(const float* __restrict__ inv_freq, float* __restrict__ cs,
                                  float* __restrict__ sn, int max_pos, int half) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)max_pos * half) return;
  const int pos = (int)(idx / half), i = (int)(idx % half);
  const float ang = (float)pos * inv_freq[i];
  cs[idx] = cosf(ang);
  sn[idx] = sinf(ang);
}

// Prompt chunk: rotate q in place, rotate k into the K cache, copy v into the V
// cache.  qkv bf16 [n][ldq] = [q heads | k heads | v heads]; token i is at
// position pos0+i.  rotate_half convention (modeling_llama.py:138-146).
__global__ void rope_prefill_kernel(__nv_bfloat16* __restrict__ qkv, int64_t ldq,
                                    __nv_bfloat16* __restrict__ kc, __nv_bfloat16* __restrict__ vc,
                                    int64_t kv_stride, const float* __restrict__ cs,
                                    const float* __restrict__ sn, int pos0, int H, int Hkv,
                                    int DH) {
  const int row = blockIdx.x;
  const int pos = pos0 + row;
  const int half = DH / 2;
  __nv_bfloat16* base = qkv + (int64_t)row * ldq;
  const int n_rot = (H + Hkv) * half;
  for (int i = threadIdx.x; i < n_rot; i += blockDim.x) {
    const int head = i / half, j = i % half;
    const float c = cs[(int64_t)pos * half + j], s = sn[(int64_t)pos * half + j];
    __nv_bfloat16* hp = base + head * DH;
    const float a = __bfloat162float(hp[j]), b = __bfloat162float(hp[j + half]);
    const __nv_bfloat16 r0 = __float2bfloat16_rn(a * c - b * s);
    const __nv_bfloat16 r1 = __float2bfloat16_rn(b * c + a * s);
    if (head < H) {
      hp[j] = r0;
      hp[j + half] = r1;
    } else {
      __nv_bfloat16* dst = kc + ((int64_t)(head - H) * kv_stride + pos) * DH;
      dst[j] = r0;
      dst[j + half] = r1;
    }
  }
  for (int i = threadIdx.x; i < Hkv * DH; i += blockDim.x) {
    const int head = i / DH, j = i % DH;
    vc[((int64_t)head * kv_stride + pos) * DH + j] = base[(H + Hkv) * DH + i];
  }
}

// Decode step: per opponent b, rotate q -> fp32 q_out [b][H][DH]; rotate k and
// copy v into that opponent's suffix cache at index suf_len[slot].
__global__ void rope_decode_kernel(const __nv_bfloat16* __restrict__ qkv, float* __restrict__ q_out,
                                   __nv_bfloat16* __restrict__ sk, __nv_bfloat16* __restrict__ sv,
                                   int64_t sstride, const int* __restrict__ slots,
                                   const int* __restrict__ suf_len, int prefix_len,
                                   const float* __restrict__ cs, const float* __restrict__ sn, int H,
                                   int Hkv, int DH) {
  ktrace_mark(TK_ROPE);
  pdl_wait();
  const int b = blockIdx.x;
  const int slot = slots[b];
  const int t = suf_len[slot];
  const int pos = prefix_len + t;
  const int half = DH / 2;
  const int QKV = (H + 2 * Hkv) * DH;
  const __nv_bfloat16* base = qkv + (int64_t)b * QKV;
  for (int i = threadIdx.x; i < (H + Hkv) * half; i += blockDim.x) {
    const int head = i / half, j = i % half;
    const float c = cs[(int64_t)pos * half + j], s = sn[(int64_t)pos * half + j];
    const float a = __bfloat162float(base[head * DH + j]);
    const float bb = __bfloat162float(base[head * DH + j + half]);
    // round like the prefill path (bf16 q and k after rotation)
    const float r0 = round_bf16(a * c - bb * s), r1 = round_bf16(bb * c + a * s);
    if (head < H) {
      q_out[((int64_t)b * H + head) * DH + j] = r0;
      q_out[((int64_t)b * H + head) * DH + j + half] = r1;
    } else {
      __nv_bfloat16* dst = sk + (((int64_t)slot * Hkv + (head - H)) * sstride + t) * DH;
      dst[j] = __float2bfloat16_rn(r0);
      dst[j + half] = __float2bfloat16_rn(r1);
    }
  }
  for (int i = threadIdx.x; i < Hkv * DH; i += blockDim.x) {
    const int head = i / DH, j = i % DH;
    sv[(((int64_t)slot * Hkv + head) * sstride + t) * DH + j] = base[(H + Hkv) * DH + i];
  }
  pdl_launch_dependents();
}

// ---------------------------------------------------------------------------
// sampling: token = argmax_v( logit_v / T + Gumbel(u_v) ), u from uniform01()
// (T == 0: plain argmax, lowest index wins ties).  One CTA per opponent.  The
// same kernel appends the token to the output, advances the opponent's state
// and writes the next step's input embedding into the residual stream.
// ---------------------------------------------------------------------------
struct SampleParams {
  const float* logits;       // [b][V], or [1][V] when broadcast_logits
  int broadcast_logits;      // first token: every opponent samples the prefill logits
  int V;                     // vocabulary entries in `logits` (a tensor-parallel rank holds V_full / tp_size)
  int v_off;                 // token id of logits[0] (tp_rank * V)
  int n_ranks, rank_stride;  // merge: part_best/part_idx of rank r start r * rank_stride words further on
  float temperature;
  const int* slots;          // [b]
  const uint64_t* seeds;     // [max_seqs]
  int* suf_len;              // [max_seqs]  (advanced when advance != 0)
  int* n_out;                // [max_seqs]  tokens emitted so far
  int* done;                 // [max_seqs]
  int* out_tokens;           // [max_seqs][out_stride]
  int out_stride;
  int eos_id;
  int advance;               // 1: the KV of the previous input token is now in the cache
  int* pos_b;                // [b] absolute position of each opponent's next input token
  int prefix_len;
  const float* part_best;    // [b][kSampleChunks] from sample_partial_kernel, or null: scan here
  const int* part_idx;
  const int* forced;         // teacher forcing: [b] tokens to use instead of sampling
  int* cur_tok;              // [max_seqs] next input token
  const __nv_bfloat16* embed;
  float* x;                  // [b][d] residual stream for the next step
  int d;
  float embed_scale;
};

// Stage 1 of sampling: grid (chunks, b); every CTA scans a slice of the vocabulary and leaves its best
// (perturbed score, index) in part_best / part_idx [b][chunks].  sample_kernel then only merges those.
constexpr int kSampleChunks = 64;
__global__ void __launch_bounds__(256) sample_partial_kernel(SampleParams p, float* __restrict__ part_best,
                                                             int* __restrict__ part_idx) {
  ktrace_mark(TK_SAMPLE_SCAN);
  if (!g_ktrace_on) pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_best[8];
  __shared__ int s_idx[8];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int slot = p.slots[b];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t step = (uint32_t)p.n_out[slot];
  const float* lg = p.logits + (p.broadcast_logits ? 0 : (int64_t)b * p.V);
  const uint64_t seed = p.seeds[slot];
  const float invT = p.temperature > 0.f ? 1.0f / p.temperature : 1.0f;
  const int per = (p.V + kSampleChunks - 1) / kSampleChunks;
  const int v0 = chunk * per, v1 = min(p.V, v0 + per);
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  for (int v = v0 + tid; v < v1; v += 256) {
    const int tok = p.v_off + v;  // global token id: noise and tie rule do not depend on the sharding
    float sc = lg[v] * invT;
    if (p.temperature > 0.f) sc -= logf(-logf(uniform01(seed, step, (uint32_t)tok)));
    if (sc > best || (sc == best && tok < bidx)) { best = sc; bidx = tok; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
  }
  if (lane == 0) { s_best[warp] = best; s_idx[warp] = bidx; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_best[w] > best || (s_best[w] == best && s_idx[w] < bidx)) { best = s_best[w]; bidx = s_idx[w]; }
    part_best[b * kSampleChunks + chunk] = best;
    part_idx[b * kSampleChunks + chunk] = bidx;
  }
  pdl_launch_dependents();
}

__global__ void __launch_bounds__(1024) sample_kernel(SampleParams p) {
  ktrace_mark(TK_SAMPLE);
  if (!g_ktrace_on) pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_best[32];
  __shared__ int s_idx[32];
  __shared__ int s_tok;
  const int b = blockIdx.x;
  const int slot = p.slots[b];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    int sl = p.suf_len[slot];
    if (p.advance) p.suf_len[slot] = ++sl;  // previous token's KV is in place
    p.pos_b[b] = p.prefix_len + sl;         // absolute position of the token about to be fed
  }
  const uint32_t step = (uint32_t)p.n_out[slot];
  const bool was_done = p.done[slot] != 0;

  if (p.forced == nullptr && p.part_best != nullptr) {
    // merge the per-chunk winners of sample_partial_kernel (same tie rule: lowest index)
    if (warp == 0) {
      float best = -INFINITY;
      int bidx = 0x7fffffff;
      for (int c = lane; c < kSampleChunks * p.n_ranks; c += 32) {
        const int r = c / kSampleChunks, cc = c % kSampleChunks;
        const float ob = p.part_best[(int64_t)r * p.rank_stride + b * kSampleChunks + cc];
        const int oi = p.part_idx[(int64_t)r * p.rank_stride + b * kSampleChunks + cc];
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
      }
      if (lane == 0) s_tok = bidx;
    }
  } else if (p.forced == nullptr) {
    const float* lg = p.logits + (p.broadcast_logits ? 0 : (int64_t)b * p.V);
    const uint64_t seed = p.seeds[slot];
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    const float invT = p.temperature > 0.f ? 1.0f / p.temperature : 1.0f;
    for (int v = tid; v < p.V; v += blockDim.x) {
      float sc = lg[v] * invT;
      if (p.temperature > 0.f) {
        const float u = uniform01(seed, step, (uint32_t)v);
        sc -= logf(-logf(u));
      }
      if (sc > best || (sc == best && v < bidx)) { best = sc; bidx = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) { s_best[warp] = best; s_idx[warp] = bidx; }
    __syncthreads();
    if (warp == 0) {
      best = (lane < (int)(blockDim.x >> 5)) ? s_best[lane] : -INFINITY;
      bidx = (lane < (int)(blockDim.x >> 5)) ? s_idx[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
      }
      if (lane == 0) s_tok = bidx;
    }
  } else if (tid == 0) {
    s_tok = p.forced[b];
  }
  __syncthreads();
  const int tok = s_tok;
  if (tid == 0) {
    p.cur_tok[slot] = tok;
    if (!was_done && p.out_tokens != nullptr) {
      p.out_tokens[(int64_t)slot * p.out_stride + step] = tok;
      p.n_out[slot] = (int)step + 1;
      if (tok == p.eos_id) p.done[slot] = 1;
    }
  }
  const __nv_bfloat16* e = p.embed + (int64_t)tok * p.d;
  for (int k = tid; k < p.d; k += blockDim.x)
    p.x[(int64_t)b * p.d + k] = __bfloat162float(e[k]) * p.embed_scale;
  pdl_launch_dependents();
}

// ---------------------------------------------------------------------------
// seeded synthetic weights: N(0, std) via Box-Muller on the sampler's hash
// ---------------------------------------------------------------------------
__global__ void init_normal_bf16_kernel(__nv_bfloat16* w, int64_t n, uint64_t seed, float std) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const uint64_t h = mix64(seed ^ mix64((uint64_t)i));
    const float u1 = ((float)(uint32_t)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(uint32_t)((h >> 16) & 0xFFFFFFu) + 0.5f) * (1.0f / 16777216.0f);
    const float z = sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
    w[i] = __float2bfloat16_rn(z * std);
  }
}
__global__ void fill_f32_kernel(float* w, int64_t n, float v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) w[i] = v;
}

}  // namespace advspec
