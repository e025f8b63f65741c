// attn.cuh — attention over the KV cache.
//   attn_prefill_kernel   causal flash attention for a prompt chunk (mma.sync
//                         m16n8k16 bf16, fp32 online softmax); ~7-24 % of the
//                         prefill flops at 4K-32K tokens.
//   attn_prefill_check    scalar restatement, verification only.
//   attn_decode_kernel    one new token per opponent: split-KV over the SHARED
//                         prefix (read once for all opponents and all query
//                         heads of a KV head) plus each opponent's own suffix.
//   attn_decode_combine   log-sum-exp merge of the split partials.
// KV layout: bf16 [n_kv_heads][stride][head_dim] per layer, K and V separate.
#pragma once

#include "common.cuh"

namespace advspec {

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct AttnPrefillParams {
  const __nv_bfloat16* q;  // [n_q][ldq], head h at column h*DH (already RoPE'd)
  int64_t ldq;
  const __nv_bfloat16* kc;  // [Hkv][kv_stride][DH]
  const __nv_bfloat16* vc;
  int64_t kv_stride;
  __nv_bfloat16* out;  // [n_q][H*DH]
  int n_q, q_pos0, H, Hkv;
  float scale;  // head_dim^-0.5
  int dh;       // head_dim in GLOBAL memory; the tensor-core kernel's tile width DH >= dh is zero-padded
};

// ---------------------------------------------------------------- prefill (TC)
// CTA = BM/16 warps x 16 query rows = BM queries of one head (BM = 128 halves the K/V re-reads
// from L2 at long prompts); keys in tiles of 64,
// double-buffered with cp.async; smem rows are 16-byte-chunk XOR swizzled so
// ldmatrix is conflict-free.
template <int DH, int BM>
__global__ void __launch_bounds__(BM * 2) attn_prefill_kernel(AttnPrefillParams p) {
  constexpr int BN = 64;
  constexpr int NT = BM * 2;  // one warp per 16 query rows
  constexpr int CPR = DH / 8;  // 16-byte chunks per row
  extern __shared__ __align__(128) uint8_t attn_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* sK = sQ + BM * DH;      // [2][BN][DH]
  __nv_bfloat16* sV = sK + 2 * BN * DH;  // [2][BN][DH]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int n_qtiles = (p.n_q + BM - 1) / BM;
  const int qt = n_qtiles - 1 - (int)blockIdx.x;  // heaviest (latest) tiles first
  const int q0 = qt * BM;
  const int h = blockIdx.y;
  const int hk = h / (p.H / p.Hkv);
  const int dhg = p.dh;  // <= DH; chunks past it are zero-filled (Phi-3: 96 inside a 128-wide tile)
  const __nv_bfloat16* kbase = p.kc + (int64_t)hk * p.kv_stride * dhg;
  const __nv_bfloat16* vbase = p.vc + (int64_t)hk * p.kv_stride * dhg;
  const int total_kv = p.q_pos0 + p.n_q;
  const int last_q = min(q0 + BM, p.n_q) - 1;
  const int n_kvt = (p.q_pos0 + last_q) / BN + 1;

  auto swz = [](int row, int chunk) { return row * DH + ((chunk ^ (row & 7)) << 3); };

  // Q tile
  for (int id = tid; id < BM * CPR; id += NT) {
    const int r = id / CPR, c = id % CPR;
    const bool ok = (q0 + r) < p.n_q && c * 8 < dhg;
    const __nv_bfloat16* src = p.q + (int64_t)(ok ? q0 + r : 0) * p.ldq + h * dhg + (ok ? c * 8 : 0);
    cp_async16(sQ + swz(r, c), src, ok);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * BN;
    for (int id = tid; id < BN * CPR; id += NT) {
      const int r = id / CPR, c = id % CPR;
      const bool ok = (k0 + r) < total_kv && c * 8 < dhg;
      const int64_t off = (int64_t)(ok ? k0 + r : 0) * dhg + (ok ? c * 8 : 0);
      cp_async16(sK + buf * BN * DH + swz(r, c), kbase + off, ok);
      cp_async16(sV + buf * BN * DH + swz(r, c), vbase + off, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  uint32_t qf[DH / 16][4];
  float o[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const float sl2 = p.scale * 1.4426950408889634f;
  const int qpos_lo = p.q_pos0 + q0 + warp * 16 + g;  // row g; row g+8 is +8

  for (int j = 0; j < n_kvt; ++j) {
    const int buf = j & 1;
    if (j + 1 < n_kvt) load_kv(j + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const int mi = lane >> 3;
        const int r = warp * 16 + (lane & 7) + (mi & 1) * 8;
        const int c = kk * 2 + (mi >> 1);
        ldmatrix_x4(qf[kk], sQ + swz(r, c));
      }
    }
    const __nv_bfloat16* bK = sK + buf * BN * DH;
    const __nv_bfloat16* bV = sV + buf * BN * DH;

    float s[BN / 8][4];
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
      for (int nb2 = 0; nb2 < BN / 16; ++nb2) {
        uint32_t r[4];
        const int mi = lane >> 3;
        const int row = nb2 * 16 + (mi >> 1) * 8 + (lane & 7);
        const int c = kk * 2 + (mi & 1);
        ldmatrix_x4(r, bK + swz(row, c));
        mma_bf16_16816(s[2 * nb2], qf[kk], r[0], r[1]);
        mma_bf16_16816(s[2 * nb2 + 1], qf[kk], r[2], r[3]);
      }
    }
    // causal mask (only tiles that reach past this CTA's earliest query)
    const int k0 = j * BN;
    if (k0 + BN - 1 > p.q_pos0 + q0) {
#pragma unroll
      for (int nb = 0; nb < BN / 8; ++nb) {
        const int kp = k0 + nb * 8 + 2 * t4;
        if (kp > qpos_lo) s[nb][0] = -INFINITY;
        if (kp + 1 > qpos_lo) s[nb][1] = -INFINITY;
        if (kp > qpos_lo + 8) s[nb][2] = -INFINITY;
        if (kp + 1 > qpos_lo + 8) s[nb][3] = -INFINITY;
      }
    }
    // online softmax, rows g (i=0) and g+8 (i=1)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < BN / 8; ++nb) mx = fmaxf(mx, fmaxf(s[nb][2 * i], s[nb][2 * i + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run[i], mx);
      const float m_off = (m_new == -INFINITY) ? 0.f : m_new * sl2;
      const float corr = (m_run[i] == -INFINITY) ? 0.f : ex2_approx(m_run[i] * sl2 - m_off);
      m_run[i] = m_new;
      float rs = 0.f;
#pragma unroll
      for (int nb = 0; nb < BN / 8; ++nb) {
        const float p0 = ex2_approx(fmaf(s[nb][2 * i], sl2, -m_off));
        const float p1 = ex2_approx(fmaf(s[nb][2 * i + 1], sl2, -m_off));
        s[nb][2 * i] = p0;
        s[nb][2 * i + 1] = p1;
        rs += p0 + p1;
      }
      l_run[i] = l_run[i] * corr + rs;
#pragma unroll
      for (int d = 0; d < DH / 8; ++d) {
        o[d][2 * i] *= corr;
        o[d][2 * i + 1] *= corr;
      }
    }
    // O += P V
#pragma unroll
    for (int kk2 = 0; kk2 < BN / 16; ++kk2) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk2][0], s[2 * kk2][1]);
      a[1] = pack_bf16(s[2 * kk2][2], s[2 * kk2][3]);
      a[2] = pack_bf16(s[2 * kk2 + 1][0], s[2 * kk2 + 1][1]);
      a[3] = pack_bf16(s[2 * kk2 + 1][2], s[2 * kk2 + 1][3]);
#pragma unroll
      for (int db2 = 0; db2 < DH / 16; ++db2) {
        uint32_t r[4];
        const int mi = lane >> 3;
        const int row = kk2 * 16 + (mi & 1) * 8 + (lane & 7);
        const int c = db2 * 2 + (mi >> 1);
        ldmatrix_x4_trans(r, bV + swz(row, c));
        mma_bf16_16816(o[2 * db2], a, r[0], r[1]);
        mma_bf16_16816(o[2 * db2 + 1], a, r[2], r[3]);
      }
    }
    __syncthreads();  // everyone is done with buf before the next iteration refills it
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float l = l_run[i];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = (l > 0.f) ? 1.0f / l : 0.f;
    const int qr = q0 + warp * 16 + g + i * 8;
    if (qr < p.n_q) {
      __nv_bfloat16* dst = p.out + (int64_t)qr * (p.H * dhg) + h * dhg;
#pragma unroll
      for (int d = 0; d < DH / 8; ++d) {
        if (d * 8 < dhg) {
          const uint32_t v = pack_bf16(o[d][2 * i] * inv, o[d][2 * i + 1] * inv);
          *reinterpret_cast<uint32_t*>(dst + d * 8 + 2 * t4) = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------ prefill (check)
// One warp per (query, head); fp32 throughout except the bf16 inputs.
__global__ void attn_prefill_check_kernel(AttnPrefillParams p, int DH) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= p.n_q * p.H) return;
  const int qi = wid / p.H, h = wid % p.H;
  const int hk = h / (p.H / p.Hkv);
  const int per = DH / 32;  // 2,3,4,8
  float qv[8], acc[8];
  for (int e = 0; e < per; ++e) {
    qv[e] = __bfloat162float(p.q[(int64_t)qi * p.ldq + h * DH + lane * per + e]);
    acc[e] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const int qpos = p.q_pos0 + qi;
  for (int k = 0; k <= qpos; ++k) {
    const __nv_bfloat16* kr = p.kc + ((int64_t)hk * p.kv_stride + k) * DH + lane * per;
    float d = 0.f;
    for (int e = 0; e < per; ++e) d = fmaf(qv[e], __bfloat162float(kr[e]), d);
    d = warp_sum(d) * p.scale;
    const float m_new = fmaxf(m, d);
    const float corr = expf(m - m_new);
    const float pw = expf(d - m_new);
    l = l * corr + pw;
    const __nv_bfloat16* vr = p.vc + ((int64_t)hk * p.kv_stride + k) * DH + lane * per;
    for (int e = 0; e < per; ++e) acc[e] = acc[e] * corr + pw * __bfloat162float(vr[e]);
    m = m_new;
  }
  for (int e = 0; e < per; ++e)
    p.out[(int64_t)qi * (p.H * DH) + h * DH + lane * per + e] = __float2bfloat16_rn(acc[e] / l);
}

// ------------------------------------------------------------------- decode
// Work item: up to 4 query rows that share one KV source — either a slice of
// the shared prefix (seq < 0) or the private suffix of opponent `seq`.
struct AttnItem {
  int kv_head;
  int seq;        // -1: shared prefix; >= 0: suffix of this opponent slot
  int tok_begin;  // prefix slice [tok_begin, tok_end); suffix: [0, suf_len[seq]+1)
  int tok_end;
  int n_rows;
  int row_b[4];     // batch index of each row (position in this decode call)
  int row_head[4];  // query head of each row
  int slot;         // partial slot written by this item
};

struct AttnDecodeParams {
  const AttnItem* items;
  const float* q;             // [b][H][DH] fp32, RoPE'd
  const __nv_bfloat16* pk;    // prefix K [Hkv][pstride][DH]
  const __nv_bfloat16* pv;
  int64_t pstride;
  const __nv_bfloat16* sk;    // suffix K [max_seqs][Hkv][sstride][DH] (this layer)
  const __nv_bfloat16* sv;
  int64_t sstride;
  const int* suf_len;         // [max_seqs] tokens already in each suffix (before this step)
  float* part_m;              // [b*H][n_slots]
  float* part_l;
  float* part_o;              // [b*H][n_slots][DH]
  int H, Hkv, n_slots;
  float scale;
};

// 128 threads = 16 groups of 8 lanes; a group owns one token per step (each
// lane DH/8 contiguous dims, 16-byte loads) and keeps its own online-softmax
// state for the item's <= 4 rows; groups are merged at the end.
template <int DH>
__global__ void __launch_bounds__(128) attn_decode_kernel(AttnDecodeParams p) {
  constexpr int DPL = DH / 8;  // dims per lane
  static_assert(DPL % 4 == 0, "head_dim must be a multiple of 32");
  ktrace_mark(TK_ATTN);
  pdl_wait();
  const AttnItem it = p.items[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = tid >> 3, l8 = tid & 7;

  const __nv_bfloat16 *kb, *vb;
  int tb, te;
  if (it.seq < 0) {
    kb = p.pk + (int64_t)it.kv_head * p.pstride * DH;
    vb = p.pv + (int64_t)it.kv_head * p.pstride * DH;
    tb = it.tok_begin;
    te = it.tok_end;
  } else {
    const int64_t base = ((int64_t)it.seq * p.Hkv + it.kv_head) * p.sstride * DH;
    kb = p.sk + base;
    vb = p.sv + base;
    tb = 0;
    te = p.suf_len[it.seq] + 1;
  }

  const float sl2 = p.scale * 1.4426950408889634f;
  float q[4][DPL], o[4][DPL], m[4], l[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
    const bool ok = r < it.n_rows;
    const float* qr = p.q + ((int64_t)(ok ? it.row_b[r] : 0) * p.H + (ok ? it.row_head[r] : 0)) * DH +
                      l8 * DPL;
#pragma unroll
    for (int e = 0; e < DPL; ++e) {
      q[r][e] = ok ? qr[e] * sl2 : 0.f;
      o[r][e] = 0.f;
    }
  }

  const uint32_t gmask = 0xFFu << (lane & 24);
  for (int t = tb + grp; t < te; t += 16) {
    float kv[DPL];
    const uint2* ks = reinterpret_cast<const uint2*>(kb + (int64_t)t * DH + l8 * DPL);
#pragma unroll
    for (int e = 0; e < DPL / 4; ++e) {
      const uint2 w = ks[e];
      kv[4 * e + 0] = bf16lo(w.x);
      kv[4 * e + 1] = bf16hi(w.x);
      kv[4 * e + 2] = bf16lo(w.y);
      kv[4 * e + 3] = bf16hi(w.y);
    }
    float sc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < DPL; ++e) d = fmaf(q[r][e], kv[e], d);
      // the 4 groups of a warp can run different trip counts: shuffle within the group only
      d += __shfl_xor_sync(gmask, d, 1);
      d += __shfl_xor_sync(gmask, d, 2);
      d += __shfl_xor_sync(gmask, d, 4);
      sc[r] = d;  // already in log2 units
    }
    const uint2* vs = reinterpret_cast<const uint2*>(vb + (int64_t)t * DH + l8 * DPL);
#pragma unroll
    for (int e = 0; e < DPL / 4; ++e) {
      const uint2 w = vs[e];
      kv[4 * e + 0] = bf16lo(w.x);
      kv[4 * e + 1] = bf16hi(w.x);
      kv[4 * e + 2] = bf16lo(w.y);
      kv[4 * e + 3] = bf16hi(w.y);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (sc[r] > m[r]) {  // rare after the first few tokens
        const float corr = exp2f(m[r] - sc[r]);  // m = -inf -> 0
        l[r] *= corr;
#pragma unroll
        for (int e = 0; e < DPL; ++e) o[r][e] *= corr;
        m[r] = sc[r];
      }
      const float pw = exp2f(sc[r] - m[r]);
      l[r] += pw;
#pragma unroll
      for (int e = 0; e < DPL; ++e) o[r][e] = fmaf(pw, kv[e], o[r][e]);
    }
  }

  // merge the 4 groups of a warp (lanes l8, l8+8, l8+16, l8+24 hold the same dims)
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m[r], off);
      const float l2 = __shfl_xor_sync(0xffffffffu, l[r], off);
      const float mn = fmaxf(m[r], m2);
      const float c1 = (m[r] == -INFINITY) ? 0.f : exp2f(m[r] - mn);
      const float c2 = (m2 == -INFINITY) ? 0.f : exp2f(m2 - mn);
      l[r] = l[r] * c1 + l2 * c2;
#pragma unroll
      for (int e = 0; e < DPL; ++e) {
        const float o2 = __shfl_xor_sync(0xffffffffu, o[r][e], off);
        o[r][e] = o[r][e] * c1 + o2 * c2;
      }
      m[r] = mn;
    }
  }
  // merge the 4 warps through shared memory
  __shared__ float s_m[4][4], s_l[4][4];
  __shared__ float s_o[4][4][DH];
  if (lane < 8) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (l8 == 0) {
        s_m[warp][r] = m[r];
        s_l[warp][r] = l[r];
      }
#pragma unroll
      for (int e = 0; e < DPL; ++e) s_o[warp][r][l8 * DPL + e] = o[r][e];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < it.n_rows * DH; idx += 128) {
    const int r = idx / DH, d = idx % DH;
    float mn = -INFINITY;
    for (int w = 0; w < 4; ++w) mn = fmaxf(mn, s_m[w][r]);
    float lt = 0.f, ot = 0.f;
    for (int w = 0; w < 4; ++w) {
      const float c = (s_m[w][r] == -INFINITY) ? 0.f : exp2f(s_m[w][r] - mn);
      lt += s_l[w][r] * c;
      ot += s_o[w][r][d] * c;
    }
    const int64_t ps = ((int64_t)it.row_b[r] * p.H + it.row_head[r]) * p.n_slots + it.slot;
    p.part_o[ps * DH + d] = ot;
    if (d == 0) {
      p.part_m[ps] = mn;
      p.part_l[ps] = lt;
    }
  }
  pdl_launch_dependents();
}

// out[b][h*DH + d] (bf16) = sum_s o_s * 2^(m_s - M) / sum_s l_s * 2^(m_s - M)
__global__ void attn_decode_combine_kernel(const float* __restrict__ part_m,
                                           const float* __restrict__ part_l,
                                           const float* __restrict__ part_o,
                                           __nv_bfloat16* __restrict__ out, int n_slots, int DH) {
  ktrace_mark(TK_COMBINE);
  pdl_wait();
  const int row = blockIdx.x;  // b*H + h
  float M = -INFINITY;
  for (int s = 0; s < n_slots; ++s) M = fmaxf(M, part_m[(int64_t)row * n_slots + s]);
  float L = 0.f;
  for (int s = 0; s < n_slots; ++s) {
    const float ms = part_m[(int64_t)row * n_slots + s];
    if (ms != -INFINITY) L += part_l[(int64_t)row * n_slots + s] * exp2f(ms - M);
  }
  const float inv = L > 0.f ? 1.0f / L : 0.f;
  for (int d = threadIdx.x; d < DH; d += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < n_slots; ++s) {
      const float ms = part_m[(int64_t)row * n_slots + s];
      if (ms != -INFINITY) acc += part_o[((int64_t)row * n_slots + s) * DH + d] * exp2f(ms - M);
    }
    out[(int64_t)row * DH + d] = __float2bfloat16_rn(acc * inv);
  }
  pdl_launch_dependents();
}

}  // namespace advspec
