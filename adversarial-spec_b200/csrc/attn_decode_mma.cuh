// attn_decode_mma.cuh — decode attention, second design: one kernel per layer does
//   RoPE of the new q/k  ->  append k/v to the opponent's private suffix KV
//   ->  split-KV flash-decoding on tensor cores (mma.sync m16n8k16, bf16)
//   ->  per-split partial (m, l, o); attn_decode_combine2_kernel (below) merges the splits.
// A work item is (KV head, opponent group, KV source): the shared prefix is cut
// into splits that are read ONCE for all opponents and all query heads of the KV
// head (up to 16 query rows = one MMA M tile); each opponent's suffix is its own
// item.  Prefix K/V tiles of 64 keys arrive by TMA (tensor maps over the KV region) into a
// ring of up to six stages, suffix tiles by cp.async into the same layout; two warp groups
// take alternate tiles, each warp 16 keys of a tile with a private online-softmax state.
// Head dims 64 / 128 / 256 natively, 96 in the 128-wide tile with zero padding.
// Replaces rope_decode_kernel + attn_decode_kernel + attn_decode_combine_kernel.
#pragma once

#include "attn.cuh"
#include "common.cuh"

namespace advspec {

// Work decomposition is pure arithmetic on blockIdx (no table to fetch): per KV head, opponents are
// grouped so that a group's query rows (opponents x G heads) fill one 16-row MMA tile; the prefix is
// cut into n_splits slices shared by the whole group; each opponent's suffix is its own item.
struct AttnDecode2Params {
  const __nv_bfloat16* qkv;  // [b][(H+2Hkv)*DH] raw projections of the new token (bias applied)
  const float* rope_cos;     // [max_pos][DH/2]
  const float* rope_sin;
  const __nv_bfloat16* pk;   // prefix K [Hkv][pstride][DH]
  const __nv_bfloat16* pv;
  int64_t pstride;
  __nv_bfloat16* sk;         // suffix K [max_seqs][Hkv][sstride][DH] (this layer)
  __nv_bfloat16* sv;
  int64_t sstride;
  CUtensorMap map_k, map_v;  // prefix K / V of this layer: 2-D (Hkv*pstride tokens) x DH, box 64 tokens x 64 dims,
                             // 128-byte swizzle.  By value in the (grid-constant) parameter block: the TMA unit
                             // then reads the descriptors from the constant bank instead of global memory
  const int* pos_b;          // [b] absolute position of each opponent's new token (device state)
  int slots[8];              // batch index -> opponent slot (fixed for the decode call)
  int prefix_len;
  float* part_m;             // [b*H][n_slots]
  float* part_l;
  float* part_o;             // [b*H][n_slots][DH]
  int b, H, Hkv, G;
  int opg, n_og, n_splits, n_slots;  // opponents per group, groups per KV head, prefix splits, n_splits+1
  float scale;
  int dh;                    // head_dim in GLOBAL memory (<= DH): Phi-3's 96 runs in the 128-wide tile, zero-padded
};

template <int DH, int NST>
__global__ void __launch_bounds__(256, 1) attn_decode_mma_kernel(const __grid_constant__ AttnDecode2Params p) {
  constexpr int BN = 64;
  constexpr int CPR = DH / 8;
  constexpr int TILE = BN * DH;  // elements of one K (or V) tile
  constexpr int NH = DH / 64;    // 128-byte-wide halves of a row (TMA swizzle atoms are 128 B wide)
  constexpr int HALF = BN * 64;  // elements of one 64-token x 64-dim half tile (8 KB)
  extern __shared__ uint8_t ad_smem_raw[];
  // ring first, 1024-byte aligned (128-byte TMA swizzle atoms); then the query tile; then barriers
  uint8_t* ad_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ad_smem_raw) + 1023) &
                                                ~static_cast<uintptr_t>(1023));
  __nv_bfloat16* sKV = reinterpret_cast<__nv_bfloat16*>(ad_smem);  // [NST][K halves | V halves][64][64]
  __nv_bfloat16* sQ = sKV + (size_t)NST * 2 * TILE;                // [16][DH]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sQ + 16 * DH);  // [NST]

  ktrace_mark(TK_ATTN);
  phase_mark(0);
  if (!g_ktrace_on) pdl_launch_dependents();  // dependents only pre-stage weights before their own wait
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = warp >> 2, w4 = warp & 3;  // two warp groups take alternate key tiles
  const int g = lane >> 2, t4 = lane & 3;
  const int dhg = p.dh;  // global head_dim; tile columns >= dhg are zero (TMA out-of-bounds fill / zfill)
  const int QKV = (p.H + 2 * p.Hkv) * dhg;
  const int half = dhg / 2;
  auto swz = [](int row, int chunk) { return row * DH + ((chunk ^ (row & 7)) << 3); };  // query tile
  // K/V tiles: element offset of 16-byte chunk `chunk` (8 dims) of token row `row` in the layout TMA
  // writes with CU_TENSOR_MAP_SWIZZLE_128B: half tiles of [64 tokens][64 dims], chunk index XOR (row % 8)
  auto kvz = [](int row, int chunk) { return (chunk >> 3) * HALF + row * 64 + (((chunk & 7) ^ (row & 7)) << 3); };

  // ---- which item am I
  // Block order: every prefix item first (group-major), then the per-opponent suffix items.  When the grid
  // exceeds one wave (many KV heads: Phi-3, Gemma) the long prefix CTAs then all start at once and the short
  // suffix CTAs backfill behind them, instead of a second wave that again contains full-length prefix CTAs.
  const int n_prefix = p.Hkv * p.n_og * p.n_splits;
  int grp, j;
  if ((int)blockIdx.x < n_prefix) {
    grp = blockIdx.x / p.n_splits;
    j = blockIdx.x % p.n_splits;
  } else {
    const int r = blockIdx.x - n_prefix;
    grp = r / p.opg;
    j = p.n_splits + r % p.opg;
  }
  const int hk = grp / p.n_og, og = grp % p.n_og;
  const int o0 = og * p.opg;
  const int n_opp = min(p.opg, p.b - o0);
  const bool is_prefix = j < p.n_splits;
  if (!is_prefix && (j - p.n_splits) >= n_opp) return;  // padding item of a short last group
  const int row_off = is_prefix ? 0 : (j - p.n_splits) * p.G;  // rows inside the group's 16-row tile
  const int n_rows = is_prefix ? n_opp * p.G : p.G;
  const int slot_out = is_prefix ? j : p.n_splits;

  // The lane-0 thread of warp s owns ring stage s: it initialises the stage's barrier and (prefix items) issues
  // the stage's first TMA at once — no CTA-wide barrier between kernel entry and the first load.  Everybody
  // else meets the initialised barriers at the __syncthreads() below, before the first wait.
  if ((tid & 31) == 0 && (tid >> 5) < NST) {
    if (tid == 0) {
      tma_prefetch_desc(&p.map_k);
      tma_prefetch_desc(&p.map_v);
    }
    mbar_init(&full_bar[tid >> 5], 1);
    fence_mbar_init();
  }

  const __nv_bfloat16 *kb, *vb;
  int tb, te;
  if (is_prefix) {
    kb = p.pk + (int64_t)hk * p.pstride * dhg;
    vb = p.pv + (int64_t)hk * p.pstride * dhg;
    tb = (int)((int64_t)p.prefix_len * j / p.n_splits);
    te = (int)((int64_t)p.prefix_len * (j + 1) / p.n_splits);
  } else {
    pdl_wait();  // needs this step's projections and positions
    // append the new token's k (rotated) and v to this opponent's suffix, then attend over it
    const int bi = o0 + (j - p.n_splits);
    const int pos = p.pos_b[bi];
    const int t = pos - p.prefix_len;
    const int64_t base = ((int64_t)p.slots[bi] * p.Hkv + hk) * p.sstride * dhg;
    __nv_bfloat16* kdst = p.sk + base + (int64_t)t * dhg;
    __nv_bfloat16* vdst = p.sv + base + (int64_t)t * dhg;
    const __nv_bfloat16* ksrc = p.qkv + (int64_t)bi * QKV + (p.H + hk) * dhg;
    const __nv_bfloat16* vsrc = p.qkv + (int64_t)bi * QKV + (p.H + p.Hkv + hk) * dhg;
    for (int jj = tid; jj < half; jj += 256) {
      const float c = p.rope_cos[(int64_t)pos * half + jj], s = p.rope_sin[(int64_t)pos * half + jj];
      const float a = __bfloat162float(ksrc[jj]), bb = __bfloat162float(ksrc[jj + half]);
      kdst[jj] = __float2bfloat16_rn(a * c - bb * s);
      kdst[jj + half] = __float2bfloat16_rn(bb * c + a * s);
    }
    for (int jj = tid; jj < dhg; jj += 256) vdst[jj] = vsrc[jj];
    __threadfence_block();
    kb = p.sk + base;
    vb = p.sv + base;
    tb = 0;
    te = t + 1;
  }

  const int n_tiles = (te - tb + BN - 1) / BN;
  // Prefix tiles: one elected thread issues 2*NH TMA box copies per tile (K and V halves) that complete
  // on the stage's mbarrier.  Suffix tiles (a handful of tokens the CTA itself just extended): per-lane
  // cp.async into the same swizzled layout.
  auto load_tile = [&](int tile, int issuer) {
    const int st = tile % NST;
    const int k0 = tb + tile * BN;
    __nv_bfloat16* dK = sKV + (size_t)st * 2 * TILE;
    __nv_bfloat16* dV = dK + TILE;
    if (is_prefix) {
      if (tid == issuer) {
        mbar_arrive_expect_tx(&full_bar[st], 2u * TILE * 2u);
        const int row0 = hk * (int)p.pstride + k0;
#pragma unroll
        for (int h2 = 0; h2 < NH; ++h2) {
          tma_load_2d(dK + h2 * HALF, &p.map_k, &full_bar[st], h2 * 64, row0);
          tma_load_2d(dV + h2 * HALF, &p.map_v, &full_bar[st], h2 * 64, row0);
        }
      }
    } else {
      for (int id = tid; id < BN * CPR; id += 256) {
        const int r = id / CPR, c = id % CPR;
        const bool ok = (k0 + r) < te && c * 8 < dhg;
        const int64_t off = (int64_t)(ok ? k0 + r : tb) * dhg + (ok ? c * 8 : 0);
        cp_async16(dK + kvz(r, c), kb + off, ok);
        cp_async16(dV + kvz(r, c), vb + off, ok);
      }
    }
  };
  // Put the whole slice in flight at once when it fits the ring (prefix KV is constant during decode,
  // so this happens before the dependency wait)
  if (is_prefix) {
    for (int s = 0; s < NST; ++s) {
      if (s < n_tiles) load_tile(s, s * 32);  // the stage's own thread issues: the TMA ops of different stages
      cp_async_commit();                       // enter the queue side by side, not one after another
    }
    phase_mark(1);
    pdl_wait();
    phase_mark(2);
  }
  __syncthreads();  // barrier initialisation (and the suffix item's appended row) visible to the whole CTA

  // ---- stage the rotated, bf16-rounded query rows (rows past n_rows are zero): one thread handles
  // 8 consecutive rotation pairs of one row with 16-byte loads and stores
  for (int i = tid; i < 16 * (half / 8); i += 256) {
    const int r = i / (half / 8), jc = i % (half / 8);  // row, chunk of 8 pairs
    uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = lo;
    if (r < n_rows) {
      const int gr = row_off + r;
      const int bi = o0 + gr / p.G, head = hk * p.G + gr % p.G;
      const int pos = p.pos_b[bi];
      const __nv_bfloat16* qs = p.qkv + (int64_t)bi * QKV + head * dhg + jc * 8;
      const uint4 a4 = *reinterpret_cast<const uint4*>(qs);
      const uint4 b4 = *reinterpret_cast<const uint4*>(qs + half);
      const float4 c0 = *reinterpret_cast<const float4*>(p.rope_cos + (int64_t)pos * half + jc * 8);
      const float4 c1 = *reinterpret_cast<const float4*>(p.rope_cos + (int64_t)pos * half + jc * 8 + 4);
      const float4 s0 = *reinterpret_cast<const float4*>(p.rope_sin + (int64_t)pos * half + jc * 8);
      const float4 s1 = *reinterpret_cast<const float4*>(p.rope_sin + (int64_t)pos * half + jc * 8 + 4);
      const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
      uint32_t lo_w[4], hi_w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a0 = bf16lo(aw[e]), a1 = bf16hi(aw[e]), b0 = bf16lo(bw[e]), b1 = bf16hi(bw[e]);
        lo_w[e] = pack_bf16(a0 * cs[2 * e] - b0 * sn[2 * e], a1 * cs[2 * e + 1] - b1 * sn[2 * e + 1]);
        hi_w[e] = pack_bf16(b0 * cs[2 * e] + a0 * sn[2 * e], b1 * cs[2 * e + 1] + a1 * sn[2 * e + 1]);
      }
      lo = make_uint4(lo_w[0], lo_w[1], lo_w[2], lo_w[3]);
      hi = make_uint4(hi_w[0], hi_w[1], hi_w[2], hi_w[3]);
    }
    *reinterpret_cast<uint4*>(sQ + swz(r, jc)) = lo;
    *reinterpret_cast<uint4*>(sQ + swz(r, jc + half / 8)) = hi;
  }
  for (int i = tid; i < 16 * (CPR - dhg / 8); i += 256) {  // padding columns of the query tile
    const int r = i / (CPR - dhg / 8), c = dhg / 8 + i % (CPR - dhg / 8);
    *reinterpret_cast<uint4*>(sQ + swz(r, c)) = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();  // sQ complete; the appended k/v row is visible to this CTA's loads
  if (!is_prefix) {
    for (int s = 0; s < NST; ++s) {
      if (s < n_tiles) load_tile(s, 0);
      cp_async_commit();
    }
  }
  phase_mark(3);

  uint32_t qf[DH / 16][4];
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) {
    const int mi = lane >> 3;
    ldmatrix_x4(qf[kk], sQ + swz((lane & 7) + (mi & 1) * 8, kk * 2 + (mi >> 1)));
  }
  float o[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sl2 = p.scale * 1.4426950408889634f;

  for (int jt = 0; jt < n_tiles; ++jt) {
    // tiles 0..NST-1 were issued up front; tile jt >= NST was issued at iteration jt - NST
    if (is_prefix) {
      mbar_wait(&full_bar[jt % NST], (uint32_t)(jt / NST) & 1u, 0x700u + (jt % NST));
    } else {
      cp_async_wait<NST - 1>();
      __syncthreads();
    }
    phase_mark(6 + (jt < 9 ? jt : 9));  // tile jt's data has landed (diagnostic stamps 6..15)
    if ((jt & 1) == wg) {
      const __nv_bfloat16* bK = sKV + (size_t)(jt % NST) * 2 * TILE;
      const __nv_bfloat16* bV = bK + TILE;
      const int kw0 = tb + jt * BN + w4 * 16;  // first key of this warp's 16-key slice
      if (kw0 < te) {
        float s[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          uint32_t r[4];
          const int mi = lane >> 3;
          ldmatrix_x4(r, bK + kvz(w4 * 16 + (mi >> 1) * 8 + (lane & 7), kk * 2 + (mi & 1)));
          mma_bf16_16816(s[0], qf[kk], r[0], r[1]);
          mma_bf16_16816(s[1], qf[kk], r[2], r[3]);
        }
        if (kw0 + 16 > te) {
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            const int kp = kw0 + nb * 8 + 2 * t4;
            if (kp >= te) s[nb][0] = s[nb][2] = -INFINITY;
            if (kp + 1 >= te) s[nb][1] = s[nb][3] = -INFINITY;
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float mx = fmaxf(fmaxf(s[0][2 * i], s[0][2 * i + 1]), fmaxf(s[1][2 * i], s[1][2 * i + 1]));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          const float m_new = fmaxf(m_run[i], mx);
          const float m_off = (m_new == -INFINITY) ? 0.f : m_new * sl2;
          const float corr = (m_run[i] == -INFINITY) ? 0.f : exp2f(m_run[i] * sl2 - m_off);
          m_run[i] = m_new;
          float rs = 0.f;
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            const float p0 = exp2f(s[nb][2 * i] * sl2 - m_off);
            const float p1 = exp2f(s[nb][2 * i + 1] * sl2 - m_off);
            s[nb][2 * i] = p0;
            s[nb][2 * i + 1] = p1;
            rs += p0 + p1;
          }
          l_run[i] = l_run[i] * corr + rs;
          if (corr != 1.0f) {
#pragma unroll
            for (int d = 0; d < DH / 8; ++d) {
              o[d][2 * i] *= corr;
              o[d][2 * i + 1] *= corr;
            }
          }
        }
        uint32_t a[4];
        a[0] = pack_bf16(s[0][0], s[0][1]);
        a[1] = pack_bf16(s[0][2], s[0][3]);
        a[2] = pack_bf16(s[1][0], s[1][1]);
        a[3] = pack_bf16(s[1][2], s[1][3]);
#pragma unroll
        for (int db2 = 0; db2 < DH / 16; ++db2) {
          uint32_t r[4];
          const int mi = lane >> 3;
          ldmatrix_x4_trans(r, bV + kvz(w4 * 16 + (mi & 1) * 8 + (lane & 7), db2 * 2 + (mi >> 1)));
          mma_bf16_16816(o[2 * db2], a, r[0], r[1]);
          mma_bf16_16816(o[2 * db2 + 1], a, r[2], r[3]);
        }
      }
    }
    if (jt + NST < n_tiles) {
      __syncthreads();  // stage jt % NST is free again
      load_tile(jt + NST, 0);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();
  __syncthreads();  // every warp is done with the ring before it is reused for the merge
  phase_mark(4);

  // ---- merge the 8 warps (disjoint key subsets) through shared memory
  constexpr int OP = DH + 8;                    // padded row pitch (floats): 2-way instead of 8-way conflicts
  float* s_m = reinterpret_cast<float*>(sKV);  // [8][16]
  float* s_l = s_m + 128;                       // [8][16]
  float* s_o = s_l + 128;                       // [8][16][OP]
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float l = l_run[i];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const int r = g + 8 * i;
    if (t4 == 0) {
      s_m[warp * 16 + r] = (m_run[i] == -INFINITY) ? -INFINITY : m_run[i] * sl2;
      s_l[warp * 16 + r] = l;
    }
#pragma unroll
    for (int d = 0; d < DH / 8; ++d) {
      *reinterpret_cast<float2*>(&s_o[(warp * 16 + r) * OP + d * 8 + 2 * t4]) = make_float2(o[d][2 * i], o[d][2 * i + 1]);
    }
  }
  __syncthreads();
  // weights 2^(m_w - M) of the 8 warps' states, once per row (threads 0..127: warp w, row r)
  float* s_c = s_o + 8 * 16 * OP;  // [8][16]
  float* s_ML = s_c + 128;         // [16][2]: M, L of the merged row
  if (tid < 16) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) M = fmaxf(M, s_m[w * 16 + tid]);
    float L = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float mw = s_m[w * 16 + tid];
      const float c = (mw == -INFINITY) ? 0.f : exp2f(mw - M);
      s_c[w * 16 + tid] = c;
      L += s_l[w * 16 + tid] * c;
    }
    s_ML[2 * tid] = M;
    s_ML[2 * tid + 1] = L;
  }
  __syncthreads();
  for (int idx = tid; idx < n_rows * dhg; idx += 256) {
    const int r = idx / dhg, d = idx % dhg;
    float O = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) O += s_o[(w * 16 + r) * OP + d] * s_c[w * 16 + r];
    const int gr = row_off + r;
    const int bi = o0 + gr / p.G, head = hk * p.G + gr % p.G;
    const int64_t ps = ((int64_t)bi * p.H + head) * p.n_slots + slot_out;
    p.part_o[ps * dhg + d] = O;
    if (d == 0) {
      p.part_m[ps] = s_ML[2 * r];
      p.part_l[ps] = s_ML[2 * r + 1];
    }
  }
  phase_mark(5);
  pdl_launch_dependents();
}

// Merge of the split partials of one (opponent, query head) row per CTA: weights
// 2^(m_s - M) / L are formed once in shared memory, then each thread owns one head
// dimension and keeps 8 independent loads in flight over the slots.
__global__ void __launch_bounds__(128) attn_decode_combine2_kernel(const float* __restrict__ part_m,
                                                                   const float* __restrict__ part_l,
                                                                   const float* __restrict__ part_o,
                                                                   __nv_bfloat16* __restrict__ out,
                                                                   int n_slots, int DH) {
  __shared__ float w[320];
  __shared__ float s_M;
  ktrace_mark(TK_COMBINE);
  if (!g_ktrace_on) pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x, tid = threadIdx.x;
  const int S = n_slots;
  // the partial outputs do not depend on the weights: put the first 16 slots' loads of this thread's
  // dimension in flight together with the m/l loads (one L2 round trip instead of two)
  constexpr int PRE = 16;
  float pre[PRE];
  const bool own = tid < DH;
  const float* po = part_o + (int64_t)row * S * DH + tid;
#pragma unroll
  for (int j = 0; j < PRE; ++j) pre[j] = (own && j < S) ? po[(int64_t)j * DH] : 0.f;
  float lv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int s = tid + i * 128;
    if (s < 320) w[s] = (s < S) ? part_m[(int64_t)row * S + s] : -INFINITY;
    lv[i] = (s < S) ? part_l[(int64_t)row * S + s] : 0.f;
  }
  __syncthreads();
  if (tid < 32) {
    float M = -INFINITY;
    for (int s = tid; s < S; s += 32) M = fmaxf(M, w[s]);
    M = warp_max(M);
    if (tid == 0) s_M = M;
  }
  __syncthreads();
  const float M = s_M;
  float part = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int s = tid + i * 128;
    if (s < S) {
      const float ws = (w[s] == -INFINITY) ? 0.f : exp2f(w[s] - M);
      part += lv[i] * ws;
      w[s] = ws;
    }
  }
  __shared__ float s_sum[4];
  part = warp_sum(part);
  if ((tid & 31) == 0) s_sum[tid >> 5] = part;
  __syncthreads();
  const float L = s_sum[0] + s_sum[1] + s_sum[2] + s_sum[3];  // same fixed order in every thread
  const float inv = L > 0.f ? 1.0f / L : 0.f;
  for (int d = tid; d < DH; d += 128) {
    const float* pd = part_o + (int64_t)row * S * DH + d;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    int s = 0;
    if (d == tid) {
#pragma unroll
      for (int j = 0; j < PRE; ++j)
        if (j < S) a[j & 7] = fmaf(pre[j], w[j], a[j & 7]);
      s = min(S, PRE);
    }
    for (; s + 8 <= S; s += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = pd[(int64_t)(s + j) * DH];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaf(v[j], w[s + j], a[j]);
    }
    for (; s < S; ++s) a[0] = fmaf(pd[(int64_t)s * DH], w[s], a[0]);
    const float tot = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    out[(int64_t)row * DH + d] = __float2bfloat16_rn(tot * inv);
  }
  pdl_launch_dependents();
}

}  // namespace advspec
