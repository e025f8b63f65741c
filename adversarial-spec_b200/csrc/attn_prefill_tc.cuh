// attn_prefill_tc.cuh — causal flash attention for the prompt on the 5th-gen tensor cores.
//
// One CTA = 128 query rows of one head; keys in tiles of 128.  Per tile
//   S = Q K^T   tcgen05.mma 128x128x128, Q and K K-major from TMA-swizzled shared memory, S in TMEM
//   softmax     8 warps, TWO THREADS PER QUERY ROW (64 keys each): tcgen05.ld the scores, mask, exp2 against
//               a lazily-updated row maximum (rescale O only when the maximum grew by > 2^8), write P
//               (bf16) back to shared memory in the 128B-swizzled K-major layout.  The exponentials were
//               the limiter (16,384 MUFU.EX2 per tile = the tile's MMA time): every other 8-key chunk now
//               takes a degree-3 polynomial on the FMA pipe instead (relative error 1e-4, P is rounded to
//               bf16 = 4e-3 anyway), and P is DOUBLE-BUFFERED so the softmax of tile j+1 runs under the
//               P V of tile j instead of waiting for it
//   O += P V    tcgen05.mma 128x128x128, A = P (K-major), B = V straight from its [key][dim] tile as an
//               MN-major operand; O accumulates in TMEM across all tiles
// Warp roles: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-9 softmax/epilogue.
// S is double-buffered in TMEM so QK^T of tile j+1 runs under the softmax of tile j.
// TMEM columns: S0 [0,128) S1 [128,256) O [256,384).  head_dim 128, and 64 / 96 in the same 128-wide tiles:
// Q K^T issues only head_dim/16 k-steps (whatever the TMA box holds past the head is never multiplied), the K/V
// tensor maps are head_dim wide so the box columns past it arrive as zeros, and O's columns past head_dim are
// not written (Phi-3's 32 x 96 heads).  head_dim 256 (Gemma) uses attn.cuh.
// Shared memory: Q, P x2, (K, V) x2 stages = 7 x 32 KB + barriers + the row-maximum exchange = the whole
// 227 KB of the SM.
#pragma once

#include "attn.cuh"
#include "common.cuh"

namespace advspec {

constexpr int kAtBM = 128, kAtBN = 128, kAtDH = 128;
constexpr int kAtStages = 2;
constexpr int kAtHalf = 128 * 64 * 2;               // bytes of a [128 rows][64 elems] half tile (16 KB)
constexpr int kAtTile = 2 * kAtHalf;                // 32 KB: Q, K, V or P tile
constexpr int kAtBarBytes = 256;
constexpr int kAtMxBytes = 2 * 2 * 128 * 4;         // [tile parity][column half][row] row-maximum exchange
// Q, P x2, K/V stages, barriers, exchange, alignment slack (the base is 256-byte aligned: at most 768 bytes)
constexpr int kAtSmem = kAtTile * (3 + 2 * kAtStages) + kAtBarBytes + kAtMxBytes + 768;
static_assert(kAtSmem <= 232448, "the prompt-attention CTA needs the whole 227 KB");
constexpr int kAtThreads = 320;  // TMA warp, MMA warp, 8 softmax warps (2 threads per query row)

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major operand, 128-byte swizzle: 64 MN elements contiguous (128 B); consecutive k rows 128 B apart;
// groups of 8 k rows SBO = 1024 B apart; the next 64 MN elements LBO bytes away.
__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// 2^x on the FMA pipe: x = n + f with n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 polynomial
// (max relative error 1.0e-4), 2^n by adding n to the exponent field.  x is clamped at -126 (a masked
// score of -inf becomes 1e-38, nothing in a sum of O(1) terms).
__device__ __forceinline__ float exp2_poly3(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;  // 1.5 * 2^23: the low mantissa bits now hold round(x)
  const float f = x - (t - 12582912.0f);
  const float p = fmaf(f, fmaf(f, fmaf(f, 0.05500892922282219f, 0.24221095442771912f), 0.6932829022407532f), 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ int g_attn_poly = 1;  // ADVSPEC_ATTN_POLY=0: all exponentials on the MUFU (A/B)

struct AttnPrefillTcParams {
  __nv_bfloat16* out;  // [n_q][H*128]
  int n_q, q_pos0, H, Hkv;
  int kv_rows_per_head;  // kv_stride: row of (kv head hk, token t) in the K/V tensor maps = hk*kv_rows_per_head + t
  float scale;
  int dh;  // head_dim: 64, 96 or 128
};

// tmQ: [n_q rows][ldq cols] bf16 (head h at column h*dh), box 64 cols x 128 rows, 128B swizzle.
// tmK/tmV: [Hkv*kv_stride rows][dh cols], box 64 cols x 128 rows, 128B swizzle.
__global__ void __launch_bounds__(kAtThreads, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, AttnPrefillTcParams p) {
  extern __shared__ __align__(256) uint8_t at_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sP = smem + kAtTile;                         // [2]: tile parity
  uint8_t* sK = smem + 3 * kAtTile;                     // [stages]
  uint8_t* sV = smem + (3 + kAtStages) * kAtTile;       // [stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (3 + 2 * kAtStages) * kAtTile);
  uint64_t* q_full = bars;                 // 1
  // K and V of a stage have their own barriers: the K slot is free as soon as Q K^T of its tile has retired
  // (long before P V), so the load of K(j+2) — and with it S(j+2) — no longer waits behind softmax and P V of tile j
  uint64_t* k_full = bars + 1;             // [stages]
  uint64_t* v_full = k_full + kAtStages;   // [stages]
  uint64_t* k_empty = v_full + kAtStages;  // [stages]
  uint64_t* v_empty = k_empty + kAtStages; // [stages]
  uint64_t* s_full = v_empty + kAtStages;  // [2]
  uint64_t* s_empty = s_full + 2;          // [2]
  uint64_t* p_full = s_empty + 2;          // [2]: P of tile j is in buffer j & 1
  uint64_t* pv_done = p_full + 2;          // [2]: P V of tile j retired (frees P buffer j & 1 and, in order, O)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  // [tile parity][column half][row]: row-maximum exchange between the two threads of a row (and, after the
  // last tile, their shares of the row sum)
  float (*s_mx)[2][128] = reinterpret_cast<float (*)[2][128]>(smem + (3 + 2 * kAtStages) * kAtTile + kAtBarBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qtiles = (p.n_q + kAtBM - 1) / kAtBM;
  const int qt = n_qtiles - 1 - (int)blockIdx.x;  // heaviest (latest) tiles first
  const int q0 = qt * kAtBM;
  const int h = blockIdx.y;
  const int hk = h / (p.H / p.Hkv);
  const int total_kv = p.q_pos0 + p.n_q;
  const int kv_needed = min(p.q_pos0 + q0 + kAtBM, total_kv);  // keys any row of this tile may see
  const int n_t = (kv_needed + kAtBN - 1) / kAtBN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kAtStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 256);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&p_full[s], 256);
      mbar_init(&pv_done[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS[2] = {tmem_base, tmem_base + 128u};
  const uint32_t tO = tmem_base + 256u;

  if (warp == 0 && (lane == 0 || lane == 16)) {
    // ------------------------------ TMA producers: lane 0 streams K (and Q), lane 16 streams V -----------
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kAtTile);
      tma_load_2d(sQ, &tmQ, q_full, h * p.dh, q0);
      tma_load_2d(sQ + kAtHalf, &tmQ, q_full, h * p.dh + 64, q0);
    }
    uint8_t* dst = lane == 0 ? sK : sV;
    const CUtensorMap* tm = lane == 0 ? &tmK : &tmV;
    uint64_t* full = lane == 0 ? k_full : v_full;
    uint64_t* empty = lane == 0 ? k_empty : v_empty;
    for (int t = 0; t < n_t; ++t) {
      const int s = t % kAtStages;
      mbar_wait(&empty[s], (((uint32_t)(t / kAtStages)) & 1u) ^ 1u, 0x900u + s + (lane ? 8 : 0));
      mbar_arrive_expect_tx(&full[s], kAtTile);
      const int row = hk * p.kv_rows_per_head + t * kAtBN;
      tma_load_2d(dst + s * kAtTile, tm, &full[s], 0, row);
      tma_load_2d(dst + s * kAtTile + kAtHalf, tm, &full[s], 64, row);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128) | (1u << 16);  // B (= V) is MN-major
    auto issue_qk = [&](int t) {
      const int s = t % kAtStages, b = t & 1;
      mbar_wait(&k_full[s], ((uint32_t)(t / kAtStages)) & 1u, 0xA00u + s);
      mbar_wait(&s_empty[b], (((uint32_t)(t >> 1)) & 1u) ^ 1u, 0xA10u + b);
      tc_fence_after();
      const int n_ks = p.dh / 16;  // k-steps of the head dimension actually present
#pragma unroll
      for (int k = 0; k < kAtDH / 16; ++k) {
        if (k >= n_ks) break;
        const uint32_t off = (uint32_t)(k >> 2) * kAtHalf + (uint32_t)(k & 3) * 32u;  // 64-dim half, 32 B per k-step
        tc_mma_f16(tS[b], make_smem_desc_sw128(smem_u32(sQ) + off), make_smem_desc_sw128(smem_u32(sK + s * kAtTile) + off),
                   idesc_qk, k != 0 ? 1u : 0u);
      }
      tc_commit(&s_full[b]);
      tc_commit(&k_empty[s]);  // the K slot is reusable once these MMAs retire
    };
    mbar_wait(q_full, 0, 0xA20u);
    issue_qk(0);
    for (int j = 0; j < n_t; ++j) {
      if (j + 1 < n_t) issue_qk(j + 1);
      const int s = j % kAtStages, pb = j & 1;
      mbar_wait(&p_full[pb], ((uint32_t)(j >> 1)) & 1u, 0xA30u + pb);
      mbar_wait(&v_full[s], ((uint32_t)(j / kAtStages)) & 1u, 0xA40u + s);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < kAtBN / 16; ++k) {
        const uint32_t offp = (uint32_t)(k >> 2) * kAtHalf + (uint32_t)(k & 3) * 32u;  // P: K-major over keys
        const uint32_t offv = (uint32_t)k * 16u * 128u;                                 // V: 16 key rows of 128 B
        tc_mma_f16(tO, make_smem_desc_sw128(smem_u32(sP + pb * kAtTile) + offp),
                   make_smem_desc_sw128_mn(smem_u32(sV + s * kAtTile) + offv, kAtHalf), idesc_pv, (j | k) != 0 ? 1u : 0u);
      }
      tc_commit(&pv_done[pb]);
      tc_commit(&v_empty[s]);
    }
  } else if (warp >= 2) {
    // ------------------------------ softmax + epilogue ------------------------
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int ch = (warp - 2) >> 2;       // which 64-key half of the row this thread handles
    const int row = quad * 32 + lane;     // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t col_off = (uint32_t)(ch * 64);
    const int qpos = p.q_pos0 + q0 + row;
    const float sl2 = p.scale * 1.4426950408889634f;
    const bool poly = g_attn_poly != 0;
    float m_used = -INFINITY;  // row maximum the exponentials are currently taken against (raw score units)
    float l_run = 0.f;         // this thread's share of the row sum
    for (int j = 0; j < n_t; ++j) {
      const int b = j & 1;
      mbar_wait(&s_full[b], ((uint32_t)(j >> 1)) & 1u, 0xB00u + b);
      tc_fence_after();
      uint32_t sv[64];
#pragma unroll
      for (int c = 0; c < 2; ++c)
        tmem_ld_32x32(tS[b] + lane_off + col_off + (uint32_t)(c * 32), *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[b]);
      // mask (causal and past-the-end keys) and row maximum
      const int k0 = j * kAtBN + ch * 64;
      const bool need_mask = (j * kAtBN + kAtBN - 1 > p.q_pos0 + q0) || (j * kAtBN + kAtBN > total_kv);
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float v = __uint_as_float(sv[c]);
        if (need_mask && (k0 + c > qpos || k0 + c >= total_kv)) v = -INFINITY;
        sv[c] = __float_as_uint(v);
        mx = fmaxf(mx, v);
      }
      s_mx[b][ch][row] = mx;
      named_bar_sync(2, 256);
      mx = fmaxf(mx, s_mx[b][ch ^ 1][row]);
      // lazy rescale: keep exponentiating against m_used until the row maximum has grown by > 2^8
      const bool grow = (mx > m_used + 8.0f / sl2) || (m_used == -INFINITY && mx != -INFINITY);
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = fmaxf(m_used, mx);
        const float corr = (m_used == -INFINITY) ? 0.f : exp2f((m_used - m_new) * sl2);
        l_run *= corr;
        if (j > 0) {
          // the O accumulator may only be touched once the previous tile's P V has retired
          mbar_wait(&pv_done[(j - 1) & 1], ((uint32_t)((j - 1) >> 1)) & 1u, 0xB10u);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32(tO + lane_off + col_off + (uint32_t)(c * 32), ov);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * corr);
            tmem_st_32x32(tO + lane_off + col_off + (uint32_t)(c * 32), ov);
          }
          tmem_st_wait();
        }
        m_used = m_new;
      }
      const float m_off = (m_used == -INFINITY) ? 0.f : m_used * sl2;
      // this tile's P buffer is free once the P V of tile j-2 has retired (it has, except under a long stall)
      if (j >= 2) mbar_wait(&pv_done[j & 1], ((uint32_t)((j - 2) >> 1)) & 1u, 0xB18u);
      float rs = 0.f;
      uint8_t* prow = sP + (size_t)(j & 1) * kAtTile + (size_t)ch * kAtHalf + (size_t)row * 128;  // keys [ch*64, +64)
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {  // 16-byte chunks of 8 keys, alternately on the MUFU and the FMA pipe
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = __uint_as_float(sv[c8 * 8 + e]) * sl2 - m_off;
          pv[e] = ((c8 & 1) && poly) ? exp2_poly3(x) : exp2f(x);
          rs += pv[e];
        }
        uint4 o;
        o.x = pack_bf16(pv[0], pv[1]);
        o.y = pack_bf16(pv[2], pv[3]);
        o.z = pack_bf16(pv[4], pv[5]);
        o.w = pack_bf16(pv[6], pv[7]);
        *reinterpret_cast<uint4*>(prow + ((c8 ^ (row & 7)) << 4)) = o;
      }
      l_run += rs;
      fence_proxy_async();  // P was written by the generic proxy; the MMA reads it through the async proxy
      tc_fence_before();
      mbar_arrive(&p_full[j & 1]);
    }
    // ---- epilogue: O / l -> bf16 -> global (two threads share one 256-byte output row)
    float (*s_lsum)[128] = s_mx[n_t & 1];  // the parity buffer the last tile did not use
    s_lsum[ch][row] = l_run;
    named_bar_sync(2, 256);
    const float l_tot = s_lsum[0][row] + s_lsum[1][row];
    mbar_wait(&pv_done[(n_t - 1) & 1], ((uint32_t)((n_t - 1) >> 1)) & 1u, 0xB20u);
    tc_fence_after();
    const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
    const int qr = q0 + row;
    __nv_bfloat16* dst = p.out + (int64_t)qr * (p.H * p.dh) + h * p.dh + ch * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32];
      __syncwarp();
      tmem_ld_32x32(tO + lane_off + col_off + (uint32_t)(c * 32), ov);
      tmem_ld_wait();
      if (qr < p.n_q && ch * 64 + c * 32 < p.dh) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(ov[e + 0]) * inv, __uint_as_float(ov[e + 1]) * inv);
          o.y = pack_bf16(__uint_as_float(ov[e + 2]) * inv, __uint_as_float(ov[e + 3]) * inv);
          o.z = pack_bf16(__uint_as_float(ov[e + 4]) * inv, __uint_as_float(ov[e + 5]) * inv);
          o.w = pack_bf16(__uint_as_float(ov[e + 6]) * inv, __uint_as_float(ov[e + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 32 + e) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace advspec
