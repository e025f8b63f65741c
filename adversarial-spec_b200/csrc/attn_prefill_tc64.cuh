// attn_prefill_tc64.cuh — the tcgen05 prompt attention with 64-key tiles and TWO CTAs per SM.
//
// attn_prefill_tc.cuh (128-key tiles, one CTA per SM) spends ~2.4 us per tile for 0.55 us of tensor work: one
// query tile per SM means S ready -> tcgen05.ld -> row max -> exchange -> exp -> P store -> fence -> P V -> commit
// is a serial chain, and overlapping more of it inside one CTA (double-buffered P, split K/V barriers, exp2 on the
// FMA pipe) measured no faster.  Here a CTA needs half the shared memory and half the TMEM (64-key tiles: Q 32 KB,
// P 16 KB, K x2 and V x1 stages of 16 KB; S0/S1 64 columns each + O 128 = 256 TMEM columns), so two CTAs are
// resident per SM and the hardware interleaves one CTA's softmax with the other's MMAs.
// Same contract, operands and numerics as attn_prefill_tc_kernel (head_dim 64 / 96 / 128).
#pragma once

#include "attn_prefill_tc.cuh"

namespace advspec {

constexpr int kA6BM = 128, kA6BN = 64;
constexpr int kA6QHalf = 128 * 64 * 2;   // bytes of a [128 rows][64 dims] half of Q (16 KB)
constexpr int kA6KHalf = 64 * 64 * 2;    // bytes of a [64 keys][64 dims] half of K or V (8 KB)
constexpr int kA6QTile = 2 * kA6QHalf;   // 32 KB
constexpr int kA6PTile = 128 * 64 * 2;   // 16 KB: P [128 rows][64 keys], one 128-byte swizzle row per query row
constexpr int kA6KTile = 2 * kA6KHalf;   // 16 KB
constexpr int kA6BarBytes = 128;
constexpr int kA6MxBytes = 2 * 2 * 128 * 4;
constexpr int kA6Smem = kA6QTile + kA6PTile + 3 * kA6KTile + kA6BarBytes + kA6MxBytes + 1024;  // K x2, V x1
constexpr int kA6Threads = 320;
static_assert(2 * (kA6Smem + 1024) <= 233472, "two CTAs of the 64-key attention must fit one SM");

__global__ void __launch_bounds__(kA6Threads, 2)
attn_prefill_tc64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, AttnPrefillTcParams p) {
  extern __shared__ uint8_t a6_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a6_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + kA6QTile;
  uint8_t* sK = sP + kA6PTile;             // [2]
  uint8_t* sV = sK + 2 * kA6KTile;         // [1]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kA6KTile);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // [2]
  uint64_t* k_empty = bars + 3;            // [2]
  uint64_t* v_full = bars + 5;             // 1
  uint64_t* v_empty = bars + 6;            // 1
  uint64_t* s_full = bars + 7;             // [2]
  uint64_t* s_empty = bars + 9;            // [2]
  uint64_t* p_full = bars + 11;            // 1
  uint64_t* pv_done = bars + 12;           // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  float (*s_mx)[2][128] = reinterpret_cast<float (*)[2][128]>(reinterpret_cast<uint8_t*>(bars) + kA6BarBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qtiles = (p.n_q + kA6BM - 1) / kA6BM;
  const int qt = n_qtiles - 1 - (int)blockIdx.x;  // heaviest (latest) tiles first
  const int q0 = qt * kA6BM;
  const int h = blockIdx.y;
  const int hk = h / (p.H / p.Hkv);
  const int total_kv = p.q_pos0 + p.n_q;
  const int kv_needed = min(p.q_pos0 + q0 + kA6BM, total_kv);  // keys any row of this tile may see
  const int n_t = (kv_needed + kA6BN - 1) / kA6BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 256);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(p_full, 256);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS[2] = {tmem_base, tmem_base + 64u};
  const uint32_t tO = tmem_base + 128u;

  if (warp == 0 && lane == 0) {
    // ------------------------------ TMA producer ------------------------------
    auto load_k = [&](int t) {
      const int s = t & 1;
      mbar_wait(&k_empty[s], (((uint32_t)(t >> 1)) & 1u) ^ 1u, 0xC00u + s);
      mbar_arrive_expect_tx(&k_full[s], kA6KTile);
      const int row = hk * p.kv_rows_per_head + t * kA6BN;
      tma_load_2d(sK + s * kA6KTile, &tmK, &k_full[s], 0, row);
      tma_load_2d(sK + s * kA6KTile + kA6KHalf, &tmK, &k_full[s], 64, row);
    };
    mbar_arrive_expect_tx(q_full, kA6QTile);
    tma_load_2d(sQ, &tmQ, q_full, h * p.dh, q0);
    tma_load_2d(sQ + kA6QHalf, &tmQ, q_full, h * p.dh + 64, q0);
    load_k(0);
    for (int t = 0; t < n_t; ++t) {
      if (t + 1 < n_t) load_k(t + 1);  // its slot frees when Q K^T of tile t-1 retires: before the V slot does
      mbar_wait(v_empty, (((uint32_t)t) & 1u) ^ 1u, 0xC10u);
      mbar_arrive_expect_tx(v_full, kA6KTile);
      const int row = hk * p.kv_rows_per_head + t * kA6BN;
      tma_load_2d(sV, &tmV, v_full, 0, row);
      tma_load_2d(sV + kA6KHalf, &tmV, v_full, 64, row);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 64);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128) | (1u << 16);  // B (= V) is MN-major
    const int n_ks = p.dh / 16;
    auto issue_qk = [&](int t) {
      const int s = t & 1;
      mbar_wait(&k_full[s], ((uint32_t)(t >> 1)) & 1u, 0xD00u + s);
      mbar_wait(&s_empty[s], (((uint32_t)(t >> 1)) & 1u) ^ 1u, 0xD10u + s);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k >= n_ks) break;
        const uint32_t offq = (uint32_t)(k >> 2) * kA6QHalf + (uint32_t)(k & 3) * 32u;  // 64-dim half, 32 B per k-step
        const uint32_t offk = (uint32_t)(k >> 2) * kA6KHalf + (uint32_t)(k & 3) * 32u;
        tc_mma_f16(tS[s], make_smem_desc_sw128(smem_u32(sQ) + offq),
                   make_smem_desc_sw128(smem_u32(sK + s * kA6KTile) + offk), idesc_qk, k != 0 ? 1u : 0u);
      }
      tc_commit(&s_full[s]);
      tc_commit(&k_empty[s]);
    };
    mbar_wait(q_full, 0, 0xD20u);
    issue_qk(0);
    for (int j = 0; j < n_t; ++j) {
      if (j + 1 < n_t) issue_qk(j + 1);
      mbar_wait(p_full, (uint32_t)j & 1u, 0xD30u);
      mbar_wait(v_full, (uint32_t)j & 1u, 0xD40u);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < kA6BN / 16; ++k) {
        const uint32_t offp = (uint32_t)k * 32u;          // P: K-major over the tile's 64 keys, one swizzle row per query
        const uint32_t offv = (uint32_t)k * 16u * 128u;   // V: 16 key rows of 128 B
        tc_mma_f16(tO, make_smem_desc_sw128(smem_u32(sP) + offp),
                   make_smem_desc_sw128_mn(smem_u32(sV) + offv, kA6KHalf), idesc_pv, (j | k) != 0 ? 1u : 0u);
      }
      tc_commit(pv_done);
      tc_commit(v_empty);
    }
  } else if (warp >= 2) {
    // ------------------------------ softmax + epilogue ------------------------
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int ch = (warp - 2) >> 2;       // which 32-key half of the tile (and 64-dim half of O) this thread handles
    const int row = quad * 32 + lane;     // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int qpos = p.q_pos0 + q0 + row;
    const float sl2 = p.scale * 1.4426950408889634f;
    float m_used = -INFINITY;  // row maximum the exponentials are currently taken against (raw score units)
    float l_run = 0.f;         // this thread's share of the row sum
    for (int j = 0; j < n_t; ++j) {
      const int b = j & 1;
      mbar_wait(&s_full[b], ((uint32_t)(j >> 1)) & 1u, 0xE00u + b);
      tc_fence_after();
      uint32_t sv[32];
      tmem_ld_32x32(tS[b] + lane_off + (uint32_t)(ch * 32), sv);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[b]);
      // mask (causal and past-the-end keys) and row maximum
      const int k0 = j * kA6BN + ch * 32;
      const bool need_mask = (j * kA6BN + kA6BN - 1 > p.q_pos0 + q0) || (j * kA6BN + kA6BN > total_kv);
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float v = __uint_as_float(sv[c]);
        if (need_mask && (k0 + c > qpos || k0 + c >= total_kv)) v = -INFINITY;
        sv[c] = __float_as_uint(v);
        mx = fmaxf(mx, v);
      }
      s_mx[b][ch][row] = mx;
      named_bar_sync(2, 256);
      mx = fmaxf(mx, s_mx[b][ch ^ 1][row]);
      // P smem and the O accumulator are only touched once the previous tile's P V has retired
      if (j > 0) mbar_wait(pv_done, ((uint32_t)(j - 1)) & 1u, 0xE10u);
      // lazy rescale: keep exponentiating against m_used until the row maximum has grown by > 2^8
      const bool grow = (mx > m_used + 8.0f / sl2) || (m_used == -INFINITY && mx != -INFINITY);
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = fmaxf(m_used, mx);
        const float corr = (m_used == -INFINITY) ? 0.f : exp2f((m_used - m_new) * sl2);
        l_run *= corr;
        if (j > 0) {
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32(tO + lane_off + (uint32_t)(ch * 64 + c * 32), ov);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * corr);
            tmem_st_32x32(tO + lane_off + (uint32_t)(ch * 64 + c * 32), ov);
          }
          tmem_st_wait();
        }
        m_used = m_new;
      }
      const float m_off = (m_used == -INFINITY) ? 0.f : m_used * sl2;
      float rs = 0.f;
      uint8_t* prow = sP + (size_t)row * 128;  // this query's 64 keys: one 128-byte swizzled row
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {  // 16-byte chunks of 8 keys
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          pv[e] = exp2f(__uint_as_float(sv[c8 * 8 + e]) * sl2 - m_off);
          rs += pv[e];
        }
        uint4 o;
        o.x = pack_bf16(pv[0], pv[1]);
        o.y = pack_bf16(pv[2], pv[3]);
        o.z = pack_bf16(pv[4], pv[5]);
        o.w = pack_bf16(pv[6], pv[7]);
        *reinterpret_cast<uint4*>(prow + (((ch * 4 + c8) ^ (row & 7)) << 4)) = o;
      }
      l_run += rs;
      fence_proxy_async();  // P was written by the generic proxy; the MMA reads it through the async proxy
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 -> global (two threads share one output row)
    float (*s_lsum)[128] = s_mx[n_t & 1];  // the parity buffer the last tile did not use
    s_lsum[ch][row] = l_run;
    named_bar_sync(2, 256);
    const float l_tot = s_lsum[0][row] + s_lsum[1][row];
    mbar_wait(pv_done, ((uint32_t)(n_t - 1)) & 1u, 0xE20u);
    tc_fence_after();
    const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
    const int qr = q0 + row;
    __nv_bfloat16* dst = p.out + (int64_t)qr * (p.H * p.dh) + h * p.dh + ch * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32];
      __syncwarp();
      tmem_ld_32x32(tO + lane_off + (uint32_t)(ch * 64 + c * 32), ov);
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
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace advspec
