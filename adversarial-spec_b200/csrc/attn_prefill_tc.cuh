// attn_prefill_tc.cuh — causal flash attention for the prompt on the 5th-gen tensor cores.
//
// One CTA = 128 query rows of one head; keys in tiles of 64; TWO CTAs resident per SM.  Per tile
//   S = Q K^T   tcgen05.mma 128x64x(head_dim), Q and K K-major from TMA-swizzled shared memory, S in TMEM
//   softmax     8 warps, two threads per query row (32 keys each): tcgen05.ld the scores, mask (diagonal and
//               last tiles only), exp2 against a lazily-updated row maximum (rescale O only when the maximum
//               grew by > 2^8), write P (bf16) back to shared memory in the 128B-swizzled K-major layout
//   O += P V    tcgen05.mma 128x128x64, A = P (K-major), B = V straight from its [key][dim] tile as an
//               MN-major operand; O accumulates in TMEM across all tiles
// Warp roles: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-9 softmax/epilogue.
// S is double-buffered in TMEM so Q K^T of tile j+1 runs under the softmax of tile j.
// TMEM columns (256 per CTA): S0 [0,64) S1 [64,128) O [128,256).  Shared memory per CTA: Q 32 KB, P 16 KB,
// K x2 and V x1 stages of 16 KB (101.5 KB with barriers): two CTAs per SM.
// head_dim 128, and 64 / 96 in the same 128-wide tiles: Q K^T issues only head_dim/16 k-steps (whatever the TMA
// box holds past the head is never multiplied), the K/V tensor maps are head_dim wide so the box columns past
// it arrive as zeros, and O's columns past head_dim are not written (Phi-3's 32 x 96 heads).  head_dim 256
// (Gemma) uses attn.cuh.
//
// How it got here (profiles/r02_prefill_experiments.md).  Round 1: 128-key tiles, one CTA per SM, 419 us per
// layer on the 5,068-token prompt, tensor pipe 28 % active: with one query tile per SM the chain S ready ->
// tcgen05.ld -> row max -> exchange -> exp -> P store -> fence -> P V -> commit is serial, and the softmax
// warps were issue-bound on top (a select per element for the causal mask on every tile; exp2f's range test
// and two scalings).  Overlapping more inside one CTA (double-buffered P, separate K / V barriers, half of
// the exponentials on the FMA pipe) measured slower.  What worked: trimming the softmax instruction stream
// (uniform mask branch, one-instruction ex2: 381 us, scheduler slots 52 -> 27 % active) and THEN halving the
// CTA so that two are resident and the hardware interleaves one CTA's softmax with the other's MMAs: 311 us,
// tensor pipe 42 % active.
#pragma once

#include "attn.cuh"
#include "common.cuh"

namespace advspec {

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

struct AttnPrefillTcParams {
  __nv_bfloat16* out;  // [n_q][H*dh]
  int n_q, q_pos0, H, Hkv;
  int kv_rows_per_head;  // kv_stride: row of (kv head hk, token t) in the K/V tensor maps = hk*kv_rows_per_head + t
  float scale;
  int dh;  // head_dim: 64, 96 or 128
};

constexpr int kAtBM = 128, kAtBN = 64;
constexpr int kAtQHalf = 128 * 64 * 2;   // bytes of a [128 rows][64 dims] half of Q (16 KB)
constexpr int kAtKHalf = 64 * 64 * 2;    // bytes of a [64 keys][64 dims] half of K or V (8 KB)
constexpr int kAtQTile = 2 * kAtQHalf;   // 32 KB
constexpr int kAtPTile = 128 * 64 * 2;   // 16 KB: P [128 rows][64 keys], one 128-byte swizzle row per query row
constexpr int kAtKTile = 2 * kAtKHalf;   // 16 KB
constexpr int kAtBarBytes = 128;
constexpr int kAtMxBytes = 2 * 2 * 128 * 4;
constexpr int kAtSmem = kAtQTile + kAtPTile + 3 * kAtKTile + kAtBarBytes + kAtMxBytes + 1024;  // K x2, V x1
constexpr int kAtThreads = 320;
static_assert(2 * (kAtSmem + 1024) <= 233472, "two CTAs of the 64-key attention must fit one SM");

__global__ void __launch_bounds__(kAtThreads, 2)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, AttnPrefillTcParams p) {
  extern __shared__ uint8_t at_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + kAtQTile;
  uint8_t* sK = sP + kAtPTile;             // [2]
  uint8_t* sV = sK + 2 * kAtKTile;         // [1]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAtKTile);
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
  float (*s_mx)[2][128] = reinterpret_cast<float (*)[2][128]>(reinterpret_cast<uint8_t*>(bars) + kAtBarBytes);

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
      mbar_arrive_expect_tx(&k_full[s], kAtKTile);
      const int row = hk * p.kv_rows_per_head + t * kAtBN;
      tma_load_2d(sK + s * kAtKTile, &tmK, &k_full[s], 0, row);
      tma_load_2d(sK + s * kAtKTile + kAtKHalf, &tmK, &k_full[s], 64, row);
    };
    mbar_arrive_expect_tx(q_full, kAtQTile);
    tma_load_2d(sQ, &tmQ, q_full, h * p.dh, q0);
    tma_load_2d(sQ + kAtQHalf, &tmQ, q_full, h * p.dh + 64, q0);
    load_k(0);
    for (int t = 0; t < n_t; ++t) {
      if (t + 1 < n_t) load_k(t + 1);  // its slot frees when Q K^T of tile t-1 retires: before the V slot does
      mbar_wait(v_empty, (((uint32_t)t) & 1u) ^ 1u, 0xC10u);
      mbar_arrive_expect_tx(v_full, kAtKTile);
      const int row = hk * p.kv_rows_per_head + t * kAtBN;
      tma_load_2d(sV, &tmV, v_full, 0, row);
      tma_load_2d(sV + kAtKHalf, &tmV, v_full, 64, row);
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
        const uint32_t offq = (uint32_t)(k >> 2) * kAtQHalf + (uint32_t)(k & 3) * 32u;  // 64-dim half, 32 B per k-step
        const uint32_t offk = (uint32_t)(k >> 2) * kAtKHalf + (uint32_t)(k & 3) * 32u;
        tc_mma_f16(tS[s], make_smem_desc_sw128(smem_u32(sQ) + offq),
                   make_smem_desc_sw128(smem_u32(sK + s * kAtKTile) + offk), idesc_qk, k != 0 ? 1u : 0u);
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
      for (int k = 0; k < kAtBN / 16; ++k) {
        const uint32_t offp = (uint32_t)k * 32u;          // P: K-major over the tile's 64 keys, one swizzle row per query
        const uint32_t offv = (uint32_t)k * 16u * 128u;   // V: 16 key rows of 128 B
        tc_mma_f16(tO, make_smem_desc_sw128(smem_u32(sP) + offp),
                   make_smem_desc_sw128_mn(smem_u32(sV) + offv, kAtKHalf), idesc_pv, (j | k) != 0 ? 1u : 0u);
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
      const int k0 = j * kAtBN + ch * 32;
      const bool need_mask = (j * kAtBN + kAtBN - 1 > p.q_pos0 + q0) || (j * kAtBN + kAtBN > total_kv);
      if (need_mask) {  // CTA-uniform: only the diagonal and last tiles carry a mask
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (k0 + c > qpos || k0 + c >= total_kv) sv[c] = 0xff800000u;  // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(sv[c]));
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
          pv[e] = ex2_approx(fmaf(__uint_as_float(sv[c8 * 8 + e]), sl2, -m_off));
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
