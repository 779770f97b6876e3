// Non-causal softmax attention for heads wider than 64 (head_dim <= 128, stored with a 128-column stride and zero padding):
// the LightningDiT-XL family has head_dim 72 (reference models/lightningdit.py:509-515: hidden 1152 / 16 heads), which no
// 64-wide tile serves.  Same structure as attn_fwd_kernel (attention_sm100.cuh) reduced to ONE 128-query tile per CTA,
// because a 128-wide O accumulator leaves TMEM room for one tile only:
//   TMEM  S [0,128) fp32 scores    P [128,192) bf16 pairs (A operand of P.V)    O [192,320) fp32 accumulator
//   smem  Q 2 x 16 KB (two 64-column swizzle atoms), K and V rings of 2 stages x 32 KB
//   warp 0 TMA loader, warp 1 whole-warp uniform MMA issuer, warp 2 TMEM allocator, warps 4-7 softmax (one query row per thread)
// QK^T walks both 64-column atoms (8 K-steps of 16); P.V has N = 128 with V read MN-major across its two atoms
// (LBO = 16 KB).  Only the first `hd` columns of O are stored: the output stays dense [B*T, H*hd] for the out-projection.
#pragma once
#include "ptx.cuh"

namespace ldmae {

constexpr int kA128Threads = 256;
constexpr int kA128Stages = 2;
constexpr int kA128Atom = 128 * 128;                 // 128 rows x 64 bf16
constexpr int kA128Tile = 2 * kA128Atom;             // 128 rows x 128 bf16
constexpr int kA128SmemBytes = 1024 + kA128Tile + 2 * kA128Stages * kA128Tile + 256 + 1024;   // + [2][128] partial row sums (kHalf)
constexpr int kA128HalfThreads = 384;                // kHalf: eight softmax warps, two threads per query row

struct Attn128Params {
  __nv_bfloat16* out;     // [B*T, ldo], head h at columns h*hd .. h*hd+hd
  float* lse2;            // [B, H, T] or nullptr
  int T, H, ldo, hd;      // hd = real head_dim (multiple of 8, <= 128)
  int q_col, k_col, v_col;
  float scale_log2;
  int poly;               // kHalf: every 4th pair of exponentials on the FMA pipe (degree-3 polynomial) instead of the MUFU unit
  float m0_log2;          // > 0: |score * scale_log2| <= m0_log2 is known (qk-normed heads): the constant replaces the running row
                          // maximum -- no max pass, no rescaling of O (softmax is shift invariant); <= 0: running maximum
};

// kHalf (needs the constant softmax offset, m0_log2 > 0): EIGHT softmax warps, two threads per query row -- warp w and w + 4
// share the TMEM lanes of 32 rows and take 64 of the block's 128 score columns each.  Without a running maximum the two halves
// of a row never talk to each other until the row sums are added at the end, and a second warp per SM sub-partition overlaps
// one warp's tcgen05.ld / FMA / pack / tcgen05.st work with the other's MUFU queue (one warp per sub-partition left the
// kernel at ~2.4x its MUFU floor).
template <bool kHalf>
__global__ void __launch_bounds__(kHalf ? kA128HalfThreads : kA128Threads, 1)
attn_fwd_hd128_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const Attn128Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kA128Tile;
  uint8_t* sV = sK + kA128Stages * kA128Tile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kA128Stages * kA128Tile);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kA128Stages;
  uint64_t* v_full = k_empty + kA128Stages;
  uint64_t* v_empty = v_full + kA128Stages;
  uint64_t* s_full = v_empty + kA128Stages;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_free + 1;
  uint64_t* o_done = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);
  float* lsum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);      // [2][128] (kHalf)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qblk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.T + 127) / 128;
  const int row_base = b * p.T;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_qkv);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kA128Stages; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1); mbar_init(s_free, kHalf ? 8 : 4); mbar_init(p_full, kHalf ? 8 : 4); mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kA128Tile);
      for (int a = 0; a < 2; ++a)
        tma_load_2d(&tmap_qkv, q_full, sQ + a * kA128Atom, p.q_col + head * 128 + a * 64, row_base + qblk * 128);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&k_empty[stage], phase ^ 1, 10);
        mbar_expect_tx(&k_full[stage], kA128Tile);
        for (int a = 0; a < 2; ++a)
          tma_load_2d(&tmap_qkv, &k_full[stage], sK + stage * kA128Tile + a * kA128Atom, p.k_col + head * 128 + a * 64, row_base + j * 128);
        mbar_wait(&v_empty[stage], phase ^ 1, 11);
        mbar_expect_tx(&v_full[stage], kA128Tile);
        for (int a = 0; a < 2; ++a)
          tma_load_2d(&tmap_qkv, &v_full[stage], sV + stage * kA128Tile + a * kA128Atom, p.v_col + head * 128 + a * 64, row_base + j * 128);
        if (++stage == kA128Stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, false, false);
    // The head occupies a 128-column slot but only its first `hd` columns are non-zero (XL: 72): the contraction of Q.K^T stops
    // after ceil(hd / 16) of the 8 k-steps and P.V produces ceil(hd / 16) * 16 output columns -- exact (the skipped columns are
    // zeros), 5/8 of the tensor-core work for head_dim 72.
    const int ksteps = (p.hd + 15) / 16;
    const uint32_t idesc_pv = umma_idesc_bf16(128, ksteps * 16, false, true);
    const bool issuer = elect_one();
    const uint64_t d_q0 = umma_smem_desc_sw128(smem_u32(sQ), 1024, 0);
    const uint64_t d_k0 = umma_smem_desc_sw128(smem_u32(sK), 1024, 0);
    const uint64_t d_v0 = umma_smem_desc_sw128(smem_u32(sV), 1024, kA128Atom);      // LBO: the next 64-wide atom of d
    auto issue_qk = [&](int kstage) {
      const uint64_t dk = d_k0 + static_cast<uint64_t>(kstage * (kA128Tile >> 4));
      if (issuer) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (a * 4 + k < ksteps)
              umma_bf16<1>(tmem_base, d_q0 + a * (kA128Atom >> 4) + 2 * k, dk + a * (kA128Atom >> 4) + 2 * k, idesc_qk,
                           (a | k) != 0 ? 1u : 0u);
        umma_commit<1>(s_full);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0, 20);
    mbar_wait(&k_full[0], 0, 21);
    tc_fence_after();
    issue_qk(0);
    if (issuer) umma_commit<1>(&k_empty[0]);
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int j = 0; j < nkv; ++j) {
      int nstage = stage + 1; uint32_t nphase = phase;
      if (nstage == kA128Stages) { nstage = 0; nphase ^= 1; }
      if (j + 1 < nkv) {
        mbar_wait(&k_full[nstage], nphase, 25);
        mbar_wait(s_free, j & 1, 26);
        tc_fence_after();
        issue_qk(nstage);
        if (issuer) umma_commit<1>(&k_empty[nstage]);
        __syncwarp();
      }
      mbar_wait(&v_full[stage], phase, 24);
      mbar_wait(p_full, j & 1, 22);
      tc_fence_after();
      const uint64_t dv = d_v0 + static_cast<uint64_t>(stage * (kA128Tile >> 4));
      if (issuer) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_ts(tmem_base + 192, tmem_base + 128 + k * 8, dv + 128 * k, idesc_pv, (j != 0 || k != 0) ? 1u : 0u);
        umma_commit<1>(o_done);
        umma_commit<1>(&v_empty[stage]);
      }
      __syncwarp();
      stage = nstage; phase = nphase;
    }
  } else if (kHalf && warp >= 4) {
    const int wq = warp & 3, half = (warp - 4) >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + half * 64, tP = tmem_base + lane_addr + 128 + half * 32, tO = tmem_base + lane_addr + 192;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    const float2 neg2 = make_float2(-p.m0_log2, -p.m0_log2);
    const bool usepoly = p.poly != 0;
    float l_run = 0.f;
#pragma unroll 1
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1, 30);
      __syncwarp();
      tc_fence_after();
      float s[64];
      tmem_ld32(tS, s); tmem_ld32(tS + 32, s + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      const int kvalid = p.T - j * 128 - half * 64;
      if (kvalid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kvalid) s[i] = -INFINITY;
      }
      float2 ls = make_float2(0.f, 0.f);
      uint32_t w[32];
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float2 x = fma2(make_float2(s[i], s[i + 1]), sc2, neg2);
        // a quarter of the exponentials on the FMA pipe (the MUFU unit, 16 / clk / SM, is the floor of this loop)
        const float2 e = (usepoly && (i & 6) == 6) ? ex2_poly2(x) : make_float2(ex2_approx(x.x), ex2_approx(x.y));
        ls = add2(ls, e);
        w[i >> 1] = pack_bf16x2(e.x, e.y);
      }
      l_run += ls.x + ls.y;
      tmem_st32(tP, w);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // row sum = the two halves' partial sums (shared memory, one named barrier over the eight softmax warps)
    lsum[half * 128 + wq * 32 + lane] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l_run = lsum[wq * 32 + lane] + lsum[128 + wq * 32 + lane];
    mbar_wait(o_done, (nkv - 1) & 1, 34);
    __syncwarp();
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const int q_tok = qblk * 128 + wq * 32 + lane;
    if (half == 0 && p.lse2 != nullptr && q_tok < p.T)
      p.lse2[(static_cast<size_t>(b) * p.H + head) * p.T + q_tok] = p.m0_log2 + log2f(l_run);
    __nv_bfloat16* dst = p.out + static_cast<size_t>(row_base + min(q_tok, p.T - 1)) * p.ldo + head * p.hd;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = half * 2 + cc;                     // this thread's 32-column chunks of O
      float o[32];
      tmem_ld32(tO + c * 32, o);
      tmem_ld_wait();
      if (q_tok < p.T) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (c * 32 + q * 8 < p.hd)
            *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) =
                make_uint4(pack_bf16x2(o[8 * q] * inv_l, o[8 * q + 1] * inv_l), pack_bf16x2(o[8 * q + 2] * inv_l, o[8 * q + 3] * inv_l),
                           pack_bf16x2(o[8 * q + 4] * inv_l, o[8 * q + 5] * inv_l), pack_bf16x2(o[8 * q + 6] * inv_l, o[8 * q + 7] * inv_l));
        }
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr, tP = tS + 128, tO = tS + 192;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    float m_used = -INFINITY, l_run = 0.f;
#pragma unroll 1
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1, 30);
      __syncwarp();
      tc_fence_after();
      float s[128];
      tmem_ld32(tS, s); tmem_ld32(tS + 32, s + 32); tmem_ld32(tS + 64, s + 64); tmem_ld32(tS + 96, s + 96);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      const int kvalid = p.T - j * 128;
      if (kvalid < 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kvalid) s[i] = -INFINITY;
      }
      const bool fixed = p.m0_log2 > 0.f;                // warp-uniform
      float m_new = m_used;
      if (!fixed) {
        float m_blk = s[0];
#pragma unroll
        for (int i = 1; i < 128; ++i) m_blk = fmaxf(m_blk, s[i]);
        m_new = fmaxf(m_used, m_blk);
      }
      if (j > 0 && !fixed) {
        // O and l follow the running maximum (every block: this variant favours simplicity over the lazy rescale)
        mbar_wait(o_done, (j - 1) & 1, 32);
        __syncwarp();
        tc_fence_after();
        if (__any_sync(0xffffffffu, m_new > m_used)) {
          const float alpha = ex2_approx((m_used - m_new) * p.scale_log2);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float o[32];
            tmem_ld32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= alpha;
            tmem_st32(tO + c * 32, reinterpret_cast<const uint32_t*>(o));
          }
          tmem_st_wait();
          l_run *= alpha;
        }
      }
      m_used = m_new;
      const float neg = fixed ? -p.m0_log2 : -m_used * p.scale_log2;
      const float2 neg2 = make_float2(neg, neg);
      float2 ls = make_float2(0.f, 0.f);
      uint32_t w[64];
#pragma unroll
      for (int i = 0; i < 128; i += 2) {
        const float2 x = fma2(make_float2(s[i], s[i + 1]), sc2, neg2);
        const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
        ls = add2(ls, e);
        w[i >> 1] = pack_bf16x2(e.x, e.y);
      }
      l_run += ls.x + ls.y;
      tmem_st32(tP, w);
      tmem_st32(tP + 32, w + 32);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_done, (nkv - 1) & 1, 34);
    __syncwarp();
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const int q_tok = qblk * 128 + wq * 32 + lane;
    if (p.lse2 != nullptr && q_tok < p.T)
      p.lse2[(static_cast<size_t>(b) * p.H + head) * p.T + q_tok] =
          (p.m0_log2 > 0.f ? p.m0_log2 : m_used * p.scale_log2) + log2f(l_run);
    __nv_bfloat16* dst = p.out + static_cast<size_t>(row_base + min(q_tok, p.T - 1)) * p.ldo + head * p.hd;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float o[32];
      tmem_ld32(tO + c * 32, o);
      tmem_ld_wait();
      if (q_tok < p.T) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (c * 32 + q * 8 < p.hd)
            *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) =
                make_uint4(pack_bf16x2(o[8 * q] * inv_l, o[8 * q + 1] * inv_l), pack_bf16x2(o[8 * q + 2] * inv_l, o[8 * q + 3] * inv_l),
                           pack_bf16x2(o[8 * q + 4] * inv_l, o[8 * q + 5] * inv_l), pack_bf16x2(o[8 * q + 6] * inv_l, o[8 * q + 7] * inv_l));
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace ldmae
