// Flash-style non-causal attention for sm_100a, head_dim 64, bf16 operands, fp32 softmax/accumulate.
// Computes softmax(q k^T * scale) v per (sample, head) -- reference models/lightningdit.py:77
// (F.scaled_dot_product_attention, no mask, dropout 0) and tokenizer/models_mae.py:135-141.
//
// Input is the token-major QKV projection output [B*T, ldq] bf16 (columns q|k|v, each head-major,
// 64 per head) read through ONE 2-D TMA tensor map with a {64 x 128} box, so no head transposes
// exist anywhere.  Output o is token-major [B*T, H*64] bf16 = the A operand of the out-projection.
//
// One CTA = one (sample, head) x 256 queries (two 128-row tiles that ping-pong on the tensor core):
//   warp 0       TMA loader (Q once; K/V 128-key blocks through a 3-deep ring)
//   warp 1       MMA issuer: S_t = Q_t K_j^T (128x128x64) and Oblk_t = P_t V_j (128x64x128) in TMEM
//   warp 2       TMEM allocator
//   warps 4-7    softmax + accumulate for tile 0 (one query row per thread, no shuffles)
//   warps 8-11   same for tile 1
// Online softmax keeps the running max / sum / O accumulator in registers; P goes to smem as the
// K-major A operand (128B swizzle) of the PV MMA; V is consumed MN-major straight from its TMA tile.
#pragma once
#include "ptx.cuh"

namespace ldmae {

constexpr int kAttnThreads = 384;
constexpr int kAttnKVStages = 3;
constexpr int kAttnTileBytes = 128 * 128;  // 128 rows x 64 bf16
constexpr int kAttnSmemBytes = 1024 + 2 * kAttnTileBytes            // Q0,Q1
                               + 2 * kAttnKVStages * kAttnTileBytes  // K,V rings
                               + 2 * 2 * kAttnTileBytes              // P0,P1 (128 x 128 bf16 each)
                               + 256;

struct AttnParams {
  __nv_bfloat16* out;  // [B*T, ldo]
  int T, H, ldo;
  int q_col, k_col, v_col;  // column offsets of the q / k / v sections inside a QKV row
  float scale_log2;         // softmax scale * log2(e)
};

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * kAttnTileBytes;
  uint8_t* sV = sK + kAttnKVStages * kAttnTileBytes;
  uint8_t* sP = sV + kAttnKVStages * kAttnTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kAttnTileBytes);
  uint64_t* q_full = bars;                         // [1]
  uint64_t* k_full = bars + 1;                     // [3]
  uint64_t* k_empty = k_full + kAttnKVStages;      // [3]
  uint64_t* v_full = k_empty + kAttnKVStages;      // [3]
  uint64_t* v_empty = v_full + kAttnKVStages;      // [3]
  uint64_t* s_full = v_empty + kAttnKVStages;      // [2]
  uint64_t* p_full = s_full + 2;                   // [2]
  uint64_t* o_full = p_full + 2;                   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qpair = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.T + 127) / 128;
  const int row_base = b * p.T;                    // first token row of this sample

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_qkv);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kAttnKVStages; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1); mbar_init(&p_full[t], 4); mbar_init(&o_full[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S0 [0,128)  S1 [128,256)  Oblk0 [256,320)  Oblk1 [320,384)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * kAttnTileBytes);
      tma_load_2d(&tmap_qkv, q_full, sQ, p.q_col + head * 64, row_base + qpair * 256);
      tma_load_2d(&tmap_qkv, q_full, sQ + kAttnTileBytes, p.q_col + head * 64, row_base + qpair * 256 + 128);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&k_empty[stage], phase ^ 1, 10);
        mbar_expect_tx(&k_full[stage], kAttnTileBytes);
        tma_load_2d(&tmap_qkv, &k_full[stage], sK + stage * kAttnTileBytes, p.k_col + head * 64, row_base + j * 128);
        mbar_wait(&v_empty[stage], phase ^ 1, 11);
        mbar_expect_tx(&v_full[stage], kAttnTileBytes);
        tma_load_2d(&tmap_qkv, &v_full[stage], sV + stage * kAttnTileBytes, p.v_col + head * 64, row_base + j * 128);
        if (++stage == kAttnKVStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, false, true);   // V is MN-major (d contiguous)
      auto issue_qk = [&](int t, int kstage) {
        const uint32_t qa = smem_u32(sQ + t * kAttnTileBytes);
        const uint32_t ka = smem_u32(sK + kstage * kAttnTileBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16<1>(tmem_base + t * 128, umma_smem_desc_sw128(qa + k * 32, 1024, 0),
                       umma_smem_desc_sw128(ka + k * 32, 1024, 0), idesc_qk, k != 0 ? 1u : 0u);
      };
      auto issue_pv = [&](int t, int vstage) {
        const uint32_t pa = smem_u32(sP + t * 2 * kAttnTileBytes);
        const uint32_t va = smem_u32(sV + vstage * kAttnTileBytes);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 8 x 16 keys; P: two 64-key swizzle atoms of 128 rows each
          umma_bf16<1>(tmem_base + 256 + t * 64,
                       umma_smem_desc_sw128(pa + (k >> 2) * kAttnTileBytes + (k & 3) * 32, 1024, 0),
                       umma_smem_desc_sw128(va + k * 2048, 1024, kAttnTileBytes), idesc_pv, k != 0 ? 1u : 0u);
      };
      mbar_wait(q_full, 0, 20);
      mbar_wait(&k_full[0], 0, 21);
      tc_fence_after();
      issue_qk(0, 0); umma_commit<1>(&s_full[0]);
      issue_qk(1, 0); umma_commit<1>(&s_full[1]);
      umma_commit<1>(&k_empty[0]);
      int stage = 0; uint32_t phase = 0;                 // ring position of block j
      for (int j = 0; j < nkv; ++j) {
        int nstage = stage + 1; uint32_t nphase = phase;
        if (nstage == kAttnKVStages) { nstage = 0; nphase ^= 1; }
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], j & 1, 22 + t);
          if (t == 0) mbar_wait(&v_full[stage], phase, 24);
          tc_fence_after();
          issue_pv(t, stage);
          umma_commit<1>(&o_full[t]);
          if (t == 1) umma_commit<1>(&v_empty[stage]);
          if (j + 1 < nkv) {
            if (t == 0) { mbar_wait(&k_full[nstage], nphase, 25); tc_fence_after(); }
            issue_qk(t, nstage);
            umma_commit<1>(&s_full[t]);
            if (t == 1) umma_commit<1>(&k_empty[nstage]);
          }
        }
        stage = nstage; phase = nphase;
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / accumulate: thread = one query row =====================
    const int t = (warp - 4) >> 2;                       // tile 0 / 1
    const int wq = warp & 3;
    const int r = wq * 32 + lane;                        // row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + t * 128;
    const uint32_t tO = tmem_base + lane_addr + 256 + t * 64;
    uint8_t* myP = sP + t * 2 * kAttnTileBytes + r * 128;
    const int sw = r & 7;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&s_full[t], j & 1, 30 + t);
      tc_fence_after();
      const int kvalid = p.T - j * 128;                  // keys of this block that exist (>=1)
      // pass 1: row max
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_ld32(tS + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < kvalid) m_blk = fmaxf(m_blk, v[i]);
      }
      const float m_new = fmaxf(m_run, m_blk);
      const float alpha = exp2f((m_run - m_new) * p.scale_log2);   // 0 on the first block
      if (j > 0) {
        // fold in Oblk_{j-1} (frame m_run), then move the accumulator to frame m_new
        mbar_wait(&o_full[t], (j - 1) & 1, 32 + t);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[32];
          tmem_ld32(tO + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = (o[c * 32 + i] + v[i]) * alpha;
        }
      }
      // pass 2: p = exp2((s - m_new) * scale), bf16 -> smem (K-major, 128B swizzle), row sum
      const float mscaled = m_new * p.scale_log2;
      float lsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_ld32(tS + c * 32, v);
        tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (c * 32 + i < kvalid) ? exp2f(fmaf(v[i], p.scale_log2, -mscaled)) : 0.f;
          float p1 = (c * 32 + i + 1 < kvalid) ? exp2f(fmaf(v[i + 1], p.scale_log2, -mscaled)) : 0.f;
          lsum += p0 + p1;
          w[i >> 1] = pack_bf16x2(p0, p1);
        }
        uint8_t* dst = myP + (c >> 1) * kAttnTileBytes;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 1) * 4 + q) ^ sw;
          *reinterpret_cast<uint4*>(dst + chunk * 16) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        }
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // last block's PV
    mbar_wait(&o_full[t], (nkv - 1) & 1, 34 + t);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const int q_tok = qpair * 256 + t * 128 + r;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32(tO + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c * 32 + i] = (o[c * 32 + i] + v[i]) * inv_l;
    }
    if (q_tok < p.T) {
      __nv_bfloat16* dst = p.out + static_cast<size_t>(row_base + q_tok) * p.ldo + head * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 8)
        *reinterpret_cast<uint4*>(dst + i) = make_uint4(pack_bf16x2(o[i], o[i + 1]), pack_bf16x2(o[i + 2], o[i + 3]),
                                                        pack_bf16x2(o[i + 4], o[i + 5]), pack_bf16x2(o[i + 6], o[i + 7]));
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace ldmae
