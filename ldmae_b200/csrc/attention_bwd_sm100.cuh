// Backward of the non-causal softmax attention (head_dim 64) for sm_100a -- the gradient of
// F.scaled_dot_product_attention as called at reference models/lightningdit.py:77, needed by
// transport.training_losses -> loss.backward() (reference train_accum.py:215-230).
//
// Inputs are token-major like the forward: qkv [B*T, 3D] bf16 (q,k already normed + rotated), dO [B*T, D] bf16,
// the forward's row statistics lse2[b,h,t] = log2(sum_j exp(s_tj * scale)) and delta[b,h,t] = sum_d dO*O.
// With P = exp2(S*scale*log2e - lse2):   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - delta) * scale,  dQ = dS K,  dK = dS^T Q.
//
// Two instantiations of one kernel, neither needs atomics (results are bit-reproducible):
//   kKV = false: CTA = 128 queries of one (sample, head); walks the key blocks;   rows = queries, writes dQ
//   kKV = true : CTA = 128 keys;                          walks the query blocks; rows = keys,    writes dK and dV
// (S is recomputed in both; 7 tensor-core tile products per block pair instead of FlashAttention-2's 5, in exchange for
// no fp32 atomics on dQ and a single simple pipeline.)
//
// Per CTA: warp 0 TMA loader, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 one row per thread.  TMEM (256 columns, so
// two CTAs share an SM and one CTA's exponentials overlap the other's tensor-core work):
//   [0,64)   S  (fp32)  -> overwritten in place by P  as bf16 pairs (columns 0..31)   = A operand of  dV += P  C_b
//   [64,128) dP (fp32)  -> overwritten in place by dS as bf16 pairs (columns 64..95)  = A operand of  dQ|dK += dS C_a
//   [128,192) accumulator 1 (dQ or dK)     [192,256) accumulator 2 (dV, kKV only)
// tcgen05.mma instructions of one thread execute in order, so issuing S_{i+1} after the accumulations of block i is
// all the write-after-read protection the aliasing needs.
#pragma once
#include "ptx.cuh"

namespace ldmae {

constexpr int kAbThreads = 256;
constexpr int kAbStages = 3;
constexpr int kAbRTile = 128 * 128;     // 128 rows x 64 bf16
constexpr int kAbCTile = 64 * 128;      // 64 rows x 64 bf16
constexpr int kAbStageBytes = 2 * kAbCTile + 512;       // C_a, C_b, lse2[64], delta[64]
constexpr int kAbSmemBytes = 1024 + 2 * kAbRTile + kAbStages * kAbStageBytes + 256;

struct AttnBwdParams {
  const float* lse2;    // [B, H, T] (+ padding)
  const float* delta;   // [B, H, T]
  __nv_bfloat16* dqkv;  // [B*T, ld] output, same column layout as qkv
  int T, H, ld;
  int q_col, k_col, v_col;
  float scale, scale_log2;
};

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <bool kKV>
__global__ void __launch_bounds__(kAbThreads, 2)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv_r, const __grid_constant__ CUtensorMap tm_qkv_c,
                const __grid_constant__ CUtensorMap tm_do_r, const __grid_constant__ CUtensorMap tm_do_c, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sRa = smem;                       // dQ: Q rows    | dKV: K rows
  uint8_t* sRb = smem + kAbRTile;            // dQ: dO rows   | dKV: V rows
  uint8_t* sC = smem + 2 * kAbRTile;         // ring: [C_a | C_b | lse2 | delta]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + kAbStages * kAbStageBytes);
  uint64_t* r_full = bars;
  uint64_t* c_full = bars + 1;               // [stages]
  uint64_t* c_empty = c_full + kAbStages;    // [stages]
  uint64_t* sd_full = c_empty + kAbStages;   // S and dP of the block are in TMEM (MMA -> rows)
  uint64_t* pd_full = sd_full + 1;           // P and dS written back (rows -> MMA)
  uint64_t* acc_done = pd_full + 1;          // all accumulations finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rblk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int ncb = (p.T + 63) / 64;           // column blocks
  const int row_base = b * p.T;
  const size_t vec_base = (static_cast<size_t>(b) * p.H + head) * p.T;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv_r); tma_prefetch_desc(&tm_qkv_c); tma_prefetch_desc(&tm_do_r); tma_prefetch_desc(&tm_do_c);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(r_full, 1);
    for (int s = 0; s < kAbStages; ++s) { mbar_init(&c_full[s], 1); mbar_init(&c_empty[s], 1); }
    mbar_init(sd_full, 1); mbar_init(pd_full, 4); mbar_init(acc_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(r_full, 2 * kAbRTile);
      if constexpr (kKV) {
        tma_load_2d(&tm_qkv_r, r_full, sRa, p.k_col + head * 64, row_base + rblk * 128);
        tma_load_2d(&tm_qkv_r, r_full, sRb, p.v_col + head * 64, row_base + rblk * 128);
      } else {
        tma_load_2d(&tm_qkv_r, r_full, sRa, p.q_col + head * 64, row_base + rblk * 128);
        tma_load_2d(&tm_do_r, r_full, sRb, head * 64, row_base + rblk * 128);
      }
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < ncb; ++i) {
        mbar_wait(&c_empty[stage], phase ^ 1, 10);
        uint8_t* st = sC + stage * kAbStageBytes;
        if constexpr (kKV) {
          mbar_expect_tx(&c_full[stage], 2 * kAbCTile + 512);
          tma_load_2d(&tm_qkv_c, &c_full[stage], st, p.q_col + head * 64, row_base + i * 64);
          tma_load_2d(&tm_do_c, &c_full[stage], st + kAbCTile, head * 64, row_base + i * 64);
          bulk_load_1d(st + 2 * kAbCTile, p.lse2 + vec_base + i * 64, 256, &c_full[stage]);
          bulk_load_1d(st + 2 * kAbCTile + 256, p.delta + vec_base + i * 64, 256, &c_full[stage]);
        } else {
          mbar_expect_tx(&c_full[stage], 2 * kAbCTile);
          tma_load_2d(&tm_qkv_c, &c_full[stage], st, p.k_col + head * 64, row_base + i * 64);
          tma_load_2d(&tm_qkv_c, &c_full[stage], st + kAbCTile, p.v_col + head * 64, row_base + i * 64);
        }
        if (++stage == kAbStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_ss = umma_idesc_bf16(128, 64, false, false);
      constexpr uint32_t idesc_ts = umma_idesc_bf16(128, 64, false, true);    // B = column tile read MN-major (d contiguous)
      const uint32_t ra = smem_u32(sRa), rb = smem_u32(sRb);
      auto issue_scores = [&](int stage) {
        const uint32_t ca = smem_u32(sC + stage * kAbStageBytes), cb = ca + kAbCTile;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16<1>(tmem_base, umma_smem_desc_sw128(ra + k * 32, 1024, 0), umma_smem_desc_sw128(ca + k * 32, 1024, 0), idesc_ss,
                       k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16<1>(tmem_base + 64, umma_smem_desc_sw128(rb + k * 32, 1024, 0), umma_smem_desc_sw128(cb + k * 32, 1024, 0),
                       idesc_ss, k != 0 ? 1u : 0u);
      };
      mbar_wait(r_full, 0, 20);
      int stage = 0; uint32_t phase = 0;
      mbar_wait(&c_full[0], 0, 21);
      tc_fence_after();
      issue_scores(0);
      umma_commit<1>(sd_full);
      for (int i = 0; i < ncb; ++i) {
        mbar_wait(pd_full, i & 1, 22);
        tc_fence_after();
        const uint32_t ca = smem_u32(sC + stage * kAbStageBytes), cb = ca + kAbCTile;
#pragma unroll
        for (int k = 0; k < 4; ++k)     // 4 x 16 columns of the block; A = dS: 8 TMEM columns (16 bf16) per step
          umma_bf16_ts(tmem_base + 128, tmem_base + 64 + k * 8, umma_smem_desc_sw128(ca + k * 2048, 1024, kAbCTile), idesc_ts,
                       (i != 0 || k != 0) ? 1u : 0u);
        if constexpr (kKV) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ts(tmem_base + 192, tmem_base + k * 8, umma_smem_desc_sw128(cb + k * 2048, 1024, kAbCTile), idesc_ts,
                         (i != 0 || k != 0) ? 1u : 0u);
        }
        umma_commit<1>(&c_empty[stage]);
        if (++stage == kAbStages) { stage = 0; phase ^= 1; }
        if (i + 1 < ncb) {
          mbar_wait(&c_full[stage], phase, 23);
          tc_fence_after();
          issue_scores(stage);
          umma_commit<1>(sd_full);
        } else {
          umma_commit<1>(acc_done);
        }
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr, tD = tS + 64;
    const int row = rblk * 128 + wq * 32 + lane;       // token index inside the sample
    float lse_r = 0.f, delta_r = 0.f;
    if constexpr (!kKV) {
      if (row < p.T) { lse_r = __ldg(p.lse2 + vec_base + row); delta_r = __ldg(p.delta + vec_base + row); }
    }
    int stage = 0;
    uint32_t cphase = 0;
#pragma unroll 1
    for (int i = 0; i < ncb; ++i) {
      if constexpr (kKV) mbar_wait(&c_full[stage], cphase, 29);   // this thread reads the stage's lse2 / delta vectors itself
      mbar_wait(sd_full, i & 1, 30);
      __syncwarp();
      tc_fence_after();
      const float* vl = reinterpret_cast<const float*>(sC + stage * kAbStageBytes + 2 * kAbCTile);
      const float* vd = vl + 64;
      const int cvalid = p.T - i * 64;                 // columns of this block that exist
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float s[32], dp[32];
        tmem_ld32(tS + hf * 32, s);
        tmem_ld32(tD + hf * 32, dp);
        tmem_ld_wait();
        uint32_t wp[16], wd[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float l0, l1, d0, d1;
          if constexpr (kKV) {
            const float2 l2 = *reinterpret_cast<const float2*>(vl + hf * 32 + c);
            const float2 d2 = *reinterpret_cast<const float2*>(vd + hf * 32 + c);
            l0 = l2.x; l1 = l2.y; d0 = d2.x; d1 = d2.y;
          } else {
            l0 = l1 = lse_r; d0 = d1 = delta_r;
          }
          const bool ok0 = hf * 32 + c < cvalid, ok1 = hf * 32 + c + 1 < cvalid;   // columns beyond the sample: P = dS = 0
          const float p0 = ok0 ? ex2_approx(fmaf(s[c], p.scale_log2, -l0)) : 0.f;
          const float p1 = ok1 ? ex2_approx(fmaf(s[c + 1], p.scale_log2, -l1)) : 0.f;
          const float e0 = ok0 ? p0 * (dp[c] - d0) * p.scale : 0.f;
          const float e1 = ok1 ? p1 * (dp[c + 1] - d1) * p.scale : 0.f;
          wp[c >> 1] = pack_bf16x2(p0, p1);
          wd[c >> 1] = pack_bf16x2(e0, e1);
        }
        if constexpr (kKV) tmem_st16(tS + hf * 16, wp);
        tmem_st16(tD + hf * 16, wd);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pd_full);
      if (++stage == kAbStages) { stage = 0; cphase ^= 1; }
    }
    // epilogue: accumulators -> bf16 rows of dqkv (the tcgen05.ld is warp-collective: only the stores are predicated)
    mbar_wait(acc_done, 0, 31);
    __syncwarp();
    tc_fence_after();
#pragma unroll
    for (int a = 0; a < (kKV ? 2 : 1); ++a) {
      const int col = (kKV ? (a == 0 ? p.k_col : p.v_col) : p.q_col) + head * 64;
      __nv_bfloat16* dst = p.dqkv + static_cast<size_t>(row_base + min(row, p.T - 1)) * p.ld + col;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float o[32];
        tmem_ld32(tS + 128 + a * 64 + c * 32, o);
        tmem_ld_wait();
        if (row < p.T) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) =
                make_uint4(pack_bf16x2(o[8 * q], o[8 * q + 1]), pack_bf16x2(o[8 * q + 2], o[8 * q + 3]),
                           pack_bf16x2(o[8 * q + 4], o[8 * q + 5]), pack_bf16x2(o[8 * q + 6], o[8 * q + 7]));
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 256);
}

// delta[b,h,t] = sum_d dO[t,h,d] * O[t,h,d]  (one warp per token row: lanes walk the row, heads are 64 wide)
__global__ void attn_delta_kernel(float* __restrict__ delta, const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O,
                                  int B, int T, int H) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * T) return;
  const int b = row / T, t = row % T;
  const size_t base = static_cast<size_t>(row) * H * 64;
  for (int h = 0; h < H; ++h) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(dO + base + h * 64 + lane * 2);
    const __nv_bfloat162 o = *reinterpret_cast<const __nv_bfloat162*>(O + base + h * 64 + lane * 2);
    float s = __bfloat162float(a.x) * __bfloat162float(o.x) + __bfloat162float(a.y) * __bfloat162float(o.y);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) delta[(static_cast<size_t>(b) * H + h) * T + t] = s;
  }
}

}  // namespace ldmae
