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
// Per CTA (one per SM): warp 0 TMA loader, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-19 one row per thread and 16 of
// the 64 columns of a block per warp.  TMEM (512 columns), S / dP double-buffered so that the tensor core computes the scores of
// block i+1 and the accumulations of block i-1 while the row warps work on block i:
//   buffer b (b = i & 1) at column 128 b:  [0,64) S (fp32)   [64,128) dP (fp32)
//       each warp overwrites the first 8 of the 16 S columns it has just read with its P (16 bf16) -- the A operand of
//       dV += P C_b (K-step k at column 16 k) -- and likewise dS inside dP, the A operand of dQ|dK += dS C_a
//   [256,320) accumulator 1 (dQ or dK)     [320,384) accumulator 2 (dV, kKV only)
//   [384,416) row operand a (Q | K rows, bf16)   [416,448) row operand b (dO | V rows): A operands of the score products
// tcgen05.mma instructions of one thread execute in order: scores(i+1) is issued after the accumulations of block i-1, the
// previous user of its buffer, which is all the write-after-read protection the aliasing needs.
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace ldmae {

constexpr int kAbThreads = 640;        // 4 control warps + 16 row warps (four 16-column quarters per TMEM lane quarter)
// HD = head width in the qkv buffer: 64 (tuned path) or 128 (heads wider than 64, zero padded: LightningDiT-XL's 72).
template <int HD>
struct AbGeo {
  static constexpr int kAtoms = HD / 64;                  // 64-column swizzle atoms per tile row
  static constexpr int kStages = HD == 64 ? 4 : 3;
  static constexpr int kRAtom = 128 * 128;                // 128 rows x 64 bf16
  static constexpr int kCAtom = 64 * 128;                 // 64 rows x 64 bf16
  static constexpr int kRTile = kAtoms * kRAtom;
  static constexpr int kCTile = kAtoms * kCAtom;
  static constexpr int kStageBytes = 2 * kCTile + 512;    // C_a, C_b, lse2[64], delta[64]
  static constexpr int kSmemBytes = 1024 + 2 * kRTile + kStages * kStageBytes + 256;
  static constexpr int kAcc1 = 256, kAcc2 = 256 + HD;     // TMEM columns of the accumulators
};

// Optional phase tracing of CTA 0 (debug builds: -DLDMAE_ATTN_TRACE): clock64 stamps per column block of its first two items
// (rows 0..15 and 16..31 of the buffer: the gap between them is the item boundary);
// trace[i][0..3] = MMA warp (c_full seen, scores issued, pd_full seen, accumulations issued), [4..7] = row warp 4
// (sd_full seen, TMEM loaded, math done, P/dS stored).
#ifdef LDMAE_ATTN_TRACE
#define ABWD_STAMP(k) do { if (p.trace && blockIdx.x == 0 && it < 2 && lane == 0 && i < 16) \
  p.trace[(it * 16 + i) * 8 + (k)] = clock64(); } while (0)
// epilogue stamps of row warp 4 go into the two MMA slots the item's last block leaves unused
#define ABWD_ESTAMP(k) do { if (p.trace && blockIdx.x == 0 && it < 2 && lane == 0 && warp == 4) \
  p.trace[(it * 16 + 15) * 8 + (k)] = clock64(); } while (0)
#else
#define ABWD_STAMP(k) do { } while (0)
#define ABWD_ESTAMP(k) do { } while (0)
#endif

struct AttnBwdParams {
  long long* trace;     // [32][8] or nullptr (debug builds)
  const float* nlse2;   // [B, H, T] (+ padding): -lse2 of the forward
  const float* delta;   // [B, H, T] (+ padding): scale * sum_d dO * O
  __nv_bfloat16* dqkv;  // [B*T, ld] output, same column layout as qkv
  int T, H, ld;
  int q_col, k_col, v_col;
  int hd;               // real head_dim (<= HD): columns beyond it are stored as zeros
  float scale, scale_log2;
};

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <bool kKV, int HD = 64>
__global__ void __launch_bounds__(kAbThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv_r, const __grid_constant__ CUtensorMap tm_qkv_c,
                const __grid_constant__ CUtensorMap tm_do_r, const __grid_constant__ CUtensorMap tm_do_c, const AttnBwdParams p,
                const int n_rblk, const int n_items) {
  // Work items (row block, head, sample), row block fastest; CTA c takes items c, c + gridDim.x, ...: with one CTA per SM the
  // kernel is persistent -- TMEM and barriers are set up once, the loader runs ahead across item boundaries (next row tiles
  // and column ring in flight under the current item's tail), and barrier phases follow running counters:
  //   g  column blocks processed so far (buffer g & 1, phase (g >> 1) & 1 of sd_full / pd_full)      it  items so far
  using Geo = AbGeo<HD>;
  constexpr int kAbStages = Geo::kStages, kAbRTile = Geo::kRTile, kAbCTile = Geo::kCTile, kAbStageBytes = Geo::kStageBytes;
  constexpr int kAtoms = Geo::kAtoms;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sRa = smem;                       // dQ: Q rows    | dKV: K rows
  uint8_t* sRb = smem + kAbRTile;            // dQ: dO rows   | dKV: V rows
  uint8_t* sC = smem + 2 * kAbRTile;         // ring: [C_a | C_b | lse2 | delta]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + kAbStages * kAbStageBytes);
  uint64_t* r_full = bars;
  uint64_t* c_full = bars + 1;               // [stages]
  uint64_t* c_empty = c_full + kAbStages;    // [stages]
  uint64_t* sd_full = c_empty + kAbStages;   // [2] S and dP of the block are in TMEM buffer b (MMA -> rows)
  uint64_t* pd_full = sd_full + 2;           // [2] P and dS written back into buffer b (rows -> MMA)
  uint64_t* acc_done = pd_full + 2;          // all accumulations finished
  uint64_t* ra_ready = acc_done + 1;         // row operands copied into TMEM (rows -> MMA)
  uint64_t* r_empty = ra_ready + 1;          // the row tiles in shared memory may be reloaded (HD 64: after the TMEM copy, by the
                                             // copying warps; HD 128: after the item's last score product, by the MMA warp)
  uint64_t* acc_free = r_empty + 1;          // the epilogue has read the accumulators (rows -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ncb = (p.T + 63) / 64;           // column blocks
  auto item_coords = [&](int w, int& rblk, int& head, int& b) {
    rblk = w % n_rblk;
    head = (w / n_rblk) % p.H;
    b = w / (n_rblk * p.H);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv_r); tma_prefetch_desc(&tm_qkv_c); tma_prefetch_desc(&tm_do_r); tma_prefetch_desc(&tm_do_c);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(r_full, 1);
    for (int s = 0; s < kAbStages; ++s) { mbar_init(&c_full[s], 1); mbar_init(&c_empty[s], 1); }
    for (int q = 0; q < 2; ++q) { mbar_init(&sd_full[q], 1); mbar_init(&pd_full[q], 16); }
    mbar_init(acc_done, 1); mbar_init(ra_ready, 8); mbar_init(r_empty, HD == 64 ? 8 : 1); mbar_init(acc_free, 16);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        int rblk, head, b;
        item_coords(w, rblk, head, b);
        const int row_base = b * p.T;
        const size_t vec_base = (static_cast<size_t>(b) * p.H + head) * p.T;
        // Row tiles (Q | K and dO | V rows of the item).  HD == 64: the row warps copy them into TMEM at once and hand the
        // shared-memory tiles back (r_empty), so the NEXT item's tiles are requested early in this item's column loop
        // (below) -- requested after the last column block they arrived ~2300 clk after the row warps wanted them.
        auto load_rows = [&](int rblk_, int head_, int row_base_) {
          mbar_expect_tx(r_full, 2 * kAbRTile);
          // (dO is dense [B*T, H*hd]: for hd < HD its second atom runs into the next head's columns -- harmless: those columns
          //  only meet the zero padding of V in dP, and the matching dV columns are stored as zeros)
          for (int a = 0; a < kAtoms; ++a) {
            if constexpr (kKV) {
              tma_load_2d(&tm_qkv_r, r_full, sRa + a * Geo::kRAtom, p.k_col + head_ * HD + a * 64, row_base_ + rblk_ * 128);
              tma_load_2d(&tm_qkv_r, r_full, sRb + a * Geo::kRAtom, p.v_col + head_ * HD + a * 64, row_base_ + rblk_ * 128);
            } else {
              tma_load_2d(&tm_qkv_r, r_full, sRa + a * Geo::kRAtom, p.q_col + head_ * HD + a * 64, row_base_ + rblk_ * 128);
              tma_load_2d(&tm_do_r, r_full, sRb + a * Geo::kRAtom, head_ * p.hd + a * 64, row_base_ + rblk_ * 128);
            }
          }
        };
        if (HD != 64 || it == 0) {
          if (it > 0) mbar_wait(r_empty, (it - 1) & 1, 12);
          load_rows(rblk, head, row_base);
        }
        for (int i = 0; i < ncb; ++i) {
          mbar_wait(&c_empty[stage], phase ^ 1, 10);
          uint8_t* st = sC + stage * kAbStageBytes;
          if constexpr (kKV) {
            mbar_expect_tx(&c_full[stage], 2 * kAbCTile + 512);
            for (int a = 0; a < kAtoms; ++a) {
              tma_load_2d(&tm_qkv_c, &c_full[stage], st + a * Geo::kCAtom, p.q_col + head * HD + a * 64, row_base + i * 64);
              tma_load_2d(&tm_do_c, &c_full[stage], st + kAbCTile + a * Geo::kCAtom, head * p.hd + a * 64, row_base + i * 64);
            }
            bulk_load_1d(st + 2 * kAbCTile, p.nlse2 + vec_base + i * 64, 256, &c_full[stage]);
            bulk_load_1d(st + 2 * kAbCTile + 256, p.delta + vec_base + i * 64, 256, &c_full[stage]);
          } else {
            mbar_expect_tx(&c_full[stage], 2 * kAbCTile);
            for (int a = 0; a < kAtoms; ++a) {
              tma_load_2d(&tm_qkv_c, &c_full[stage], st + a * Geo::kCAtom, p.k_col + head * HD + a * 64, row_base + i * 64);
              tma_load_2d(&tm_qkv_c, &c_full[stage], st + kAbCTile + a * Geo::kCAtom, p.v_col + head * HD + a * 64, row_base + i * 64);
            }
          }
          if (++stage == kAbStages) { stage = 0; phase ^= 1; }
          // (at the ring depth, this item's first block has been consumed, hence its row tiles were copied long ago: the wait
          //  below does not hold the column loads up -- at i == 0 it stalled the ring for most of an item)
          if (HD == 64 && i == min(kAbStages, ncb - 1) && w + static_cast<int>(gridDim.x) < n_items) {
            int rblk_n, head_n, b_n;
            item_coords(w + static_cast<int>(gridDim.x), rblk_n, head_n, b_n);
            mbar_wait(r_empty, it & 1, 12);                  // this item's tiles have been copied into TMEM
            load_rows(rblk_n, head_n, b_n * p.T);
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer.  The whole warp runs the control flow (so that descriptors and TMEM addresses stay warp-uniform and live in
    // uniform registers); only the tcgen05 instructions themselves are issued by one elected lane.
    constexpr uint32_t idesc_ss = umma_idesc_bf16(128, 64, false, false);
    // wide heads (HD = 128 slots, real head_dim p.hd, e.g. 72): the padding columns are zeros, so the contractions over the head
    // dimension stop after ceil(hd / 16) k-steps and the accumulations produce ceil(hd / 16) * 16 columns -- exact, 5/8 of the
    // tensor-core work for head_dim 72 (the epilogue never stores accumulator columns >= hd)
    const int ksteps = HD == 64 ? 4 : (p.hd + 15) / 16;
    const uint32_t idesc_ts = umma_idesc_bf16(128, HD == 64 ? 64 : ksteps * 16, false, true);    // B = column tile read MN-major (d contiguous)
    const bool issuer = elect_one();
    // one descriptor per tile; K-steps advance the 14-bit start-address field: +2 (32 B) K-major, +128 (2 KB) MN-major.
    // (The leading-dimension offset is unused in both forms: a single 64-wide swizzle atom along the other dimension.)
    // (MN-major form: LBO = distance between the 64-wide atoms of d; unused when HD = 64)
    const uint64_t d_c0 = umma_smem_desc_sw128(smem_u32(sC), 1024, Geo::kCAtom);
    const uint64_t d_ra = umma_smem_desc_sw128(smem_u32(sRa), 1024, 0), d_rb = umma_smem_desc_sw128(smem_u32(sRb), 1024, 0);
    auto issue_scores = [&](int stage, int buf) {
      const uint64_t d_ca = d_c0 + static_cast<uint64_t>((stage * kAbStageBytes) >> 4), d_cb = d_ca + (kAbCTile >> 4);
      const uint32_t td = tmem_base + buf * 128;
      if (issuer) {
        if constexpr (HD == 64) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(td, tmem_base + 384 + k * 8, d_ca + 2 * k, idesc_ss, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(td + 64, tmem_base + 416 + k * 8, d_cb + 2 * k, idesc_ss, k != 0 ? 1u : 0u);
        } else {
          // wide heads: TMEM is full (scores 256 + two 128-wide accumulators), both operands come from shared memory
#pragma unroll
          for (int a = 0; a < kAtoms; ++a)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (a * 4 + k < ksteps)
                umma_bf16<1>(td, d_ra + a * (Geo::kRAtom >> 4) + 2 * k, d_ca + a * (Geo::kCAtom >> 4) + 2 * k, idesc_ss, (a | k) != 0 ? 1u : 0u);
#pragma unroll
          for (int a = 0; a < kAtoms; ++a)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (a * 4 + k < ksteps)
                umma_bf16<1>(td + 64, d_rb + a * (Geo::kRAtom >> 4) + 2 * k, d_cb + a * (Geo::kCAtom >> 4) + 2 * k, idesc_ss, (a | k) != 0 ? 1u : 0u);
        }
        umma_commit<1>(&sd_full[buf]);
      }
      __syncwarp();
    };
    int stage = 0; uint32_t phase = 0;                 // ring position of the current block
    int g = 0, it = 0;
    // This warp's per-block control path (two barrier waits, 8 + 4..8 MMAs, commits) sits on the serial chain
    // rows(g-1) -> accumulate(g-1) -> scores(g+1): barrier addresses are computed once and the per-block waits trap silently.
    uint32_t m_c_full = smem_u32(c_full), m_pd_full = smem_u32(pd_full);
    asm volatile("" : "+r"(m_c_full), "+r"(m_pd_full));
    // kAhead (HD == 64): the first scores of the NEXT item are issued during the current item's last block (the row warps copy
    // the next item's row operands into TMEM as soon as they hold the last scores), so the next item's first block is in
    // TMEM when the row warps come back from the epilogue.  Issued only after the item's last accumulation, they arrived
    // 4600 clk after it (clock64 trace: 24 % of an item of 16 blocks).
    constexpr bool kAhead = HD == 64;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const bool has_next = w + static_cast<int>(gridDim.x) < n_items;
      if (!kAhead || it == 0) {
        if constexpr (HD == 64) mbar_wait(ra_ready, it & 1, 20); else mbar_wait(r_full, it & 1, 20);
        mbar_wait(&c_full[stage], phase, 21);
        tc_fence_after();
        // (buffer g & 1 was last used by block g - 2, whose accumulations were issued before this point)
        issue_scores(stage, g & 1);
      }
      if (HD != 64 && ncb == 1) { if (issuer) umma_commit<1>(r_empty); __syncwarp(); }
      for (int i = 0; i < ncb; ++i, ++g) {
        int nstage = stage + 1; uint32_t nphase = phase;
        if (nstage == kAbStages) { nstage = 0; nphase ^= 1; }
        if (i + 1 < ncb) {
          // scores of block i+1 into the other buffer (its previous user, block g-1, was consumed by MMAs issued earlier)
          mbar_wait_quiet_a(m_c_full + nstage * 8, nphase);
          tc_fence_after();
          ABWD_STAMP(0);
          issue_scores(nstage, (g + 1) & 1);
          if (HD != 64 && i + 2 == ncb) { if (issuer) umma_commit<1>(r_empty); __syncwarp(); }   // last read of the row tiles
          ABWD_STAMP(1);
        } else if (kAhead && has_next) {
          // the next item's first block (its row operands are being copied into TMEM by the row warps right now)
          mbar_wait(ra_ready, (it + 1) & 1, 20);
          mbar_wait(&c_full[nstage], nphase, 21);
          tc_fence_after();
          issue_scores(nstage, (g + 1) & 1);
        }
        mbar_wait_quiet_a(m_pd_full + (g & 1) * 8, (g >> 1) & 1);
        // the first accumulation of an item overwrites the accumulators: the previous item's epilogue must have read them
        if (i == 0 && it > 0) mbar_wait(acc_free, (it - 1) & 1, 27);
        tc_fence_after();
        ABWD_STAMP(2);
        const uint64_t d_ca = d_c0 + static_cast<uint64_t>((stage * kAbStageBytes) >> 4), d_cb = d_ca + (kAbCTile >> 4);
        const uint32_t tb = tmem_base + (g & 1) * 128;
        const uint32_t acc_first = i != 0 ? 1u : 0u;
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 4; ++k)     // 4 x 16 columns of the block; A = dS: 16 bf16 = 8 TMEM columns at column 16 k
            umma_bf16_ts(tmem_base + Geo::kAcc1, tb + 64 + k * 16, d_ca + 128 * k, idesc_ts, k != 0 ? 1u : acc_first);
          if constexpr (kKV) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_base + Geo::kAcc2, tb + k * 16, d_cb + 128 * k, idesc_ts, k != 0 ? 1u : acc_first);
          }
          umma_commit<1>(&c_empty[stage]);
          if (i == ncb - 1) umma_commit<1>(acc_done);
        }
        __syncwarp();
        ABWD_STAMP(3);
        stage = nstage; phase = nphase;
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;                           // TMEM lane quarter
    const int cq = (warp - 4) >> 2;                    // which 16 of the block's 64 columns this warp works on
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const float2 sl2 = make_float2(p.scale_log2, p.scale_log2), sc2 = make_float2(p.scale, p.scale);
    int stage = 0;
    uint32_t cphase = 0;
    int g = 0, it = 0;
    // shared addresses of the loop's barriers, computed once and kept opaque (a generic -> shared conversion in front of every
    // barrier operation and the printf path of the checked wait were a measurable share of this issue-bound loop; the same
    // change was worth 10 % in the forward kernel)
    uint32_t a_c_full = smem_u32(c_full), a_sd_full = smem_u32(sd_full), a_pd_full = smem_u32(pd_full);
    asm volatile("" : "+r"(a_c_full), "+r"(a_sd_full), "+r"(a_pd_full));
    constexpr bool kAhead = HD == 64;                  // see the MMA warp: the next item's operands are prepared one block early
    float a_n = 0.f, d_n = 0.f;                        // kAhead: the next item's row statistics, loaded (not yet used) one item early
    // copies this warp's 32 rows x 64 columns of a row tile (Q | K or dO | V) from its TMA tile into TMEM
    auto copy_row_operands = [&](int item) {
      mbar_wait(r_full, item & 1, 28);
      const uint8_t* src = (cq == 0 ? sRa : sRb) + (wq * 32 + lane) * 128;
      uint32_t wr[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 v = lds128(src + ((j ^ (lane & 7)) << 4));
        wr[4 * j] = v.x; wr[4 * j + 1] = v.y; wr[4 * j + 2] = v.z; wr[4 * j + 3] = v.w;
      }
      tmem_st32(tmem_base + lane_addr + 384 + cq * 32, wr);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(ra_ready); mbar_arrive(r_empty); }
    };
    // Item coordinates advance incrementally (w += gridDim.x): the two integer divisions of item_coords, run by sixteen in-order
    // warps at once, were ~1000 clk of every item boundary (clock64 trace).
    const int G = static_cast<int>(gridDim.x);
    const int d_r = G % n_rblk, d_h = (G / n_rblk) % p.H, d_b = G / (n_rblk * p.H);
    auto advance = [&](int& r_, int& h_, int& b_) {
      r_ += d_r; if (r_ >= n_rblk) { r_ -= n_rblk; ++h_; }
      h_ += d_h; if (h_ >= p.H) { h_ -= p.H; ++b_; }
      b_ += d_b;
    };
    int rblk, head, b;
    item_coords(blockIdx.x, rblk, head, b);
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
    int rblk_n = rblk, head_n = head, b_n = b;
    advance(rblk_n, head_n, b_n);
    const bool has_next = w + static_cast<int>(gridDim.x) < n_items;
    const int row_base = b * p.T;
    const size_t vec_base = (static_cast<size_t>(b) * p.H + head) * p.T;
    const int row = rblk * 128 + wq * 32 + lane;       // token index inside the sample
    float2 nl_r = make_float2(0.f, 0.f), dl_r = make_float2(0.f, 0.f);
    if constexpr (!kKV) {
      if (kAhead && it > 0) { nl_r = make_float2(a_n, a_n); dl_r = make_float2(-d_n, -d_n); }
      else if (row < p.T) {
        const float a = __ldg(p.nlse2 + vec_base + row), d = __ldg(p.delta + vec_base + row);
        nl_r = make_float2(a, a); dl_r = make_float2(-d, -d);
      }
    }
    // The row operands (Q | K and dO | V rows of this item) are the A operand of every score product: copy them once
    // from their TMA tiles into TMEM (bf16 pairs, one row per lane), so that the score MMAs read only the 2 KB column
    // tile from shared memory per instruction (an SS product of this shape is shared-memory-bandwidth bound).
    // (kAhead: items after the CTA's first were copied during the previous item's last block, below.)
    if (HD == 64 && cq < 2 && (!kAhead || it == 0)) copy_row_operands(it);
    if constexpr (!kKV) {
      if (kAhead && has_next) {
        // the next item's row statistics: global loads issued a whole item ahead, first used after this item's epilogue (used
        // right after the load they stalled every item boundary for a DRAM round trip)
        const int row_n = rblk_n * 128 + wq * 32 + lane;
        a_n = 0.f; d_n = 0.f;
        if (row_n < p.T) {
          const size_t vb = (static_cast<size_t>(b_n) * p.H + head_n) * p.T + row_n;
          a_n = __ldg(p.nlse2 + vb); d_n = __ldg(p.delta + vb);
        }
      }
    }
#pragma unroll 1
    for (int i = 0; i < ncb; ++i, ++g) {
      if constexpr (kKV) mbar_wait_quiet_a(a_c_full + stage * 8, cphase);   // this thread reads the stage's statistics vectors itself
      mbar_wait_quiet_a(a_sd_full + (g & 1) * 8, (g >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      if (warp == 4) ABWD_STAMP(4);
      const uint32_t tS = tmem_base + lane_addr + (g & 1) * 128 + cq * 16, tD = tS + 64;
      const float* vl = reinterpret_cast<const float*>(sC + stage * kAbStageBytes + 2 * kAbCTile) + cq * 16;
      const float* vd = vl + 64;
      const int cvalid = p.T - i * 64 - cq * 16;       // columns of this warp's quarter that exist
      float sv[16], dp[16];
      tmem_ld16(tS, sv);
      tmem_ld16(tD, dp);
      tmem_ld_wait();
      if (warp == 4) ABWD_STAMP(5);
      if (kAhead && i == ncb - 1 && has_next) {
        // this item's last scores are in registers (every score product of the item is complete): hand the next item's row
        // operands to the tensor core now
        if (cq < 2) copy_row_operands(it + 1);
      }
      uint32_t wp[8], wd[8];
      // The ragged tail of a sample (fewer than 16 live columns in this warp's quarter) is a separate instantiation of
      // the body behind a warp-uniform branch: as predicated selects inside the common body it cost 48 issue slots per
      // block in a loop that is issue-bound.
      auto body = [&](auto ragged) {
        constexpr bool kRagged = decltype(ragged)::value;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          float2 nl0 = nl_r, nl1 = nl_r, dl0 = dl_r, dl1 = dl_r;
          if constexpr (kKV) {
            const uint4 lw = lds128(vl + c), dw = lds128(vd + c);          // shared-space loads (the pointers are generic)
            nl0 = make_float2(__uint_as_float(lw.x), __uint_as_float(lw.y)); nl1 = make_float2(__uint_as_float(lw.z), __uint_as_float(lw.w));
            dl0 = make_float2(-__uint_as_float(dw.x), -__uint_as_float(dw.y)); dl1 = make_float2(-__uint_as_float(dw.z), -__uint_as_float(dw.w));
          }
          const float2 x0 = fma2(make_float2(sv[c], sv[c + 1]), sl2, nl0);
          const float2 x1 = fma2(make_float2(sv[c + 2], sv[c + 3]), sl2, nl1);
          float2 p0 = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
          // a quarter of the exponentials takes the FMA-pipe polynomial: the MUFU unit (16 / clk / SM) is shared by 16 warps
          float2 p1 = ((c & 4) != 0) ? ex2_poly2(x1) : make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
          float2 e0 = fma2(make_float2(dp[c], dp[c + 1]), sc2, dl0);      // scale * dP - scale * delta
          float2 e1 = fma2(make_float2(dp[c + 2], dp[c + 3]), sc2, dl1);
          if constexpr (kRagged) {                        // P = dS = 0 beyond the sample
            if (c >= cvalid) { p0.x = 0.f; e0.x = 0.f; }
            if (c + 1 >= cvalid) { p0.y = 0.f; e0.y = 0.f; }
            if (c + 2 >= cvalid) { p1.x = 0.f; e1.x = 0.f; }
            if (c + 3 >= cvalid) { p1.y = 0.f; e1.y = 0.f; }
          }
          e0 = fma2(p0, e0, make_float2(0.f, 0.f));
          e1 = fma2(p1, e1, make_float2(0.f, 0.f));
          wp[c >> 1] = pack_bf16x2(p0.x, p0.y);
          wp[(c >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
          wd[c >> 1] = pack_bf16x2(e0.x, e0.y);
          wd[(c >> 1) + 1] = pack_bf16x2(e1.x, e1.y);
        }
      };
      if (cvalid >= 16) body(std::false_type{}); else body(std::true_type{});
      if (warp == 4) ABWD_STAMP(6);
      if constexpr (kKV) tmem_st8(tS, wp);
      tmem_st8(tD, wd);
      tmem_st_wait();
      if (warp == 4) ABWD_STAMP(7);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_pd_full + (g & 1) * 8);
      if (++stage == kAbStages) { stage = 0; cphase ^= 1; }
    }
    // epilogue: accumulators -> bf16 rows of dqkv; this warp stores columns [HD/4 * cq, HD/4 * (cq + 1)) of each accumulator
    // (the tcgen05.ld is warp-collective: only the stores are predicated)
    mbar_wait(acc_done, it & 1, 31);
    __syncwarp();
    tc_fence_after();
    ABWD_ESTAMP(0);
    constexpr int kEC = HD / 4;                         // columns per warp: 16 or 32
    float o[kKV ? 2 : 1][kEC];
#pragma unroll
    for (int a = 0; a < (kKV ? 2 : 1); ++a) {
      const uint32_t ta = tmem_base + lane_addr + (a == 0 ? Geo::kAcc1 : Geo::kAcc2) + cq * kEC;
      if constexpr (kEC == 16) tmem_ld16(ta, o[a]); else tmem_ld32(ta, o[a]);
    }
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(acc_free);               // the next item's first accumulation may overwrite the accumulators
#pragma unroll
    for (int a = 0; a < (kKV ? 2 : 1); ++a) {
      const int col = (kKV ? (a == 0 ? p.k_col : p.v_col) : p.q_col) + head * HD + cq * kEC;
      __nv_bfloat16* dst = p.dqkv + static_cast<size_t>(row_base + min(row, p.T - 1)) * p.ld + col;
      if (row < p.T) {
#pragma unroll
        for (int q = 0; q < kEC / 8; ++q) {
          const bool real = cq * kEC + q * 8 < p.hd;    // head_dim is a multiple of 8: whole 16-byte groups are real or padding
          *reinterpret_cast<uint4*>(dst + q * 8) = real ?
              make_uint4(pack_bf16x2(o[a][8 * q], o[a][8 * q + 1]), pack_bf16x2(o[a][8 * q + 2], o[a][8 * q + 3]),
                         pack_bf16x2(o[a][8 * q + 4], o[a][8 * q + 5]), pack_bf16x2(o[a][8 * q + 6], o[a][8 * q + 7])) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    ABWD_ESTAMP(1);
    rblk = rblk_n; head = head_n; b = b_n;
    }   // work items
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// Row statistics of the backward: delta[b,h,t] = scale * sum_d dO[t,h,d] * O[t,h,d] and nlse2 = -lse2 (both in the form the
// packed fma of the kernels above consumes).  One thread per (token row, head): 16-byte loads along the head, no shuffles;
// neighbouring threads read neighbouring heads of the same row (coalesced).
__global__ void attn_delta_kernel(float* __restrict__ delta, float* __restrict__ nlse2, const float* __restrict__ lse2,
                                  const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O, int B, int T, int H,
                                  int hd, float scale) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * T * H) return;
  const int h = i % H;
  const size_t row = i / H;
  const int b = row / T, t = row % T;
  const uint4* a = reinterpret_cast<const uint4*>(dO + row * H * hd + static_cast<size_t>(h) * hd);
  const uint4* o = reinterpret_cast<const uint4*>(O + row * H * hd + static_cast<size_t>(h) * hd);
  float s = 0.f;
  for (int j = 0; j < hd / 8; ++j) {         // head_dim is a multiple of 8
    const uint4 x = __ldg(a + j), y = __ldg(o + j);
    const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      s = fmaf(__uint_as_float(xw[q] << 16), __uint_as_float(yw[q] << 16), s);
      s = fmaf(__uint_as_float(xw[q] & 0xffff0000u), __uint_as_float(yw[q] & 0xffff0000u), s);
    }
  }
  const size_t k = (static_cast<size_t>(b) * H + h) * T + t;
  delta[k] = s * scale;
  nlse2[k] = -lse2[k];
}

}  // namespace ldmae
