// attn_fwd_persist_kernel: attn_fwd_kernel (attention_sm100.cuh) as a PERSISTENT kernel -- one CTA per SM walks the
// (query pair, head, sample) work items.  A non-persistent CTA lives for only eight 128-key steps at T = 1024 and spends
// ~4500 clk of its ~31000 before the first scores arrive (TMEM allocation, barrier set-up, the HBM latency of the first
// Q / K tiles, the first QK^T) plus the tail of its epilogue with the tensor core idle (clock64 trace); with
// one CTA per SM nothing else can cover that.  Here the loader warp runs ahead across item boundaries (the K / V ring and
// the next item's Q tile are in flight while the current item finishes), TMEM and barriers are set up once, and the
// scores of the next item's first block are computed under the current item's last exponentials and epilogue.
//
// Same roles, TMEM map and softmax arithmetic as attn_fwd_kernel (8 softmax warps, thread = one query row; kFixedMax = the
// constant-offset softmax for qk-normed heads).  Barrier phases are driven by running counters instead of the block index:
//   g   number of key blocks processed so far by this CTA (all items)          -> s_full / s_free / p_full / o_done / turn
//   it  number of items processed so far                                      -> q_full / q_empty / o_free
// Extra barriers: q_empty (last QK^T of an item issued -> the Q tile may be reloaded), o_free[t] (the epilogue has read O_t ->
// the next item's first P.V may overwrite it).  The output is staged in its own shared-memory tiles (the Q tiles are being
// reloaded during the epilogue).
//
// Instantiations (ldmae_lib.cu:run_attention picks one):
//   <false>              running-maximum softmax (no qk-norm, VMAE), 384 threads, in-line epilogue
//   <true>               constant-offset softmax, key blocks that may be ragged (T % 128 != 0) or a single block per item
//   <true, true>         + kPipe: software-pipelined softmax warps, four epilogue warps (512 threads), used by the training forward
//   <true, true, true>   + kRaw: exponents straight from the scores (inference forward: scale * log2e folded into q_norm.weight)
// The MMA issuer walks ONE flat loop over the CTA's key blocks (all items) in every instantiation: the scores of block g + 1
// are issued before the P.V of block g even when g + 1 is the next item's first block.
#pragma once
#include "attention_sm100.cuh"

namespace ldmae {

constexpr int kAttnPSmemBytes = kAttnSmemBytes + 2 * kAttnTileBytes     // + output staging
                                + 2 * 2 * 128 * 4 + 64;                // + row sums handed to the epilogue warps (kPipe), more barriers
constexpr int kAttnPipeThreads = 512;                                  // kPipe: + four epilogue warps

// kPipe (requires kFixedMax): the softmax warps software-pipeline their TMEM traffic under the exponentials.  Without a row
// maximum a thread needs no look at the whole block before its first exponential, so the 128 scores of a step are taken in
// four 32-column chunks: the loads of chunks 2,3 are in flight under the exponentials of chunk 0 (S_t goes back to the
// tensor core after chunk 0 instead of before it), every chunk's P columns are stored as soon as they are packed, and the
// first two chunks of the NEXT step (next key block, or the next item's first block) are fetched under the exponentials of
// chunk 3 when their scores are already there.  A warp then issues MUFU work almost without gaps, instead of
// load -> exponentials -> store phases in lockstep with the other tile's warp on the same SM sub-partition.
#ifndef LDMAE_ATTN_SCALAR_SUM
#define LDMAE_ATTN_SCALAR_SUM 1
#endif
#ifndef LDMAE_ATTN_POLY_NUM
#define LDMAE_ATTN_POLY_NUM 1
#endif
#ifndef LDMAE_ATTN_POLY_DEN
#define LDMAE_ATTN_POLY_DEN 4
#endif
// pair q (of 64 per step) takes the FMA-pipe polynomial: kNum of every kDen pairs, evenly spread
__host__ __device__ constexpr bool attn_pair_is_poly(int q) {
  return ((q * LDMAE_ATTN_POLY_NUM) % LDMAE_ATTN_POLY_DEN) < LDMAE_ATTN_POLY_NUM;
}
// one 32-column chunk of a row: exponentials (MUFU / polynomial), partial row sums, bf16 pack -> 16 packed P columns
// kRaw: the scores already are the exponents (q pre-multiplied by scale * log2e, no offset: |s| <= m0 <= 48)
// (Concentrating the polynomial pairs in alternate chunks, swapped between the two tiles so that one warp of a sub-partition
// loads the MUFU unit while the other loads the FMA pipe, measured 796-855 vs 876 TFLOP/s: the even spread lets each warp
// overlap its own MUFU and FMA work, which matters more.)
template <int kChunk, bool kRaw = false>
__device__ __forceinline__ void attn_exp_chunk(const float* s, const float2 sc2, const float2 neg2, float2& ls0, float2& ls1,
                                               uint32_t* wv) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float2 x0 = kRaw ? make_float2(s[i], s[i + 1]) : fma2(make_float2(s[i], s[i + 1]), sc2, neg2);
    const float2 x1 = kRaw ? make_float2(s[i + 2], s[i + 3]) : fma2(make_float2(s[i + 2], s[i + 3]), sc2, neg2);
    const float2 p0 = attn_pair_is_poly(kChunk * 16 + i / 2) ? ex2_poly2_bounded(x0) : make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
    const float2 p1 = attn_pair_is_poly(kChunk * 16 + i / 2 + 1) ? ex2_poly2_bounded(x1) : make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
#if LDMAE_ATTN_SCALAR_SUM
    // scalar FADDs (either FMA pipe) instead of add.f32x2: 876 -> 899 TFLOP/s together with the 1/4 polynomial share
    ls0.x += p0.x; ls0.y += p0.y; ls1.x += p1.x; ls1.y += p1.y;
#else
    ls0 = add2(ls0, p0);
    ls1 = add2(ls1, p1);
#endif
    wv[i >> 1] = pack_bf16x2(p0.x, p0.y);
    wv[(i >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
  }
}

#ifdef LDMAE_ATTN_TRACE
#define ATTN_PSTAMP(k, v) do { if (p.trace && blockIdx.x == 0 && lane == 0 && wq == 0 && g < 48) p.trace[((t * 64 + g) * 8) + (k)] = (v); } while (0)
#define ATTN_ESTAMP(k) do { if (p.trace && blockIdx.x == 0 && lane == 0 && wq == 0 && it < 16) p.trace[((t * 64 + 48 + it) * 8) + (k)] = clock64(); } while (0)
#else
#define ATTN_PSTAMP(k, v) do { } while (0)
#define ATTN_ESTAMP(k) do { } while (0)
#endif

template <bool kFixedMax, bool kPipe = false, bool kRaw = false>
__global__ void __launch_bounds__(kPipe ? kAttnPipeThreads : kAttnThreads, 1)
attn_fwd_persist_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out, const AttnParams p,
                        const int n_qpairs, const int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * kAttnTileBytes;
  uint8_t* sV = sK + kAttnKVStages * kAttnTileBytes;
  uint8_t* sO = sV + kAttnKVStages * kAttnTileBytes;      // [2 tiles][4 warps] x 4 KB output staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * kAttnTileBytes);
  uint64_t* q_full = bars;                         // [1]
  uint64_t* q_empty = bars + 1;                    // [1]
  uint64_t* k_full = bars + 2;                     // [stages]
  uint64_t* k_empty = k_full + kAttnKVStages;
  uint64_t* v_full = k_empty + kAttnKVStages;
  uint64_t* v_empty = v_full + kAttnKVStages;
  uint64_t* s_full = v_empty + kAttnKVStages;      // [2]
  uint64_t* s_free = s_full + 2;                   // [2]
  uint64_t* p_full = s_free + 2;                   // [2]
  uint64_t* o_done = p_full + 2;                   // [2]
  uint64_t* o_free = o_done + 2;                   // [2]
  uint64_t* turn = o_free + 2;                     // [2]
  uint64_t* l_full = turn + 2;                     // [2]  kPipe: the row sums of an item are in shared memory (softmax -> epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(l_full + 2);
  float* l_s = reinterpret_cast<float*>(tmem_slot + 2);     // kPipe: [2 item parities][2 tiles][128 rows]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkv = (p.T + 127) / 128;
  // item w -> (qpair fastest: neighbouring CTAs share the K / V of one (sample, head) in L2)
  auto item_coords = [&](int w, int& qpair, int& head, int& b) {
    qpair = w % n_qpairs;
    head = (w / n_qpairs) % p.H;
    b = w / (n_qpairs * p.H);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int s = 0; s < kAttnKVStages; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1); mbar_init(&s_free[t], 4); mbar_init(&p_full[t], 4); mbar_init(&o_done[t], 1);
      mbar_init(&o_free[t], 4); mbar_init(&turn[t], 4); mbar_init(&l_full[t], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Register split (setmaxnreg moves registers only through the pool the CTA's own warps release, never the part the launch
  // left unallocated): 384 threads x 168 -> control 80 / softmax 208;  kPipe: 512 x 128 -> control 64 / epilogue 48 / softmax 200
  if (warp < 4) {
  if constexpr (kPipe) setmaxnreg_dec<64>(); else setmaxnreg_dec<80>();
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        int qpair, head, b;
        item_coords(w, qpair, head, b);
        const int row_base = b * p.T;
        for (int j = 0; j < nkv; ++j) {
          mbar_wait(&k_empty[stage], phase ^ 1, 10);
          mbar_expect_tx(&k_full[stage], kAttnTileBytes);
          tma_load_2d(&tmap_qkv, &k_full[stage], sK + stage * kAttnTileBytes, p.k_col + head * 64, row_base + j * 128);
          if (j == 0) {
            // the Q tiles of this item, once the previous item's last QK^T has read them
            if (it > 0) mbar_wait(q_empty, (it - 1) & 1, 12);
            mbar_expect_tx(q_full, 2 * kAttnTileBytes);
            tma_load_2d(&tmap_qkv, q_full, sQ, p.q_col + head * 64, row_base + qpair * 256);
            tma_load_2d(&tmap_qkv, q_full, sQ + kAttnTileBytes, p.q_col + head * 64, row_base + qpair * 256 + 128);
          }
          mbar_wait(&v_empty[stage], phase ^ 1, 11);
          mbar_expect_tx(&v_full[stage], kAttnTileBytes);
          tma_load_2d(&tmap_qkv, &v_full[stage], sV + stage * kAttnTileBytes, p.v_col + head * 64, row_base + j * 128);
          if (++stage == kAttnKVStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, false, true);   // V is MN-major (d contiguous)
    const bool issuer = elect_one();
    const uint64_t d_q0 = umma_smem_desc_sw128(smem_u32(sQ), 1024, 0);
    const uint64_t d_k0 = umma_smem_desc_sw128(smem_u32(sK), 1024, 0);
    const uint64_t d_v0 = umma_smem_desc_sw128(smem_u32(sV), 1024, kAttnTileBytes);
    auto issue_qk = [&](int t, int kstage) {
      const uint64_t dq = d_q0 + static_cast<uint64_t>(t * (kAttnTileBytes >> 4));
      const uint64_t dk = d_k0 + static_cast<uint64_t>(kstage * (kAttnTileBytes >> 4));
      const uint32_t td = tmem_base + t * 128;
      if (issuer) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16<1>(td, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        umma_commit<1>(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int vstage, uint32_t accumulate) {
      const uint64_t dv = d_v0 + static_cast<uint64_t>(vstage * (kAttnTileBytes >> 4));
      const uint32_t td = tmem_base + 384 + t * 64, ta = tmem_base + 256 + t * 64;
      if (issuer) {
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 8 x 16 keys; P: 8 TMEM columns (16 bf16) per step
          umma_bf16_ts(td, ta + k * 8, dv + 128 * k, idesc_pv, k != 0 ? 1u : accumulate);
        umma_commit<1>(&o_done[t]);
      }
      __syncwarp();
    };
    int stage = 0; uint32_t phase = 0;                 // ring position of the block whose P.V comes next
    // One flat loop over the CTA's key blocks (all items): the scores of block g + 1 are issued before the P.V of block g even
    // when g + 1 is the NEXT item's first block -- issued after the item's last P.V they reached the softmax warps ~1500 clk
    // late once the epilogue had moved to its own warps (clock64 trace).
    const int my_items = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int total = my_items * nkv;
    if (total > 0) {
      mbar_wait(q_full, 0, 20);
      mbar_wait(&k_full[0], 0, 21);
      for (int t = 0; t < 2; ++t) {
        tc_fence_after();
        issue_qk(t, 0);
      }
      if (issuer) {
        umma_commit<1>(&k_empty[0]);
        if (nkv == 1) umma_commit<1>(q_empty);
      }
      __syncwarp();
    }
    int it = 0, j = 0;                                 // item / key block of step g
    for (int g = 0; g < total; ++g) {
      int nstage = stage + 1; uint32_t nphase = phase;
      if (nstage == kAttnKVStages) { nstage = 0; nphase ^= 1; }
      if (g + 1 < total) {
        // refill S_t with the next block as soon as the softmax warps hold block g in registers
        const int jn = (j + 1 == nkv) ? 0 : j + 1;
        if (jn == 0) mbar_wait(q_full, (it + 1) & 1, 20);     // the next item's Q tiles
        mbar_wait(&k_full[nstage], nphase, 25);
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&s_free[t], g & 1, 26 + t);
          tc_fence_after();
          issue_qk(t, nstage);
        }
        if (issuer) {
          umma_commit<1>(&k_empty[nstage]);
          if (jn == nkv - 1) umma_commit<1>(q_empty);         // that was an item's last QK^T: the Q tiles may be reloaded
        }
        __syncwarp();
      }
      mbar_wait(&v_full[stage], phase, 24);
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_full[t], g & 1, 22 + t);
        // the first P.V of an item overwrites O_t: the previous item's epilogue must have read it
        if (j == 0 && it > 0) mbar_wait(&o_free[t], (it - 1) & 1, 28 + t);
        tc_fence_after();
        issue_pv(t, stage, j == 0 ? 0u : 1u);
      }
      if (issuer) umma_commit<1>(&v_empty[stage]);
      __syncwarp();
      stage = nstage; phase = nphase;
      if (++j == nkv) { j = 0; ++it; }
    }
  }
  } else if (kPipe && warp >= 12) {
    // ===================== kPipe: epilogue warps (O_t / l -> bf16 -> global), off the softmax warps' critical path =====================
    // The softmax warps of the eight-warp kernel spent ~3000 of an item's ~22000 clk normalising and storing O with the
    // tensor core idle (clock64 trace, tools/trace_attn_pipe.py); here they hand the row sums over and go straight on to the
    // next item's scores, which are already in TMEM.  Warp e serves TMEM lanes 32 (e & 3) .. +31 of tile 0, then of tile 1.
    setmaxnreg_dec<48>();
    const int wq = warp & 3;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      int qpair, head, b;
      item_coords(w, qpair, head, b);
      const int row_base = b * p.T;
      const int g_last = (it + 1) * nkv - 1;               // the item's last key block in the CTA's running block count
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const uint32_t tO = tmem_base + lane_addr + 384 + t * 64;
        uint8_t* stage_o = sO + t * kAttnTileBytes + wq * 4096;
        mbar_wait_sleep(&l_full[t], it & 1, 200, 40 + t);
        const float l_run = l_s[((it & 1) * 2 + t) * 128 + wq * 32 + lane];
        const float inv_l = 1.f / l_run;
        const int q_tok0 = qpair * 256 + t * 128 + wq * 32;
        if (p.lse2 != nullptr && q_tok0 + lane < p.T)
          p.lse2[(static_cast<size_t>(b) * p.H + head) * p.T + q_tok0 + lane] = (kRaw ? 0.f : p.m0_log2) + log2f(l_run);
        // l_full(it) implies o_done(g_last - 1) has completed (the softmax warps waited for it before their last P store),
        // so the parity wait below cannot be satisfied by an older phase
        mbar_wait(&o_done[t], g_last & 1, 42 + t);
        __syncwarp();
        tc_fence_after();
        if (lane == 0) tma_store_wait_read<1>();            // this staging tile's previous store (two groups back) has been read
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float o[32];
          tmem_ld32(tO + c * 32, o);
          tmem_ld_wait();
          tmem_ld_pin32(o);
          if (c == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[t]);        // O_t is in registers: the next item's first P.V may overwrite it
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 v = make_uint4(pack_bf16x2(o[8 * q] * inv_l, o[8 * q + 1] * inv_l), pack_bf16x2(o[8 * q + 2] * inv_l, o[8 * q + 3] * inv_l),
                                       pack_bf16x2(o[8 * q + 4] * inv_l, o[8 * q + 5] * inv_l), pack_bf16x2(o[8 * q + 6] * inv_l, o[8 * q + 7] * inv_l));
            sts128(stage_o + lane * 128 + (((c * 4 + q) ^ (lane & 7)) << 4), v);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_out, stage_o, head * 64, row_base + q_tok0);   // (T % 128 == 0: every 32-row tile is whole)
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
  } else {
    if constexpr (kPipe) setmaxnreg_inc<200>(); else setmaxnreg_inc<208>();
    // ===================== softmax: thread = one query row =====================
    const int t = (warp - 4) >> 2;                       // tile 0 / 1
    const int wq = warp & 3;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + t * 128;
    const uint32_t tP = tmem_base + lane_addr + 256 + t * 64;
    const uint32_t tO = tmem_base + lane_addr + 384 + t * 64;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    uint8_t* stage_o = sO + t * kAttnTileBytes + wq * 4096;
    int g = 0, it = 0;
    // kPipe: shared addresses of this tile's barriers, computed once (a generic -> shared conversion per barrier operation
    // was ~6 % of the loop's issue slots), and the four score chunks of a step (sA, sB live across steps)
    uint32_t a_s_full = smem_u32(&s_full[t]), a_s_free = smem_u32(&s_free[t]), a_p_full = smem_u32(&p_full[t]),
             a_o_done = smem_u32(&o_done[t]);
    asm volatile("" : "+r"(a_s_full), "+r"(a_s_free), "+r"(a_p_full), "+r"(a_o_done));   // opaque: keep them in registers
    float sA[32], sB[32], sC[32], sD[32];
    bool pre = false;                                    // kPipe: sA, sB already hold the next step's first two chunks
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      int qpair, head, b;
      item_coords(w, qpair, head, b);
      const int row_base = b * p.T;
      float m_used = kRaw ? 0.f : (kFixedMax ? p.m0_log2 / p.scale_log2 : -INFINITY);
      float l_run = 0.f;
      if constexpr (kPipe) {
        static_assert(kFixedMax, "the pipelined softmax needs the constant-offset instantiation");
        static_assert(kPipe || !kRaw, "raw exponents exist only in the pipelined instantiation");
        const float2 neg2 = make_float2(-p.m0_log2, -p.m0_log2);
        // (the host selects this instantiation only for T % 128 == 0: no masked keys)
        float2 ls0 = make_float2(0.f, 0.f), ls1 = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int j = 0; j < nkv; ++j, ++g) {
          bool s_released;
          ATTN_PSTAMP(0, clock64());
          ATTN_PSTAMP(7, pre ? 1 : 0);
          if (!pre) {
            // nothing fetched ahead (first step of the CTA, or the scores were late): the whole row, then release S_t
            mbar_wait_quiet_a(a_s_full, g & 1);
            __syncwarp();
            tc_fence_after();
            tmem_ld32(tS, sA);
            tmem_ld32(tS + 32, sB);
            tmem_ld32(tS + 64, sC);
            tmem_ld32(tS + 96, sD);
            tmem_ld_wait();
            tmem_ld_pin32(sA); tmem_ld_pin32(sB); tmem_ld_pin32(sC); tmem_ld_pin32(sD);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(a_s_free);
            s_released = true;
          } else {
            // chunks 0,1 arrived under the previous step's last exponentials; 2,3 load under chunk 0's
            tmem_ld32(tS + 64, sC);
            tmem_ld32(tS + 96, sD);
            s_released = false;
          }
          ATTN_PSTAMP(1, clock64());
          uint32_t wv0[16], wv[16];
          attn_exp_chunk<0, kRaw>(sA, sc2, neg2, ls0, ls1, wv0);
          ATTN_PSTAMP(2, clock64());
          if (!s_released) {
            tmem_ld_wait();
            tmem_ld_pin32(sC); tmem_ld_pin32(sD);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(a_s_free);          // the tensor core may overwrite S_t with the next block
          }
          ATTN_PSTAMP(3, clock64());
          attn_exp_chunk<1, kRaw>(sB, sc2, neg2, ls0, ls1, wv);
          // P_t may only be overwritten after the previous P_t V has been read (the previous block of this CTA, whichever
          // item): that product was issued at the end of the previous step, so the first store waits until half of this
          // step's exponentials are done (the first chunk's probabilities sit in registers meanwhile)
          if (g > 0) mbar_wait_quiet_a(a_o_done, (g - 1) & 1);
          __syncwarp();
          tc_fence_after();
          tmem_st16(tP, wv0);
          tmem_st16(tP + 16, wv);
          attn_exp_chunk<2, kRaw>(sC, sc2, neg2, ls0, ls1, wv);
          tmem_st16(tP + 32, wv);
          // the next step's scores (next key block / next item's first block), if the tensor core has delivered them
          ATTN_PSTAMP(4, clock64());
          pre = __all_sync(0xffffffffu, mbar_probe_a(a_s_full, (g + 1) & 1) != 0);
          if (pre) {
            tc_fence_after();
            tmem_ld32(tS, sA);
            tmem_ld32(tS + 32, sB);
          }
          ATTN_PSTAMP(5, clock64());
          attn_exp_chunk<3, kRaw>(sD, sc2, neg2, ls0, ls1, wv);
          tmem_st16(tP + 48, wv);
          if (pre) {
            tmem_ld_wait();
            tmem_ld_pin32(sA); tmem_ld_pin32(sB);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(a_p_full);
          ATTN_PSTAMP(6, clock64());
        }
        // hand the row sums to the epilogue warps and go on to the next item
        l_s[((it & 1) * 2 + t) * 128 + wq * 32 + lane] = (ls0.x + ls0.y) + (ls1.x + ls1.y);
        __syncwarp();
        if (lane == 0) mbar_arrive(&l_full[t]);
        // (kPipe has no in-line epilogue: the guard below)
      } else {
#pragma unroll 1
      for (int j = 0; j < nkv; ++j, ++g) {
        mbar_wait(&s_full[t], g & 1, 30 + t);
        __syncwarp();
        tc_fence_after();
        float s[128];
        tmem_ld32(tS, s);
        tmem_ld32(tS + 32, s + 32);
        tmem_ld32(tS + 64, s + 64);
        tmem_ld32(tS + 96, s + 96);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);            // the tensor core may overwrite S_t with the next block
        const int kvalid = p.T - j * 128;
        if (kvalid < 128) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= kvalid) s[i] = -INFINITY;
        }
        if constexpr (!kFixedMax) {
          float mx[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) mx[i] = s[i];
#pragma unroll
          for (int i = 8; i < 128; i += 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) mx[q] = fmaxf(mx[q], s[i + q]);
          }
          const float m_blk = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
          if (j == 0) {
            m_used = m_blk;
          } else {
            const bool grow = (m_blk - m_used) * p.scale_log2 > kAttnRescaleLog2;
            if (__any_sync(0xffffffffu, grow)) {
              const float m_new = fmaxf(m_used, m_blk);
              const float alpha = ex2_approx((m_used - m_new) * p.scale_log2);
              mbar_wait(&o_done[t], (g - 1) & 1, 32 + t);
              __syncwarp();
              tc_fence_after();
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                float o[32];
                tmem_ld32(tO + c * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] *= alpha;
                tmem_st32(tO + c * 32, reinterpret_cast<const uint32_t*>(o));
              }
              tmem_st_wait();
              l_run *= alpha;
              m_used = m_new;
            }
          }
        }
        // P_t may only be overwritten after the previous P_t V has been read (the previous block of this CTA, whichever item)
        const uint32_t p_free = (g > 0) ? mbar_probe(&o_done[t], (g - 1) & 1) : 1u;
        const float neg = kFixedMax ? -p.m0_log2 : -m_used * p.scale_log2;
        const float2 neg2 = make_float2(neg, neg);
        if (p.alternate) {
          if (t == 1) mbar_wait(&turn[0], g & 1, 38);
          else if (g > 0) mbar_wait(&turn[1], (g - 1) & 1, 39);
        }
        float2 ls0 = make_float2(0.f, 0.f), ls1 = make_float2(0.f, 0.f);
        uint32_t wv[64];
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          const float2 x0 = fma2(make_float2(s[i], s[i + 1]), sc2, neg2);
          const float2 x1 = fma2(make_float2(s[i + 2], s[i + 3]), sc2, neg2);
          const float2 p0 = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
          const float2 p1 = ((i / 4) % kAttnPolyEvery == kAttnPolyEvery - 1) ? ex2_poly2(x1) : make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
          ls0 = add2(ls0, p0);
          ls1 = add2(ls1, p1);
          wv[i >> 1] = pack_bf16x2(p0.x, p0.y);
          wv[(i >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
        }
        l_run += (ls0.x + ls0.y) + (ls1.x + ls1.y);
        if (p.alternate) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&turn[t]);
        }
        if (!p_free) mbar_wait(&o_done[t], (g - 1) & 1, 36 + t);
        __syncwarp();
        tc_fence_after();
        tmem_st32(tP, wv);
        tmem_st32(tP + 32, wv + 32);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      }
      if constexpr (!kPipe) {
      // epilogue of the item: O_t / l -> bf16 -> global (own staging tile + one TMA store per warp)
      ATTN_ESTAMP(0);
      mbar_wait(&o_done[t], (g - 1) & 1, 34 + t);
      ATTN_ESTAMP(1);
      __syncwarp();
      tc_fence_after();
      const float inv_l = 1.f / l_run;
      const int q_tok0 = qpair * 256 + t * 128 + wq * 32;
      if (p.lse2 != nullptr && q_tok0 + lane < p.T)
        p.lse2[(static_cast<size_t>(b) * p.H + head) * p.T + q_tok0 + lane] = fmaf(m_used, p.scale_log2, log2f(l_run));
      const bool whole = q_tok0 + 32 <= p.T;
      __nv_bfloat16* dst = p.out + static_cast<size_t>(row_base + q_tok0 + lane) * p.ldo + head * 64;
      if (lane == 0) tma_store_wait_read<0>();              // the previous item's store has read this staging tile
      __syncwarp();
      ATTN_ESTAMP(2);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float o[32];
        tmem_ld32(tO + c * 32, o);
        tmem_ld_wait();
        tmem_ld_pin32(o);
        if (c == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[t]);          // O_t is in registers: the next item's first P.V may overwrite it
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 v = make_uint4(pack_bf16x2(o[8 * q] * inv_l, o[8 * q + 1] * inv_l), pack_bf16x2(o[8 * q + 2] * inv_l, o[8 * q + 3] * inv_l),
                                     pack_bf16x2(o[8 * q + 4] * inv_l, o[8 * q + 5] * inv_l), pack_bf16x2(o[8 * q + 6] * inv_l, o[8 * q + 7] * inv_l));
          if (whole) sts128(stage_o + lane * 128 + (((c * 4 + q) ^ (lane & 7)) << 4), v);
          else if (q_tok0 + lane < p.T) *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = v;
        }
      }
      ATTN_ESTAMP(3);
      if (whole) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_out, stage_o, head * 64, row_base + q_tok0);
          tma_store_commit();
        }
      }
      ATTN_ESTAMP(4);
      }   // !kPipe
    }
    if constexpr (!kPipe) { if (lane == 0) tma_store_wait_read<0>(); }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace ldmae
