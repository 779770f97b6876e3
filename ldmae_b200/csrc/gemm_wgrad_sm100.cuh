// Weight-gradient GEMM for sm_100a:   C[N1,N2] (fp32) += sum_m P[m,N1] * Q[m,N2]
// (P = gradient of a Linear's output, Q = the Linear's input; both row-major bf16 with the contraction index m as the
// SLOW axis, exactly as the forward/backward kernels leave them in HBM -- no transposed copies exist.)
//
// Both operands are therefore "MN-major" for the tensor core: a TMA box of {64 columns x 64 rows(m)} lands in shared
// memory as 64 m-rows of 128 swizzled bytes, which is the canonical SWIZZLE_128B MN-major atom (8 m-rows x 64 elements)
// repeated along m (stride SBO = 1024 B); 64-column atoms along the MN dimension are separate boxes (stride LBO = 8 KB).
// tcgen05.mma reads them with the a/b "major" bits of the instruction descriptor set.
//
// The output is small (a weight matrix) and the contraction is huge (all tokens of the batch), so the work is split
// along m: unit = (output tile, m-range); every unit adds its partial tile into C with TMA reduce-add
// (cp.reduce.async.bulk.tensor .add on an fp32 tensor map), which is also the gradient-accumulation semantic
// (`.grad +=`) of the training loop.  Same warp roles, CTA pairs (cta_group::2, 256 x BN tile) and TMEM double
// buffering as gemm_tn_kernel.
#pragma once
#include "gemm_sm100.cuh"

namespace ldmae {

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

// fp32 accumulate-into-C epilogue (optionally scaled): staging and ring as EpiStore<float>
struct EpiAccum : StoreRing {
  struct Params {
    CUtensorMap cmap;    // C [N1, N2] fp32: box {32, 32}, SWIZZLE_128B
    float alpha;         // partial tile is multiplied by alpha before it is added
  };
  static __device__ __forceinline__ void prefetch_maps(const Params& p) { tma_prefetch_desc(&p.cmap); }
  static __device__ __forceinline__ void run(const Params& p, int N2, const EpiCtx& c, State& st, uint32_t acc, int row0, int n0,
                                             int cbase, int gc) {
#pragma unroll 1
    for (int c0 = cbase; c0 < cbase + gc; c0 += 32) {
      if (n0 + c0 >= N2) break;
      uint8_t* tile = acquire(c, st);
      float v[32];
      tmem_ld32(acc + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        st_tile16(sw128_chunk(tile, c.lane, q),
                  make_uint4(__float_as_uint(v[4 * q] * p.alpha), __float_as_uint(v[4 * q + 1] * p.alpha),
                             __float_as_uint(v[4 * q + 2] * p.alpha), __float_as_uint(v[4 * q + 3] * p.alpha)));
      fence_proxy_async_smem();
      __syncwarp();
      if (c.el) {
        tma_reduce_add_2d(&p.cmap, tile, n0 + c0, row0);
        tma_store_commit();
      }
      ++st.seq;
    }
  }
};

struct WgradShape {
  int N1, N2, M;         // C is [N1, N2]; contraction length M
  int splits, kb_per_split;
};

template <int BN, int CG>
struct WgradCfg {
  static constexpr int kLoadBN = BN / CG;
  static constexpr int kABytes = kBM * kBK * 2;              // 128 columns of P x 64 m
  static constexpr int kBBytes = kLoadBN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiWarps = 8;
  static constexpr int kGroups = BN >= 256 ? 2 : 1;
  static constexpr int kGroupCols = BN / kGroups;
  static constexpr int kEpiBytes = kEpiWarps * EpiAccum::kWarpBytes;
  static constexpr int kBarBytes = 1024;
  static constexpr int kStagesRaw = (kSmemLimit - kEpiBytes - kBarBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BN <= 256 ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes;
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
  static_assert(kStages >= 3, "smem ring too shallow");
  static_assert(kLoadBN % 64 == 0 && BN <= 256, "operand boxes are 64 columns wide");
};

template <int BN, int CG>
__global__ void __launch_bounds__(WgradCfg<BN, CG>::kThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_q, const WgradShape g,
                  const __grid_constant__ EpiAccum::Params ep) {
  using Cfg = WgradCfg<BN, CG>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* smem_epi = smem + kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::kEpiBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint64_t* tempty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_p);
    tma_prefetch_desc(&tmap_q);
    EpiAccum::prefetch_maps(ep);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], CG);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], CG * 4 * Cfg::kGroups);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<CG>(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n2_tiles = (g.N2 + BN - 1) / BN;
  const int n1_tiles = (g.N1 + kBM * CG - 1) / (kBM * CG);
  const int tiles = n1_tiles * n2_tiles;
  const int total_units = tiles * g.splits;
  const int num_kb = (g.M + kBK - 1) / kBK;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;

  if (warp == 0) {
    // whole-warp control flow, elected issue (see gemm_sm100.cuh: a single-lane branch makes nvcc wrap every TMA / MMA
    // instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY convergence loop)
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int u = cluster_id; u < total_units; u += num_clusters) {
        const int tile = u % tiles, split = u / tiles;
        const int col_p = ((tile / n2_tiles) * CG + static_cast<int>(cta_rank)) * kBM;
        const int col_q = (tile % n2_tiles) * BN + static_cast<int>(cta_rank) * Cfg::kLoadBN;
        const int kb0 = split * g.kb_per_split, kb1 = min(num_kb, kb0 + g.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
          uint8_t* da = smem_a + stage * Cfg::kABytes;
          uint8_t* db = smem_b + stage * Cfg::kBBytes;
          if (issuer) {
          if constexpr (CG == 1) {
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
            for (int a = 0; a < kBM / 64; ++a) tma_load_2d(&tmap_p, &full_bar[stage], da + a * 8192, col_p + a * 64, kb * kBK);
#pragma unroll
            for (int b = 0; b < Cfg::kLoadBN / 64; ++b) tma_load_2d(&tmap_q, &full_bar[stage], db + b * 8192, col_q + b * 64, kb * kBK);
          } else {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            else mbar_arrive_cluster(&full_bar[stage], 0);
#pragma unroll
            for (int a = 0; a < kBM / 64; ++a) tma_load_2d_pair(&tmap_p, &full_bar[stage], da + a * 8192, col_p + a * 64, kb * kBK);
#pragma unroll
            for (int b = 0; b < Cfg::kLoadBN / 64; ++b) tma_load_2d_pair(&tmap_q, &full_bar[stage], db + b * 8192, col_q + b * 64, kb * kBK);
          }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      for (int s = 0; s < kStages; ++s) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 150 + stage);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const bool issuer = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(kBM * CG, BN, true, true);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = cluster_id; u < total_units; u += num_clusters, ++it) {
        const int split = u / tiles;
        const int kb0 = split * g.kb_per_split, kb1 = min(num_kb, kb0 + g.kb_per_split);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1, 200 + as);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase, 300 + stage);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // 16 m-rows per instruction = two 8-row core groups (SBO 1024 B apart); 64-column atoms 8 KB apart (LBO)
              const uint64_t ad = umma_smem_desc_sw128(a_addr + k * 2048, 1024, 8192);
              const uint64_t bd = umma_smem_desc_sw128(b_addr + k * 2048, 1024, 8192);
              umma_bf16<CG>(tmem_d, ad, bd, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            }
            umma_commit<CG>(&empty_bar[stage]);
            if (kb == kb1 - 1) umma_commit<CG>(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && (warp - 4) / 4 < Cfg::kGroups) {
    const int wq = warp & 3;
    const int grp = (warp - 4) / 4;
    const int ew = warp - 4;
    const EpiCtx ctx{smem_epi + ew * EpiAccum::kWarpBytes, nullptr, nullptr, lane, elect_one()};
    EpiAccum::State est;
    est.seq = 0;
    int it = 0;
    for (int u = cluster_id; u < total_units; u += num_clusters, ++it) {
      const int tile = u % tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = ((tile / n2_tiles) * CG + static_cast<int>(cta_rank)) * kBM + wq * 32;
      const int n0 = (tile % n2_tiles) * BN;
      mbar_wait(&tfull_bar[as], aphase, 400 + as);
      __syncwarp();
      tc_fence_after();
      const uint32_t acc = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * BN;
      if (row0 < g.N1) EpiAccum::run(ep, g.N2, ctx, est, acc, row0, n0, grp * Cfg::kGroupCols, Cfg::kGroupCols);
      tc_fence_before();
      __syncwarp();
      if (ctx.el) {
        if constexpr (CG == 1) mbar_arrive(&tempty_bar[as]);
        else mbar_arrive_cluster(&tempty_bar[as], 0);
      }
    }
    if (ctx.el) tma_store_wait_read<0>();
  }

  __syncwarp();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, Cfg::kTmemCols);
}

}  // namespace ldmae
