// Persistent, warp-specialised tcgen05 GEMM for sm_100a:   C[M,N] = A[M,K] . W[N,K]^T
// (A = activations, row-major bf16; W = nn.Linear weight, row-major bf16 => both operands K-major).
//
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (128B swizzle) into a kStages-deep smem ring
//   warp 1      MMA issuer:   one elected thread issues tcgen05.mma (kind::f16, fp32 accumulate in TMEM)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue:     thread = one accumulator row (tcgen05.ld 32x32b), fused math in registers, results
//                             staged in warp-private 128B-swizzled smem tiles and written with TMA stores
//                             (operands the epilogue must read, i.e. the fp32 residual stream, arrive by TMA too)
//
// kCG = 2 pairs two CTAs (cta_group::2, UMMA M = 256): each CTA loads its own 128 rows of A and half of
// the W tile, the leader CTA issues the MMAs for both, the accumulator rows of each CTA live in its own
// TMEM.  Two accumulator stages in TMEM overlap the epilogue of tile i with the main loop of tile i+1.
//
// The fused epilogues implement the LightningDiT block algebra (reference models/lightningdit.py:239-250):
// see struct comments below.  With a row per thread, a 32-row x 128-byte staging tile in the TMA 128B-swizzle
// layout is written / read with conflict-free 16-byte shared-memory accesses (chunk index XOR row%8), and all
// global traffic of the epilogue is full-line TMA bulk traffic.
#pragma once
#include "ptx.cuh"

namespace ldmae {

constexpr int kBM = 128;        // rows per CTA
constexpr int kBK = 64;         // K per stage = one 128-byte swizzle atom of bf16
// CTA = 4 control warps (TMA, MMA, TMEM alloc, spare) + Epi::kWarps epilogue warps (4 or 8; with 8 the two groups of
// four split the tile's columns, which doubles the issue slots and the latency hiding of math-heavy epilogues)
template <class Epi>
constexpr int gemm_threads() { return 128 + 32 * Epi::kWarps; }
constexpr int kSmemLimit = 232448;   // 227 KB opt-in shared memory per CTA

struct GemmShape {
  int M, N, K;
  // > 0: only the first n_live of the BN / CG weight rows each CTA owns in a tile are loaded and multiplied (the rest are known
  // to be zero: head slots wider than the head).  The accumulator then holds CG * n_live columns: CTA r's rows at
  // [r * n_live, (r + 1) * n_live).  Only epilogues that know about it (EpiQKVWide) may be launched this way.
  int n_live = 0;
};

// which tiles this CTA works on, and where this epilogue warp's rows are
struct TileSched {
  int first, stride, total, n_tiles, cg, cta_rank, wq;
  template <int BN>
  __device__ __forceinline__ void coords(int tile, int& row0, int& n0) const {
    row0 = ((tile / n_tiles) * cg + cta_rank) * kBM + wq * 32;
    n0 = (tile % n_tiles) * BN;
  }
};
struct EpiCtx {
  uint8_t* smem;     // this warp's staging area (1024-byte aligned, Epi::kWarpBytes)
  uint8_t* cta;      // CTA-wide epilogue area (Epi::kCtaBytes), filled by Epi::cta_init before the role split
  uint64_t* bars;    // this warp's mbarriers (Epi::kBars)
  int lane;
  bool el;           // this lane is the warp's elected issuer of TMA / bulk-group instructions (same lane for the whole kernel:
                     // bulk async-groups are per thread).  Every lane runs the control flow; only the instruction is predicated.
};

template <int BN, int CG, class Epi>
struct GemmCfg {
  static constexpr int kLoadBN = BN / CG;                     // W rows loaded by each CTA
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = kLoadBN * kBK * 2;
  static constexpr int kBBytesPadded = (kBBytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kABytes + kBBytesPadded;
  static constexpr int kGroups = (Epi::kWarps == 8 && BN >= 256) ? 2 : 1;   // column groups of epilogue warps at work
  static constexpr int kGroupCols = BN / kGroups;
  static constexpr int kEpiBytes = Epi::kWarps * Epi::kWarpBytes + Epi::kCtaBytes;
  static constexpr int kBarBytes = 1024;
  static constexpr int kBudget = kSmemLimit - kEpiBytes - 1024 /*align slack*/ - kBarBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kEpiBytes + kBarBytes;
  static_assert(kStages >= 3, "smem ring too shallow");
  static_assert((2 * kStages + 4 + Epi::kWarps * Epi::kBars) * 8 + 8 <= kBarBytes, "barrier area");
  static_assert(Epi::kWarpBytes % 1024 == 0 && Epi::kCtaBytes % 1024 == 0, "staging areas must keep the 1024-byte swizzle alignment");
  static_assert(2 * BN <= 512, "two accumulator stages must fit TMEM");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
  static_assert(kLoadBN % 8 == 0, "W rows per CTA");
};

// ---------------------------------------------------------------------------------------------
// Epilogue building blocks
// ---------------------------------------------------------------------------------------------
// 4 consecutive entries of a per-column vector (bias, gate, ...); columns >= N read as 0.
__device__ __forceinline__ float4 ldvec4(const float* __restrict__ base, int col, int N) {
  const float* p = base + col;
  if (col + 3 < N && (reinterpret_cast<uintptr_t>(p) & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p));
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N) r.x = __ldg(p);
  if (col + 1 < N) r.y = __ldg(p + 1);
  if (col + 2 < N) r.z = __ldg(p + 2);
  if (col + 3 < N) r.w = __ldg(p + 3);
  return r;
}
// 16-byte chunk `j` (0..7) of row `r` inside a [32 x 128 B] tile in TMA SWIZZLE_128B layout (tile 1024-B aligned)
__device__ __forceinline__ uint8_t* sw128_chunk(uint8_t* tile, int r, int j) { return tile + r * 128 + ((j ^ (r & 7)) << 4); }
// 16-byte chunk `j` (0..3) of row `r` inside a [32 x 64 B] tile in TMA SWIZZLE_64B layout (tile 512-B aligned)
__device__ __forceinline__ uint8_t* sw64_chunk(uint8_t* tile, int r, int j) { return tile + r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }

// rsqrt(mean(x^2) + eps) of a residual-stream row from its per-tile partial sums (fixed summation order)
__device__ __forceinline__ float row_rinv(const float* ssq, int row, int slots, float inv_D, float eps) {
  if (ssq == nullptr) return 1.f;
  float s = 0.f;
  for (int j = 0; j < slots; ++j) s += __ldg(ssq + static_cast<size_t>(row) * slots + j);
  return rsqrtf(s * inv_D + eps);
}

// Per-column vectors (bias, gate, cvec ...) are staged once per tile in this warp's shared memory and read back as
// broadcasts: with the whole shared-memory carve-out taken, L1 is a few KB and every __ldg would be an L2 round trip.
// dst[i] = src[col0 + i] for i < ncols (0 beyond N).  ncols is a multiple of 4.
__device__ __forceinline__ void stage_vec(float* dst, const float* __restrict__ src, int col0, int ncols, int N, int lane) {
  for (int i = lane * 4; i < ncols; i += 128) *reinterpret_cast<float4*>(dst + i) = ldvec4(src, col0 + i, N);
}
// Plain C++ accesses to the staging areas: pointers are derived from the extern __shared__ array without integer casts,
// so they compile to LDS / STS and -- unlike asm volatile wrappers -- may be overlapped by the scheduler.
__device__ __forceinline__ float4 lds_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ uint4 ld_tile16(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st_tile16(uint8_t* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
// all rows of this warp (row0 .. row0+31, clipped to M) belong to one sample?
__device__ __forceinline__ bool warp_rows_one_sample(int row0, int M, int rows_per_sample) {
  const int last = min(row0 + 31, M - 1);
  return row0 < M && row0 / rows_per_sample == last / rows_per_sample;
}

// Output staging shared by the store-only epilogues: two 4 KB tiles per warp (one TMA store per 32 rows x 128 B)
// followed by 2 KB for staged per-column vectors.
struct StoreRing {
  static constexpr int kWarpBytes = 2 * 4096 + 2048;
  static constexpr int kVecOff = 2 * 4096;
  static constexpr int kBars = 0;
  static constexpr int kCtaBytes = 0;
  static constexpr int kWarps = 8;
  template <class P>
  static __device__ __forceinline__ void cta_init(const P&, uint8_t*, int, int) {}
  struct State { int seq; };
  static __device__ __forceinline__ uint8_t* acquire(const EpiCtx& c, State& st) {
    if (c.el) tma_store_wait_read<1>();                // the store issued two chunks ago no longer reads its tile
    __syncwarp();
    return c.smem + (st.seq & 1) * 4096;
  }
  static __device__ __forceinline__ void release(const EpiCtx& c, State& st, const CUtensorMap* map, uint8_t* tile, int col,
                                                 int row) {
    fence_proxy_async_smem();
    __syncwarp();
    if (c.el) {
      tma_store_2d(map, tile, col, row);
      tma_store_commit();
    }
    ++st.seq;
  }
};

// ---------------------------------------------------------------------------------------------
// Epilogues.  Contract (GC = columns of the tile this warp's group works on, cbase = first of them):
//   cta_init(p, cta_smem, tid, nthreads)                        all threads of the CTA, before the role split
//   begin<BN>(p, g, sched, ctx, st)                             once per epilogue warp before the first tile
//   run<BN,GC>(p, g, sched, ctx, st, acc_taddr, row0, n0, cbase) per tile; acc_taddr = TMEM address of this warp's 32
//                                                               lanes, column 0 of the accumulator stage; thread `lane`
//                                                               owns row row0+lane
//   end(ctx, st)                                                once per warp after the last tile (drains TMA stores)
// ---------------------------------------------------------------------------------------------

// out[M,N] (bf16 or fp32) = act(acc + bias[col]);  used for adaLN, the per-sample shift vectors and the VMAE linears.
template <typename OutT, int ACT /*0 none, 1 gelu-erf, 2 gelu-tanh*/>
struct EpiStore : StoreRing {
  struct Params {
    CUtensorMap omap;    // out [M, N]: box {64 bf16 | 32 fp32, 32 rows}, SWIZZLE_128B
    const float* bias;   // [N] or nullptr
  };
  static __device__ __forceinline__ void prefetch_maps(const Params& p) { tma_prefetch_desc(&p.omap); }
  template <int BN>
  static __device__ __forceinline__ void begin(const Params&, const GemmShape&, const TileSched&, const EpiCtx&, State& st) {
    st.seq = 0;
  }
  template <int BN, int GC>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& g, const TileSched&, const EpiCtx& c, State& st,
                                             uint32_t acc, int row0, int n0, int cbase) {
    constexpr bool kBf16 = sizeof(OutT) == 2;
    constexpr int kChunk = kBf16 ? 64 : 32;            // accumulator columns per 128-byte output row
    static_assert(GC % kChunk == 0 && GC <= 512, "EpiStore: group width must be a multiple of the staged chunk");
    float* vb = reinterpret_cast<float*>(c.smem + kVecOff);
    if (p.bias != nullptr) {
      stage_vec(vb, p.bias, n0 + cbase, GC, g.N, c.lane);
      __syncwarp();
    }
#pragma unroll 1
    for (int c0 = cbase; c0 < cbase + GC; c0 += kChunk) {
      if (n0 + c0 >= g.N) break;
      uint8_t* tile = acquire(c, st);
      float v[kChunk];
      tmem_ld32(acc + c0, v);
      if constexpr (kChunk == 64) tmem_ld32(acc + c0 + 32, v + 32);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < kChunk; j += 4) {
        if (p.bias != nullptr) {
          const float4 b = lds_f4(vb + (c0 - cbase) + j);
          v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
        if constexpr (ACT == 1) {
#pragma unroll
          for (int q = 0; q < 4; ++q) v[j + q] = gelu_erf_f(v[j + q]);
        }
        if constexpr (ACT == 2) {
#pragma unroll
          for (int q = 0; q < 4; ++q) v[j + q] = gelu_tanh_f(v[j + q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 w;
        if constexpr (kBf16) {
          w = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                         pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
        } else {
          w = make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                         __float_as_uint(v[4 * q + 3]));
        }
        st_tile16(sw128_chunk(tile, c.lane, q), w);
      }
      release(c, st, &p.omap, tile, n0 + c0, row0);
    }
  }
  static __device__ __forceinline__ void end(const EpiCtx& c, State&) {
    if (c.el) tma_store_wait_read<0>();
  }
};

// Residual update of the fp32 token stream, fused with the *next* norm's operand preparation:
//   x[row,col] += gate[b,col] * (acc + bias[col])                      (lightningdit.py:248-249)
//   anext[row,col] = bf16( x_new * gnext[b,col] )   with gnext = norm.weight * (1 + scale_b)
//   ssq[row, tile] = sum_{col in tile} x_new^2       (RMSNorm statistics, models/rmsnorm.py:63)
// so that  modulate(RMSNorm(x)) . W^T  ==  rsqrt(ssq/D+eps) * (anext . W^T) + shift_b . W^T  is finished in
// the next GEMM's epilogue without another pass over x.  gate/gnext/ssq/anext are optional (VMAE uses the plain
// residual).  b = row / rows_per_sample.
// x travels by TMA in 32-row x 32-column fp32 boxes: loaded kPF chunks ahead (across tile boundaries, i.e. while
// the tensor core still works on the tile), updated in place in shared memory and stored back from the same tile.
// This epilogue is HBM-bound (12 bytes per accumulator element), so four warps suffice.
// kTrain: the training forward additionally keeps, for the backward, the branch output before the gate
// (m = acc + bias, bf16 -> dgate = sum_t dx * m) and writes the updated stream to a second buffer (xmap_out), so that
// the stream at every norm input stays available (RMSNorm Jacobian).
template <int NXP /*pair buffers (2 x 4 KB x tiles each) per warp*/, int PFP /*prefetch distance in pairs*/, bool kTrain = false>
struct EpiResidualT {
  // The accumulator tile is walked in PAIRS of 32-column chunks (64 columns = 256 contiguous bytes of the fp32 stream per
  // row, one 128-byte row of the bf16 operand): one barrier wait, one proxy fence and one bulk-store group per pair, two
  // independent instruction streams for the scheduler to interleave.  Measured on the one-chunk-per-iteration version
  // (clock64 stamps, profiles/r02_resid_trace.txt): ~2000 clk per chunk with NO wait for memory in it -- 420 clk in the
  // lane-0 load issue (integer divisions of the tile coordinates + the ELECT / R2UR.BROADCAST / BRA.U.ANY loop nvcc wraps
  // around a TMA instruction behind a single-lane branch), 240 clk in the store issue (same), 850 clk of dependent
  // LDS -> FFMA -> STS chains with a lone warp per SM sub-partition.  Here every lane runs the control flow with
  // incrementally updated (division-free) coordinates and only the TMA / mbarrier instructions are predicated on the
  // elected lane.
  static constexpr int kWarps = 4;
  static constexpr int kCtaBytes = 0;
  static constexpr int kNXP = NXP;
  static constexpr int kPFP = PFP;
  static_assert(PFP >= 1 && PFP + 1 <= NXP, "the buffer of pair p+PFP held pair p+PFP-NXP, whose store must have been waited for (pair <= p-1)");
  static constexpr int kXBytes = 4096, kPairBytes = 2 * kXBytes, kATile = 4096;
  static constexpr int kAOff = NXP * kPairBytes;                         // one bf16 tile [32 x 64] (128-byte rows) for anext
  static constexpr int kMOff = kAOff + kATile;                           // kTrain: two bf16 tiles for m (double-buffered)
  static constexpr int kVecOff = kMOff + (kTrain ? 2 * kATile : 0);      // [3][256] floats: bias, gate, gnext of the tile
  static constexpr int kWarpBytes = kVecOff + 3072;
  static constexpr int kBars = NXP;
  struct Params {
    CUtensorMap xmap;      // x [M, N] fp32: box {32, 32}, SWIZZLE_128B (loads)
    CUtensorMap xmap_out;  // where the updated x is stored (= xmap for the in-place inference path)
    CUtensorMap amap;      // anext [M, N] bf16: box {64, 32}, SWIZZLE_128B (stores); unused when !has_anext
    CUtensorMap mmap;      // kTrain: m [M, N] bf16, box {64, 32}, SWIZZLE_128B
    const float* bias;     // [N]
    const float* gate;     // [B, gate_ld] or nullptr (=> 1)
    const float* gnext;    // [B, gnext_ld] (has_anext)
    float* ssq;            // [M, ss_slots] per-row partial sums of squares, or nullptr.  Slot = column/128 of the
                           // producing tile: no atomics, so the statistics are bit-reproducible.
    int gate_ld, gnext_ld, rows_per_sample, ss_slots, has_anext;
    int x_row_mod;         // > 0: the tile of x that is LOADED comes from row (row0 % x_row_mod) of xmap -- a [x_row_mod, N] table
                           // added to every sample (patch embedding: x = patches . W^T + bias + pos_embed[token]); 0: row0
    long long* trace;      // debug builds (-DLDMAE_GEMM_TRACE): [256 pairs][8] clock64 stamps of CTA 0, epilogue warp 0
  };
  // every lane keeps the same copy of this state (warp-uniform control flow)
  struct State { int seq, pf_seq, pf_tile, pf_pair, pf_mi, pf_ni, st_div, st_mod; };
#ifdef LDMAE_GEMM_TRACE
#define RES_STAMP(k) do { if (p.trace && blockIdx.x == 0 && s.wq == 0 && lane == 0 && st.seq < 256) p.trace[st.seq * 8 + (k)] = clock64(); } while (0)
#else
#define RES_STAMP(k) do { } while (0)
#endif
  template <class P>
  static __device__ __forceinline__ void cta_init(const P&, uint8_t*, int, int) {}
  static __device__ __forceinline__ void prefetch_maps(const Params& p) {
    tma_prefetch_desc(&p.xmap);
    tma_prefetch_desc(&p.xmap_out);
    if (p.has_anext) tma_prefetch_desc(&p.amap);
    if constexpr (kTrain) tma_prefetch_desc(&p.mmap);
  }
  // issue the x loads of every pair with sequence number <= upto (all lanes walk the schedule, the elected lane issues;
  // the target buffers are known to be free)
  template <int BN>
  static __device__ __forceinline__ void pump(const Params& p, const GemmShape& g, const TileSched& s, const EpiCtx& c, State& st,
                                              int upto) {
    while (st.pf_seq <= upto && st.pf_tile < s.total) {
      int row0 = (st.pf_mi * s.cg + s.cta_rank) * kBM + s.wq * 32;
      if (p.x_row_mod > 0) row0 %= p.x_row_mod;
      const int n0 = st.pf_ni * BN;
      const int npairs = min(BN / 64, (g.N - n0 + 63) / 64);
      const int b = st.pf_seq % kNXP;
      uint8_t* dst = c.smem + b * kPairBytes;
      const int col = n0 + st.pf_pair * 64;
      if (c.el) {
        mbar_expect_tx(&c.bars[b], kPairBytes);
        tma_load_2d(&p.xmap, &c.bars[b], dst, col, row0);
        tma_load_2d(&p.xmap, &c.bars[b], dst + kXBytes, col + 32, row0);
      }
      ++st.pf_seq;
      if (++st.pf_pair == npairs) {
        st.pf_pair = 0;
        st.pf_tile += s.stride;
        st.pf_mi += st.st_div;
        st.pf_ni += st.st_mod;
        if (st.pf_ni >= s.n_tiles) { st.pf_ni -= s.n_tiles; ++st.pf_mi; }
      }
    }
  }
  template <int BN>
  static __device__ __forceinline__ void begin(const Params& p, const GemmShape& g, const TileSched& s, const EpiCtx& c, State& st) {
    st.seq = 0; st.pf_seq = 0; st.pf_tile = s.first; st.pf_pair = 0;
    st.pf_mi = s.first / s.n_tiles; st.pf_ni = s.first % s.n_tiles;
    st.st_div = s.stride / s.n_tiles; st.st_mod = s.stride % s.n_tiles;
    pump<BN>(p, g, s, c, st, kPFP - 1);
  }
  // one sample per warp (always true when rows_per_sample is a multiple of 32): the tile's column vectors are staged in
  // shared memory (kStaged); otherwise every thread reads its own sample's vectors from global memory.  Two
  // instantiations, so that the hot loop of the common case is branch-free.
  template <int BN, int GC>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& g, const TileSched& s, const EpiCtx& c, State& st,
                                             uint32_t acc, int row0, int n0, int cbase) {
    if (warp_rows_one_sample(row0, g.M, p.rows_per_sample)) run_t<BN, GC, true>(p, g, s, c, st, acc, row0, n0, cbase);
    else run_t<BN, GC, false>(p, g, s, c, st, acc, row0, n0, cbase);
  }
  template <int BN, int GC, bool kStaged>
  static __device__ __forceinline__ void run_t(const Params& p, const GemmShape& g, const TileSched& s, const EpiCtx& c, State& st,
                                               uint32_t acc, int row0, int n0, int /*cbase*/) {
    static_assert(GC == BN && BN <= 256 && BN % 64 == 0, "the residual epilogue walks the whole tile width in 64-column pairs");
    const int lane = c.lane;
    const int my_row = row0 + lane;
    const int my_b = (my_row < g.M ? my_row : g.M - 1) / p.rows_per_sample;
    const float* gate = p.gate ? p.gate + static_cast<size_t>(my_b) * p.gate_ld : nullptr;
    const float* gnext = p.has_anext ? p.gnext + static_cast<size_t>(my_b) * p.gnext_ld : nullptr;
    float* vbias = reinterpret_cast<float*>(c.smem + kVecOff);
    float* vgate = vbias + 256;
    float* vgnext = vbias + 512;
    if constexpr (kStaged) {
      stage_vec(vbias, p.bias, n0, BN, g.N, lane);
      if (gate != nullptr) stage_vec(vgate, gate, n0, BN, g.N, lane);
      if (gnext != nullptr) stage_vec(vgnext, gnext, n0, BN, g.N, lane);
      __syncwarp();
    }
    float2 ss2 = make_float2(0.f, 0.f);
    uint8_t* at = c.smem + kAOff;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 64) {
      if (n0 + c0 >= g.N) break;
      RES_STAMP(0);
      uint8_t* xt = c.smem + (st.seq % kNXP) * kPairBytes;
      uint8_t* mt = c.smem + kMOff + (st.seq & 1) * kATile;
      mbar_wait(&c.bars[st.seq % kNXP], (st.seq / kNXP) & 1, 500);
      RES_STAMP(1);
      float v[64];
      tmem_ld32(acc + c0, v);
      tmem_ld32(acc + c0 + 32, v + 32);
      tmem_ld_wait();
      RES_STAMP(2);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int col = n0 + c0 + 4 * q;
        float4 bi, gt = make_float4(1.f, 1.f, 1.f, 1.f);
        if constexpr (kStaged) {
          bi = lds_f4(vbias + c0 + 4 * q);
          if (gate != nullptr) gt = lds_f4(vgate + c0 + 4 * q);
        } else {
          bi = ldvec4(p.bias, col, g.N);
          if (gate != nullptr) gt = ldvec4(gate, col, g.N);
        }
        uint8_t* xp = sw128_chunk(xt + (q >> 3) * kXBytes, lane, q & 7);
        const uint4 xr = ld_tile16(xp);
        // packed fp32x2 arithmetic: half the issue slots of the scalar form (this loop is issue/latency-bound with one warp
        // per SM sub-partition); the row statistics accumulate in two lanes (even / odd columns), summed at the end
        const float2 m01 = add2(make_float2(v[4 * q], v[4 * q + 1]), make_float2(bi.x, bi.y));
        const float2 m23 = add2(make_float2(v[4 * q + 2], v[4 * q + 3]), make_float2(bi.z, bi.w));
        if constexpr (kTrain) {
          // 4 bf16 = 8 bytes: half of 16-byte chunk q/2 of the 128-byte row
          *reinterpret_cast<uint2*>(sw128_chunk(mt, lane, q >> 1) + (q & 1) * 8) = make_uint2(pack_bf16x2(m01.x, m01.y), pack_bf16x2(m23.x, m23.y));
        }
        const float2 x01 = fma2(m01, make_float2(gt.x, gt.y), make_float2(__uint_as_float(xr.x), __uint_as_float(xr.y)));
        const float2 x23 = fma2(m23, make_float2(gt.z, gt.w), make_float2(__uint_as_float(xr.z), __uint_as_float(xr.w)));
        st_tile16(xp, make_uint4(__float_as_uint(x01.x), __float_as_uint(x01.y), __float_as_uint(x23.x), __float_as_uint(x23.y)));
        // columns >= N hold zeros (TMA zero fill + zero bias/acc), so they do not disturb the statistics
        ss2 = fma2(x01, x01, ss2);
        ss2 = fma2(x23, x23, ss2);
        v[4 * q] = x01.x; v[4 * q + 1] = x01.y; v[4 * q + 2] = x23.x; v[4 * q + 3] = x23.y;
      }
      RES_STAMP(3);
      // every bulk store issued so far (pair seq-1 and older) has read its tiles: the anext tile and the pair buffer the
      // next load goes into are free
      if (c.el) tma_store_wait_read<0>();
      __syncwarp();
      pump<BN>(p, g, s, c, st, st.seq + kPFP);
      RES_STAMP(4);
      if (p.has_anext) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int col = n0 + c0 + 8 * q;
          float4 g0, g1;
          if constexpr (kStaged) { g0 = lds_f4(vgnext + c0 + 8 * q); g1 = lds_f4(vgnext + c0 + 8 * q + 4); }
          else { g0 = ldvec4(gnext, col, g.N); g1 = ldvec4(gnext, col + 4, g.N); }
          const float2 a0 = mul2(make_float2(v[8 * q], v[8 * q + 1]), make_float2(g0.x, g0.y));
          const float2 a1 = mul2(make_float2(v[8 * q + 2], v[8 * q + 3]), make_float2(g0.z, g0.w));
          const float2 a2 = mul2(make_float2(v[8 * q + 4], v[8 * q + 5]), make_float2(g1.x, g1.y));
          const float2 a3 = mul2(make_float2(v[8 * q + 6], v[8 * q + 7]), make_float2(g1.z, g1.w));
          st_tile16(sw128_chunk(at, lane, q),
                    make_uint4(pack_bf16x2(a0.x, a0.y), pack_bf16x2(a1.x, a1.y), pack_bf16x2(a2.x, a2.y), pack_bf16x2(a3.x, a3.y)));
        }
      }
      RES_STAMP(5);
      fence_proxy_async_smem();
      __syncwarp();
      if (c.el) {
        tma_store_2d(&p.xmap_out, xt, n0 + c0, row0);
        if (n0 + c0 + 32 < g.N) tma_store_2d(&p.xmap_out, xt + kXBytes, n0 + c0 + 32, row0);
        if (p.has_anext) tma_store_2d(&p.amap, at, n0 + c0, row0);
        if constexpr (kTrain) tma_store_2d(&p.mmap, mt, n0 + c0, row0);
        tma_store_commit();
      }
      RES_STAMP(6);
      ++st.seq;
    }
    if (p.ssq != nullptr && my_row < g.M) {
      float* dst = p.ssq + static_cast<size_t>(my_row) * p.ss_slots + n0 / 128;
      dst[0] = ss2.x + ss2.y;
      if (BN > 128 && n0 / 128 + 1 < p.ss_slots) dst[1] = 0.f;
    }
  }
  static __device__ __forceinline__ void end(const EpiCtx& c, State&) {
    if (c.el) tma_store_wait_read<0>();
    __syncwarp();
  }
};

using EpiResidual = EpiResidualT<2, 1>;        // w3 (K = 4D/1.5: tensor-bound, keeps a 4-stage operand ring)
using EpiResidualDeep = EpiResidualT<3, 2>;    // proj (K = D: HBM-bound, deeper residual prefetch, 3-stage operand ring)
using EpiResidualTrain = EpiResidualT<2, 1, true>;   // training forward (keeps m and the per-norm stream copies)

// QKV projection of LightningDiT attention (lightningdit.py:68-74) on the pre-scaled operand:
//   v = acc * rsqrt(ssq[row]/D + eps) + cvec[b,col]        (= Linear(modulate(RMSNorm(x))) incl. bias)
//   q,k heads: RMSNorm over head_dim (fp32) * weight, then 2-D axial RoPE on adjacent pairs
//   (models/pos_embed.py:38-42,135); v heads pass through.  Output bf16 [M, 3D], column = (which, head, d).
// Requires head_dim == 64 (one 128-byte output row = one head).  RoPE angles come from a compact table
// rope[axis][pos][16 cos | 16 sin] (axis 0 = token row h for dims 0..31, axis 1 = token column w for dims 32..63),
// derived from the reference's [T, 64] buffers at load time and held in shared memory (row pitch 36 floats:
// conflict-free 16-byte reads when the lanes of a warp walk consecutive positions).
struct EpiQKV : StoreRing {
  static constexpr int kCtaBytes = 10240;
  static constexpr int kRopePitch = 36;
  struct Params {
    CUtensorMap omap;       // out [M, 3D] bf16: box {64, 32}, SWIZZLE_128B
    CUtensorMap rawmap;     // has_raw (training forward): q, k before the head norm [M, 2D] bf16, same box
    int has_raw;
    const float* ssq;       // [M, ss_slots] partial sums of squares of the residual-stream row (nullptr: no row scale)
    const float* cvec;      // [B, 3D] (row pitch cvec_ld)  shift_b . W^T + bias
    int cvec_ld;
    const float* qw;        // [64] q_norm.weight or nullptr (no qk-norm)
    const float* kw;        // [64]
    const float* qb;        // [64] q_norm.bias: the head norm is nn.LayerNorm (use_rmsnorm=False, lightningdit.py:57-61); nullptr: RMSNorm
    const float* kb;        // [64]
    const float* rope;      // [2, grid, 32] or nullptr
    int D, rows_per_sample, ss_slots, grid;
    float inv_D, eps_row, eps_head;
    float q_mul;            // extra factor on q_norm.weight: 1, or softmax scale * log2(e) when the attention kernel takes
                            // its exponents straight from the scores (inference forward, run_attention prescaled)
  };
  static __device__ __forceinline__ bool rope_in_smem(const Params& p) {
    return p.rope != nullptr && 2 * p.grid * kRopePitch * 4 <= kCtaBytes;
  }
  static __device__ __forceinline__ void cta_init(const Params& p, uint8_t* cta, int tid, int nthreads) {
    if (!rope_in_smem(p)) return;
    float* dst = reinterpret_cast<float*>(cta);
    for (int i = tid; i < 2 * p.grid * 32; i += nthreads) dst[(i / 32) * kRopePitch + (i % 32)] = __ldg(p.rope + i);
  }
  static __device__ __forceinline__ void prefetch_maps(const Params& p) {
    tma_prefetch_desc(&p.omap);
    if (p.has_raw) tma_prefetch_desc(&p.rawmap);
  }
  template <int BN>
  static __device__ __forceinline__ void begin(const Params&, const GemmShape&, const TileSched&, const EpiCtx&, State& st) {
    st.seq = 0;
  }
  template <int BN, int GC>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& g, const TileSched& s, const EpiCtx& c, State& st,
                                             uint32_t acc, int row0, int n0, int cbase) {
    // fast instantiation: one sample per warp and the RoPE table in shared memory (branch-free hot loop)
    if (warp_rows_one_sample(row0, g.M, p.rows_per_sample) && (p.rope == nullptr || rope_in_smem(p)))
      run_t<BN, GC, true>(p, g, s, c, st, acc, row0, n0, cbase);
    else run_t<BN, GC, false>(p, g, s, c, st, acc, row0, n0, cbase);
  }
  template <int BN, int GC, bool kFast>
  static __device__ __forceinline__ void run_t(const Params& p, const GemmShape& g, const TileSched&, const EpiCtx& c, State& st,
                                               uint32_t acc, int row0, int n0, int cbase) {
    static_assert(GC % 64 == 0 && GC <= 256, "QKV epilogue works on whole 64-wide heads");
    const int lane = c.lane;
    const int my_row = min(row0 + lane, g.M - 1);
    const int my_b = my_row / p.rows_per_sample;
    const int tok = my_row % p.rows_per_sample;
    const float rinv = row_rinv(p.ssq, my_row, p.ss_slots, p.inv_D, p.eps_row);
    const float* cv = p.cvec + static_cast<size_t>(my_b) * p.cvec_ld;
    const bool staged = kFast || warp_rows_one_sample(row0, g.M, p.rows_per_sample);
    float* vcv = reinterpret_cast<float*>(c.smem + kVecOff);       // [GC] cvec of this tile / group
    float* vnw = vcv + 256;                                        // [64 q_norm | 64 k_norm]
    if (staged) stage_vec(vcv, cv, n0 + cbase, GC, g.N, lane);
    if (p.qw != nullptr) {
      float4 nw4 = lane < 16 ? __ldg(reinterpret_cast<const float4*>(p.qw) + lane) : __ldg(reinterpret_cast<const float4*>(p.kw) + lane - 16);
      const float mul = lane < 16 ? p.q_mul : 1.f;
      nw4.x *= mul; nw4.y *= mul; nw4.z *= mul; nw4.w *= mul;
      *reinterpret_cast<float4*>(vnw + lane * 4) = nw4;
      if (p.qb != nullptr)
        *reinterpret_cast<float4*>(vnw + 128 + lane * 4) =
            lane < 16 ? __ldg(reinterpret_cast<const float4*>(p.qb) + lane) : __ldg(reinterpret_cast<const float4*>(p.kb) + lane - 16);
    }
    __syncwarp();
    const bool rsm = kFast || rope_in_smem(p);
    // two pointer pairs so that the shared-memory copy is read with LDS (a pointer that may be either space is generic)
    const float* cta_tab = reinterpret_cast<const float*>(c.cta);
    const float* sm_h = cta_tab + (tok / p.grid) * kRopePitch;
    const float* sm_w = cta_tab + (p.grid + tok % p.grid) * kRopePitch;
    const float* gl_h = p.rope ? p.rope + static_cast<size_t>(tok / p.grid) * 32 : nullptr;
    const float* gl_w = p.rope ? p.rope + static_cast<size_t>(p.grid + tok % p.grid) * 32 : nullptr;
#pragma unroll 1
    for (int c0 = cbase; c0 < cbase + GC; c0 += 64) {
      const int colbase = n0 + c0;
      if (colbase >= g.N) break;
      const int which = colbase / p.D;                 // 0 q, 1 k, 2 v  (D % 64 == 0 => uniform per chunk)
      uint8_t* tile = acquire(c, st);
      float v[64];
      tmem_ld32(acc + c0, v);
      tmem_ld32(acc + c0 + 32, v + 32);
      tmem_ld_wait();
      // packed fp32x2 arithmetic (half the issue slots); the head's sum of squares runs in two lanes (even / odd columns)
      float2 ms2 = make_float2(0.f, 0.f);
      const float2 r2 = make_float2(rinv, rinv);
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        float4 cc;
        if constexpr (kFast) cc = lds_f4(vcv + (c0 - cbase) + j);
        else cc = staged ? lds_f4(vcv + (c0 - cbase) + j) : ldvec4(cv, colbase + j, g.N);
        const float2 a01 = fma2(make_float2(v[j], v[j + 1]), r2, make_float2(cc.x, cc.y));
        const float2 a23 = fma2(make_float2(v[j + 2], v[j + 3]), r2, make_float2(cc.z, cc.w));
        ms2 = fma2(a01, a01, ms2);
        ms2 = fma2(a23, a23, ms2);
        v[j] = a01.x; v[j + 1] = a01.y; v[j + 2] = a23.x; v[j + 3] = a23.y;
      }
      const float ms = ms2.x + ms2.y;
      if (which < 2 && p.has_raw) {
        // the head-norm Jacobian of the backward needs the un-normed head
#pragma unroll
        for (int q = 0; q < 8; ++q)
          st_tile16(sw128_chunk(tile, lane, q),
                 make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                            pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
        release(c, st, &p.rawmap, tile, colbase, row0);
        tile = acquire(c, st);
      }
      if (which < 2 && p.qw != nullptr && p.qb == nullptr) {
        const float hs = rsqrtf(ms * (1.f / 64.f) + p.eps_head);
        const float2 h2 = make_float2(hs, hs);
        const float* nw = vnw + which * 64;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          const float4 w4 = lds_f4(nw + j);
          const float2 a01 = mul2(make_float2(v[j], v[j + 1]), mul2(h2, make_float2(w4.x, w4.y)));
          const float2 a23 = mul2(make_float2(v[j + 2], v[j + 3]), mul2(h2, make_float2(w4.z, w4.w)));
          v[j] = a01.x; v[j + 1] = a01.y; v[j + 2] = a23.x; v[j + 3] = a23.y;
        }
      }
      if (which < 2 && p.qb != nullptr) {
        // nn.LayerNorm(head_dim) with affine parameters (fallback variant): centred variance in registers
        float sm = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) sm += v[j];
        const float mu = sm * (1.f / 64.f);
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) { const float d = v[j] - mu; var = fmaf(d, d, var); }
        const float hs = rsqrtf(var * (1.f / 64.f) + p.eps_head);
        const float* nw = vnw + which * 64;
        const float* nb = vnw + 128 + which * 64;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          const float4 w4 = lds_f4(nw + j), b4 = lds_f4(nb + j);
          v[j] = (v[j] - mu) * hs * w4.x + b4.x; v[j + 1] = (v[j + 1] - mu) * hs * w4.y + b4.y;
          v[j + 2] = (v[j + 2] - mu) * hs * w4.z + b4.z; v[j + 3] = (v[j + 3] - mu) * hs * w4.w + b4.w;
        }
      }
      if (which < 2 && p.rope != nullptr) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const float* stab = half == 0 ? sm_h : sm_w;
          const float* gtab = half == 0 ? gl_h : gl_w;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 cs, sn;
            if (kFast || rsm) { cs = lds_f4(stab + i); sn = lds_f4(stab + 16 + i); }
            else { cs = __ldg(reinterpret_cast<const float4*>(gtab + i)); sn = __ldg(reinterpret_cast<const float4*>(gtab + 16 + i)); }
            const float cq[4] = {cs.x, cs.y, cs.z, cs.w}, sq[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int d = half * 32 + 2 * (i + q);
              const float a = v[d], b = v[d + 1];
              v[d] = a * cq[q] - b * sq[q];             // t*cos + rotate_half(t)*sin, rotate: (x0,x1)->(-x1,x0)
              v[d + 1] = b * cq[q] + a * sq[q];
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        st_tile16(sw128_chunk(tile, lane, q),
               make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                          pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
      release(c, st, &p.omap, tile, colbase, row0);
    }
  }
  static __device__ __forceinline__ void end(const EpiCtx& c, State&) {
    if (c.el) tma_store_wait_read<0>();
  }
};

// The same QKV epilogue for heads wider than 64 (LightningDiT-XL: head_dim 72): every head occupies 128 accumulator /
// output columns, the real head_dim `hd` first and zeros behind (the packed weight and bias rows of the padding are zero,
// so the padded accumulators are exactly 0 and take no part in the statistics).  Two passes over the head (128 fp32 values
// per thread do not fit the register budget next to the rest of the kernel): sum of squares, then normalise + RoPE + store.
// Only the 16-column groups of the second half that hold data are read from TMEM and worked on (hd 72: one of four); the
// per-sample column vector and the norm weight are staged in shared memory; the RoPE angles come from a compact axial table
// tab[axis][pos][hd/4 x (cos, sin)] (axis 0 = token row for dims < hd/2, axis 1 = token column) held in shared memory
// -- derived from the reference's [T, hd] buffers and checked against them at load time (ldmae_dit_finalize); buffers without
// that structure (rope_tab == nullptr) are read element by element from global memory.
struct EpiQKVWide : StoreRing {
  static constexpr int kCtaBytes = 10240;
  struct Params {
    CUtensorMap omap;       // out [M, 3 * heads * 128] bf16: box {64, 32}, SWIZZLE_128B
    CUtensorMap rawmap;     // has_raw (training forward): q, k before the head norm [M, 2 * heads * 128] bf16, same box
    int has_raw;
    const float* ssq;       // [M, ss_slots]
    const float* cvec;      // [B, N] (row pitch cvec_ld)
    int cvec_ld;
    const float* qw;        // [128] q_norm.weight zero-padded, or nullptr
    const float* kw;        // [128]
    const float* rope_cos;  // [T, hd] or nullptr
    const float* rope_sin;
    const float* rope_tab;  // [2, grid, hd/2] interleaved (cos, sin) per angle, or nullptr (hd % 8 != 0 or no axial structure)
    int grid;
    int section, hd, rows_per_sample, ss_slots;     // section = heads * 128 = width of each of the q | k | v column ranges
    float inv_D, eps_row, eps_head;
    int acc_stride;         // accumulator columns between consecutive heads of a tile: 128, or GemmShape::n_live when only the live
                            // weight rows are multiplied
  };
  // row pitch (floats) of the table in shared memory: hd/2, plus 4 when that would put the 16-byte reads of 8 consecutive rows
  // on the same banks
  static __device__ __forceinline__ int rope_pitch(const Params& p) { return (p.hd / 2) + (((p.hd / 8) & 1) ? 0 : 4); }
  static __device__ __forceinline__ bool rope_in_smem(const Params& p) {
    return p.rope_tab != nullptr && 2 * p.grid * rope_pitch(p) * 4 <= kCtaBytes;
  }
  static __device__ __forceinline__ void cta_init(const Params& p, uint8_t* cta, int tid, int nthreads) {
    if (!rope_in_smem(p)) return;
    float* dst = reinterpret_cast<float*>(cta);
    const int w = p.hd / 2, pitch = rope_pitch(p);
    for (int i = tid; i < 2 * p.grid * w; i += nthreads) dst[(i / w) * pitch + (i % w)] = __ldg(p.rope_tab + i);
  }
  static __device__ __forceinline__ void prefetch_maps(const Params& p) {
    tma_prefetch_desc(&p.omap);
    if (p.has_raw) tma_prefetch_desc(&p.rawmap);
  }
  template <int BN>
  static __device__ __forceinline__ void begin(const Params&, const GemmShape&, const TileSched&, const EpiCtx&, State& st) {
    st.seq = 0;
  }
  static __device__ __forceinline__ void store_tile(uint8_t* tile, int lane, const float* v) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      st_tile16(sw128_chunk(tile, lane, q),
                make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                           pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
  }
  // 16 live columns followed by 48 zero columns
  static __device__ __forceinline__ void store_tile16z(uint8_t* tile, int lane, const float* v) {
    st_tile16(sw128_chunk(tile, lane, 0), make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
    st_tile16(sw128_chunk(tile, lane, 1), make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15])));
#pragma unroll
    for (int q = 2; q < 8; ++q) st_tile16(sw128_chunk(tile, lane, q), make_uint4(0u, 0u, 0u, 0u));
  }
  // v[0..NC) = v * rinv + cvec[col0 ..), returns the running (even, odd) sums of squares
  template <int NC>
  static __device__ __forceinline__ float2 add_cvec(float* v, const float* vcv, const float* cv, int colbase, int col0, int N, float rinv,
                                                    bool staged, float2 ms2) {
    const float2 r2 = make_float2(rinv, rinv);
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
      const float4 cc = staged ? lds_f4(vcv + col0 + j) : ldvec4(cv, colbase + col0 + j, N);
      const float2 a01 = fma2(make_float2(v[j], v[j + 1]), r2, make_float2(cc.x, cc.y));
      const float2 a23 = fma2(make_float2(v[j + 2], v[j + 3]), r2, make_float2(cc.z, cc.w));
      ms2 = fma2(a01, a01, ms2);
      ms2 = fma2(a23, a23, ms2);
      v[j] = a01.x; v[j + 1] = a01.y; v[j + 2] = a23.x; v[j + 3] = a23.y;
    }
    return ms2;
  }
  // v[0..NC) *= hs * weight[col0 ..)
  template <int NC>
  static __device__ __forceinline__ void scale_w(float* v, const float* vnw, int col0, float hs) {
    const float2 h2 = make_float2(hs, hs);
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
      const float4 w4 = lds_f4(vnw + col0 + j);
      const float2 a01 = mul2(make_float2(v[j], v[j + 1]), mul2(h2, make_float2(w4.x, w4.y)));
      const float2 a23 = mul2(make_float2(v[j + 2], v[j + 3]), mul2(h2, make_float2(w4.z, w4.w)));
      v[j] = a01.x; v[j + 1] = a01.y; v[j + 2] = a23.x; v[j + 3] = a23.y;
    }
  }
  template <int BN, int GC>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& g, const TileSched&, const EpiCtx& c, State& st,
                                             uint32_t acc, int row0, int n0, int cbase) {
    static_assert(GC % 128 == 0 && GC <= 256, "wide QKV epilogue works on whole 128-column heads");
    const int lane = c.lane;
    const int my_row = min(row0 + lane, g.M - 1);
    const int my_b = my_row / p.rows_per_sample;
    const int tok = my_row % p.rows_per_sample;
    const float rinv = row_rinv(p.ssq, my_row, p.ss_slots, p.inv_D, p.eps_row);
    const float* cv = p.cvec + static_cast<size_t>(my_b) * p.cvec_ld;
    const float inv_hd = 1.f / static_cast<float>(p.hd);
    const bool staged = warp_rows_one_sample(row0, g.M, p.rows_per_sample);
    float* vcv = reinterpret_cast<float*>(c.smem + kVecOff);       // [128] cvec of the head
    float* vnw = vcv + 128;                                        // [128] q_norm / k_norm weight of the head
    const int ng = (p.hd - 64 + 15) >> 4;                          // 16-column groups of the second half that hold data
    // RoPE table rows of this thread's token: shared memory (pitch rope_pitch) or the global compact table (pitch hd/2)
    const bool rsm = rope_in_smem(p);
    const int F2 = p.hd / 2;                                       // floats per axis row = 2 x angles per axis
    const int rpitch = rsm ? rope_pitch(p) : F2;
    const float* rbase = rsm ? reinterpret_cast<const float*>(c.cta) : p.rope_tab;
    const float* r_h = rbase ? rbase + static_cast<size_t>(tok / max(p.grid, 1)) * rpitch : nullptr;
    const float* r_w = rbase ? rbase + static_cast<size_t>(p.grid + tok % max(p.grid, 1)) * rpitch : nullptr;
    // angles of head dims d .. d+3 (two adjacent pairs): (cos a, sin a, cos a', sin a')
    const float* r_w2 = r_w ? r_w - F2 : nullptr;                   // so that both axes are addressed by the head dim itself
    auto angles = [&](int d) -> float4 {                           // 2 floats per angle, one angle per 2 dims: offset = dim within the axis
      const float* src = (d >= F2 ? r_w2 : r_h) + d;
      return rsm ? lds_f4(src) : __ldg(reinterpret_cast<const float4*>(src));
    };
    auto rope4 = [&](float* v4, int d) {                           // dims d .. d+3 (two adjacent pairs)
      const float4 cs = angles(d);
      const float a0 = v4[0], b0 = v4[1], a1 = v4[2], b1 = v4[3];
      v4[0] = a0 * cs.x - b0 * cs.y;                               // t*cos + rotate_half(t)*sin, rotate: (x0,x1)->(-x1,x0)
      v4[1] = b0 * cs.x + a0 * cs.y;
      v4[2] = a1 * cs.z - b1 * cs.w;
      v4[3] = b1 * cs.z + a1 * cs.w;
    };
    auto cvec4 = [&](int col_in_head, int colbase) -> float4 {
      return staged ? lds_f4(vcv + col_in_head) : ldvec4(cv, colbase + col_in_head, g.N);
    };
#pragma unroll 1
    for (int c0 = cbase; c0 < cbase + GC; c0 += 128) {
      const int colbase = n0 + c0;
      if (colbase >= g.N) break;
      const uint32_t hacc = acc + (c0 >> 7) * p.acc_stride;   // this head's accumulator columns
      const int which = colbase / p.section;           // 0 q, 1 k, 2 v
      const bool normed = which < 2 && p.qw != nullptr;
      const bool roped = which < 2 && p.rope_cos != nullptr;
      __syncwarp();                                    // the previous head's readers of vcv / vnw are done
      if (staged) stage_vec(vcv, cv, colbase, 128, g.N, lane);
      if (normed) *reinterpret_cast<float4*>(vnw + 4 * lane) = __ldg(reinterpret_cast<const float4*>(which == 0 ? p.qw : p.kw) + lane);
      __syncwarp();
      if (ng == 1 && (!roped || rbase != nullptr)) {
        // head_dim <= 80 (XL: 72): the head's 80 live columns fit the registers -- ONE pass over the accumulator
        float v[64], w[16];
        tmem_ld32(hacc, v);
        tmem_ld32(hacc + 32, v + 32);
        tmem_ld16(hacc + 64, w);
        tmem_ld_wait();
        float2 ms2 = add_cvec<64>(v, vcv, cv, colbase, 0, g.N, rinv, staged, make_float2(0.f, 0.f));
        ms2 = add_cvec<16>(w, vcv, cv, colbase, 64, g.N, rinv, staged, ms2);
        if (which < 2 && p.has_raw) {
          uint8_t* t0 = acquire(c, st);
          store_tile(t0, lane, v);
          release(c, st, &p.rawmap, t0, colbase, row0);
          uint8_t* t1 = acquire(c, st);
          store_tile16z(t1, lane, w);
          release(c, st, &p.rawmap, t1, colbase + 64, row0);
        }
        if (normed) {
          const float hs1 = rsqrtf((ms2.x + ms2.y) * inv_hd + p.eps_head);
          scale_w<64>(v, vnw, 0, hs1);
          scale_w<16>(w, vnw, 64, hs1);
        }
        if (roped) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) rope4(v + j, j);
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (64 + j < p.hd) rope4(w + j, 64 + j);
        }
        uint8_t* t0 = acquire(c, st);
        store_tile(t0, lane, v);
        release(c, st, &p.omap, t0, colbase, row0);
        uint8_t* t1 = acquire(c, st);
        store_tile16z(t1, lane, w);
        release(c, st, &p.omap, t1, colbase + 64, row0);
        continue;
      }
      float hs = 1.f;
      if (normed) {
        float2 ms2 = make_float2(0.f, 0.f);
        {
          float v[64];
          tmem_ld32(hacc, v);
          tmem_ld32(hacc + 32, v + 32);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const float4 cc = cvec4(j, colbase);
            const float2 a01 = fma2(make_float2(v[j], v[j + 1]), make_float2(rinv, rinv), make_float2(cc.x, cc.y));
            const float2 a23 = fma2(make_float2(v[j + 2], v[j + 3]), make_float2(rinv, rinv), make_float2(cc.z, cc.w));
            ms2 = fma2(a01, a01, ms2);
            ms2 = fma2(a23, a23, ms2);
          }
        }
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          if (gi < ng) {
            float v[16];
            tmem_ld16(hacc + 64 + 16 * gi, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 cc = cvec4(64 + 16 * gi + j, colbase);
              const float2 a01 = fma2(make_float2(v[j], v[j + 1]), make_float2(rinv, rinv), make_float2(cc.x, cc.y));
              const float2 a23 = fma2(make_float2(v[j + 2], v[j + 3]), make_float2(rinv, rinv), make_float2(cc.z, cc.w));
              ms2 = fma2(a01, a01, ms2);
              ms2 = fma2(a23, a23, ms2);
            }
          }
        }
        hs = rsqrtf((ms2.x + ms2.y) * inv_hd + p.eps_head);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint8_t* tile = acquire(c, st);
        float v[64];
        if (half == 0) {
          tmem_ld32(hacc, v);
          tmem_ld32(hacc + 32, v + 32);
        } else {
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            if (gi < ng) tmem_ld16(hacc + 64 + 16 * gi, v + 16 * gi);
            else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[16 * gi + j] = 0.f;
            }
          }
        }
        tmem_ld_wait();
        const int live = half == 0 ? 64 : 16 * ng;     // columns of this half that can be non-zero
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          if (j < live) {
            const float4 cc = cvec4(half * 64 + j, colbase);
            v[j] = fmaf(v[j], rinv, cc.x); v[j + 1] = fmaf(v[j + 1], rinv, cc.y);
            v[j + 2] = fmaf(v[j + 2], rinv, cc.z); v[j + 3] = fmaf(v[j + 3], rinv, cc.w);
          }
        }
        if (which < 2 && p.has_raw) {
          store_tile(tile, lane, v);
          release(c, st, &p.rawmap, tile, colbase + half * 64, row0);
          tile = acquire(c, st);
        }
        if (normed) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            if (j < live) {
              const float4 w4 = lds_f4(vnw + half * 64 + j);
              v[j] *= hs * w4.x; v[j + 1] *= hs * w4.y; v[j + 2] *= hs * w4.z; v[j + 3] *= hs * w4.w;
            }
          }
        }
        if (roped && rbase != nullptr) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const int d = half * 64 + j;
            if (d < p.hd) rope4(v + j, d);                      // hd % 8 == 0 here: the four dims are inside or outside together
          }
        } else if (roped) {
          const float* rc = p.rope_cos + static_cast<size_t>(tok) * p.hd;
          const float* rs = p.rope_sin + static_cast<size_t>(tok) * p.hd;
#pragma unroll
          for (int j = 0; j < 64; j += 2) {
            const int d = half * 64 + j;
            if (d < p.hd) {                                   // adjacent pairs share an angle (pos_embed.py:38-42,124-133)
              const float cs = __ldg(rc + d), sn = __ldg(rs + d);
              const float a = v[j], b2 = v[j + 1];
              v[j] = a * cs - b2 * sn;
              v[j + 1] = b2 * cs + a * sn;
            }
          }
        }
        store_tile(tile, lane, v);
        release(c, st, &p.omap, tile, colbase + half * 64, row0);
      }
    }
  }
  static __device__ __forceinline__ void end(const EpiCtx& c, State&) {
    if (c.el) tma_store_wait_read<0>();
  }
};

// SwiGLU first projection (models/swiglu_ffn.py:33-35) on the pre-scaled operand.  The packed weight
// interleaves w12 rows in groups of 64: [32 rows of x1 | the matching 32 rows of x2], so every 64
// accumulator columns yield 32 hidden values h = silu(x1) * x2 without leaving the thread.
struct EpiSwiGLU : StoreRing {
  struct Params {
    CUtensorMap omap;     // out [M, H] bf16: box {64, 32}, SWIZZLE_128B
    CUtensorMap premap;   // has_pre (training forward): pre-activation [M, 2H] bf16 (interleaved columns), same box
    int has_pre;
    const float* ssq;     // [M, ss_slots]
    const float* cvec;    // [B, 2H] in the same interleaved column order (row pitch cvec_ld)
    int cvec_ld;
    int rows_per_sample, ss_slots;
    float inv_D, eps_row;
  };
  static __device__ __forceinline__ void prefetch_maps(const Params& p) {
    tma_prefetch_desc(&p.omap);
    if (p.has_pre) tma_prefetch_desc(&p.premap);
  }
  template <int BN>
  static __device__ __forceinline__ void begin(const Params&, const GemmShape&, const TileSched&, const EpiCtx&, State& st) {
    st.seq = 0;
  }
  template <int BN, int GC>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& g, const TileSched& s, const EpiCtx& c, State& st,
                                             uint32_t acc, int row0, int n0, int cbase) {
    if (warp_rows_one_sample(row0, g.M, p.rows_per_sample)) run_t<BN, GC, true>(p, g, s, c, st, acc, row0, n0, cbase);
    else run_t<BN, GC, false>(p, g, s, c, st, acc, row0, n0, cbase);
  }
  template <int BN, int GC, bool kStaged>
  static __device__ __forceinline__ void run_t(const Params& p, const GemmShape& g, const TileSched&, const EpiCtx& c, State& st,
                                               uint32_t acc, int row0, int n0, int cbase) {
    static_assert(GC % 128 == 0 && GC <= 512, "SwiGLU epilogue consumes 128 accumulator columns per staged block");
    const int lane = c.lane;
    const int my_row = min(row0 + lane, g.M - 1);
    const int my_b = my_row / p.rows_per_sample;
    const float rinv = row_rinv(p.ssq, my_row, p.ss_slots, p.inv_D, p.eps_row);
    const float* cv = p.cvec + static_cast<size_t>(my_b) * p.cvec_ld;
    float* vcv = reinterpret_cast<float*>(c.smem + kVecOff);
    if constexpr (kStaged) {
      stage_vec(vcv, cv, n0 + cbase, GC, g.N, lane);
      __syncwarp();
    }
#pragma unroll 1
    for (int c0 = cbase; c0 < cbase + GC; c0 += 128) {
      if (n0 + c0 >= g.N) break;
      uint32_t w[32];                                  // 64 hidden values of this row (bf16 pairs), staged after both halves
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int cb = c0 + half * 64;
        float v[64];
        tmem_ld32(acc + cb, v);
        tmem_ld32(acc + cb + 32, v + 32);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          float4 cc;
          if constexpr (kStaged) cc = lds_f4(vcv + (cb - cbase) + j);
          else cc = ldvec4(cv, n0 + cb + j, g.N);
          v[j] = fmaf(v[j], rinv, cc.x); v[j + 1] = fmaf(v[j + 1], rinv, cc.y);
          v[j + 2] = fmaf(v[j + 2], rinv, cc.z); v[j + 3] = fmaf(v[j + 3], rinv, cc.w);
        }
        if (p.has_pre) {
          // SwiGLU's backward needs x1 and x2: keep the 64 pre-activation columns (own tile, own TMA store)
          uint8_t* pt = acquire(c, st);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            st_tile16(sw128_chunk(pt, lane, q),
                   make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                              pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
          release(c, st, &p.premap, pt, n0 + cb, row0);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 2)
          w[half * 16 + (j >> 1)] = pack_bf16x2(silu_f(v[j]) * v[32 + j], silu_f(v[j + 1]) * v[32 + j + 1]);
      }
      uint8_t* tile = acquire(c, st);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        st_tile16(sw128_chunk(tile, lane, q), make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]));
      release(c, st, &p.omap, tile, (n0 + c0) / 2, row0);
    }
  }
  static __device__ __forceinline__ void end(const EpiCtx& c, State&) {
    if (c.el) tma_store_wait_read<0>();
  }
};

// Final layer + unpatchify (lightningdit.py:267-272,376-389): out column = (pi*p + qi)*Cout + c,
// scattered to NCHW [B, Cstore, grid*p, grid*p]; learn_sigma keeps only the first Cstore channels.
// N is tiny (p*p*C, e.g. 16): plain coalesced stores (for a fixed output column the 32 lanes = 32 consecutive tokens).
struct EpiFinal {
  static constexpr int kWarps = 4;
  static constexpr int kWarpBytes = 0;
  static constexpr int kCtaBytes = 0;
  static constexpr int kBars = 0;
  struct State { int unused; };
  struct Params {
    float* out;          // [B, Cstore, G*p, G*p]
    const float* ssq;    // [M, ss_slots]
    const float* cvec;   // [B, N] (row pitch cvec_ld; 0 = one vector for every sample)
    int cvec_ld;
    int grid, patch, cout, cstore, rows_per_sample, ss_slots;
    float inv_D, eps_row;
  };
  template <class P>
  static __device__ __forceinline__ void cta_init(const P&, uint8_t*, int, int) {}
  static __device__ __forceinline__ void prefetch_maps(const Params&) {}
  template <int BN>
  static __device__ __forceinline__ void begin(const Params&, const GemmShape&, const TileSched&, const EpiCtx&, State&) {}
  template <int BN, int GC>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& g, const TileSched&, const EpiCtx& c, State&,
                                             uint32_t acc, int row0, int n0, int /*cbase*/) {
    const int row = row0 + c.lane;
    const int rowc = min(row, g.M - 1);
    const int b = rowc / p.rows_per_sample;
    const int tok = rowc % p.rows_per_sample;
    const int th = tok / p.grid, tw = tok % p.grid;
    const float rinv = row_rinv(p.ssq, rowc, p.ss_slots, p.inv_D, p.eps_row);
    const int HW = p.grid * p.patch;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      if (n0 + c0 >= g.N) break;
      float v[16];
      tmem_ld16(acc + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = n0 + c0 + j;
        if (row < g.M && col < g.N) {
          const int ch = col % p.cout;
          const int pq = col / p.cout;
          const int pi = pq / p.patch, qi = pq % p.patch;
          if (ch < p.cstore) {
            const float val = fmaf(v[j], rinv, __ldg(p.cvec + static_cast<size_t>(b) * p.cvec_ld + col));
            p.out[((static_cast<size_t>(b) * p.cstore + ch) * HW + (th * p.patch + pi)) * HW + tw * p.patch + qi] = val;
          }
        }
      }
    }
  }
  static __device__ __forceinline__ void end(const EpiCtx&, State&) {}
};

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
template <int BN, int CG, class Epi>
__global__ void __launch_bounds__(gemm_threads<Epi>(), 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const GemmShape g, const __grid_constant__ typename Epi::Params ep) {
  using Cfg = GemmCfg<BN, CG, Epi>;
  constexpr int kStages = Cfg::kStages;
  // 1024-byte alignment comes from the declaration (no integer round trip: the compiler must keep seeing a shared-space
  // pointer, otherwise every epilogue access becomes a generic LD.E / ST.E); checked once below
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* smem_epi = smem + kStages * Cfg::kStageBytes;                       // Epi::kWarps x kWarpBytes, then kCtaBytes
  uint8_t* smem_cta = smem_epi + Epi::kWarps * Epi::kWarpBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + Cfg::kEpiBytes);
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA   (leader CTA's copy is the live one)
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA   (each CTA's own copy)
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]        MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]      epilogue -> MMA (leader's copy)
  uint64_t* epi_bar = bars + 2 * kStages + 4;   // [Epi::kWarps * Epi::kBars] epilogue operand loads (per warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + Epi::kWarps * Epi::kBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    Epi::prefetch_maps(ep);
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
    for (int i = 0; i < Epi::kWarps * Epi::kBars; ++i) mbar_init(&epi_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<CG>(tmem_slot, Cfg::kTmemCols);
  Epi::cta_init(ep, smem_cta, threadIdx.x, blockDim.x);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tiles = (g.N + BN - 1) / BN;
  const int m_tiles = (g.M + kBM * CG - 1) / (kBM * CG);
  const int total_tiles = n_tiles * m_tiles;
  const int num_k = (g.K + kBK - 1) / kBK;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // whole-warp control flow, elected issue (a single-lane branch around the loop makes nvcc wrap every TMA instruction in
    // an ELECT / R2UR.BROADCAST / BRA.U.ANY convergence loop)
    {
      const bool issuer = elect_one();
      const uint32_t b_bytes = g.n_live > 0 ? static_cast<uint32_t>(g.n_live) * kBK * 2 : static_cast<uint32_t>(Cfg::kBBytes);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int m_blk = (tile / n_tiles) * CG + static_cast<int>(cta_rank);
        const int n_blk = tile % n_tiles;
        const int row_a = m_blk * kBM;
        const int row_w = n_blk * BN + static_cast<int>(cta_rank) * Cfg::kLoadBN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
          void* da = smem_a + stage * Cfg::kABytes;
          void* db = smem_b + stage * Cfg::kBBytesPadded;
          if (issuer) {
            if constexpr (CG == 1) {
              mbar_expect_tx(&full_bar[stage], Cfg::kABytes + b_bytes);
              tma_load_2d(&tmap_a, &full_bar[stage], da, kb * kBK, row_a);
              tma_load_2d(&tmap_w, &full_bar[stage], db, kb * kBK, row_w);
            } else {
              if (leader) mbar_expect_tx(&full_bar[stage], 2 * (Cfg::kABytes + b_bytes));
              else mbar_arrive_cluster(&full_bar[stage], 0);
              tma_load_2d_pair(&tmap_a, &full_bar[stage], da, kb * kBK, row_a);
              tma_load_2d_pair(&tmap_w, &full_bar[stage], db, kb * kBK, row_w);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      // drain: every slot's last consumer commit must have landed before this CTA may exit
      for (int s = 0; s < kStages; ++s) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 150 + stage);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp runs the control flow and only the tcgen05 instructions are issued by one elected lane: behind a
    // single-lane branch nvcc cannot prove the descriptors / TMEM addresses warp-uniform and wraps EVERY tcgen05.mma in an
    // ELECT / 5 x R2UR.BROADCAST / BRA.U.ANY convergence loop (~19 SASS instructions per 64-clk MMA: the issue rate sat at the
    // margin of the tensor pipe's, ncu: 64-70 % active) -- the same finding as in the attention kernels.
    if (leader) {
      const bool issuer = elect_one();
      const uint32_t idesc = umma_idesc_bf16(kBM * CG, g.n_live > 0 ? CG * g.n_live : BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1, 200 + as);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase, 300 + stage);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytesPadded);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t ad = umma_smem_desc_sw128(a_addr + k * 32, 1024, 0);
              const uint64_t bd = umma_smem_desc_sw128(b_addr + k * 32, 1024, 0);
              umma_bf16<CG>(tmem_d, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit<CG>(&empty_bar[stage]);                       // frees the smem slot (both CTAs)
            if (kb == num_k - 1) umma_commit<CG>(&tfull_bar[as]);     // accumulator ready (both CTAs)
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && (warp - 4) / 4 < Cfg::kGroups) {
    // ===================== epilogue =====================
    const int wq = warp & 3;                                         // TMEM lane quarter of this warp
    const int grp = (warp - 4) / 4;                                  // column group of this warp
    const int ew = warp - 4;
    const EpiCtx ctx{smem_epi + ew * Epi::kWarpBytes, smem_cta, epi_bar + ew * Epi::kBars, lane, elect_one()};
    const TileSched sched{cluster_id, num_clusters, total_tiles, n_tiles, CG, static_cast<int>(cta_rank), wq};
    typename Epi::State est;
    Epi::template begin<BN>(ep, g, sched, ctx, est);
    int it = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      int row0, n0;
      sched.template coords<BN>(tile, row0, n0);
      mbar_wait(&tfull_bar[as], aphase, 400 + as);
      __syncwarp();
      tc_fence_after();
      const uint32_t acc = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * BN;
      Epi::template run<BN, Cfg::kGroupCols>(ep, g, sched, ctx, est, acc, row0, n0, grp * Cfg::kGroupCols);
      tc_fence_before();
      __syncwarp();
      if (ctx.el) {
        if constexpr (CG == 1) mbar_arrive(&tempty_bar[as]);
        else mbar_arrive_cluster(&tempty_bar[as], 0);
      }
    }
    Epi::end(ctx, est);
  }

  __syncwarp();      // reconverge the single-lane role warps before the aligned barriers below
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, Cfg::kTmemCols);
}

}  // namespace ldmae
