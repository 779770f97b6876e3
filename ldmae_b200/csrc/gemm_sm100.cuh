// Persistent, warp-specialised tcgen05 GEMM for sm_100a:   C[M,N] = A[M,K] . W[N,K]^T
// (A = activations, row-major bf16; W = nn.Linear weight, row-major bf16 => both operands K-major).
//
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (128B swizzle) into a kStages-deep smem ring
//   warp 1      MMA issuer:   one elected thread issues tcgen05.mma (kind::f16, fp32 accumulate in TMEM)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue:     tcgen05.ld the accumulator (one row per thread), fused epilogue, coalesced stores
//
// kCG = 2 pairs two CTAs (cta_group::2, UMMA M = 256): each CTA loads its own 128 rows of A and half of
// the W tile, the leader CTA issues the MMAs for both, the accumulator rows of each CTA live in its own
// TMEM.  Two accumulator stages in TMEM overlap the epilogue of tile i with the main loop of tile i+1.
//
// The fused epilogues implement the LightningDiT block algebra (reference models/lightningdit.py:239-250):
// see struct comments below.
#pragma once
#include "ptx.cuh"

namespace ldmae {

constexpr int kBM = 128;        // rows per CTA
constexpr int kBK = 64;         // K per stage = one 128-byte swizzle atom of bf16
constexpr int kGemmThreads = 256;
constexpr int kStageWords = 32 * 66;   // per-epilogue-warp staging: 32 rows x (64+2) fp32

struct GemmShape {
  int M, N, K;
};

template <int BN, int CG>
struct GemmCfg {
  static constexpr int kLoadBN = BN / CG;                     // W rows loaded by each CTA
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = kLoadBN * kBK * 2;
  static constexpr int kBBytesPadded = (kBBytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kABytes + kBBytesPadded;
  static constexpr int kEpiBytes = 4 * kStageWords * 4;
  static constexpr int kBudget = 225 * 1024 - kEpiBytes - 1024 /*align slack*/ - 512 /*barriers*/;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kEpiBytes + 512;
  static_assert(2 * BN <= 512, "two accumulator stages must fit TMEM");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
  static_assert(kLoadBN % 8 == 0, "W rows per CTA");
};

// ---------------------------------------------------------------------------------------------
// Epilogues.  Contract:  Epi::run(p, acc_taddr, row0, n0, wq, lane, stage)
//   acc_taddr : TMEM address of this warp's 32 lanes, column 0 of the accumulator stage
//   row0      : global row of this warp's lane 0;   n0 : first global column of the tile
//   stage     : this warp's private smem staging (kStageWords floats)
// Thread `lane` owns accumulator row (row0 + lane) in the "row phase"; in the "column phase"
// the warp walks rows and lanes own columns, which makes global accesses 128-byte coalesced.
// ---------------------------------------------------------------------------------------------

// out[M,N] (bf16 or fp32) = act(acc + bias[col]);  used for adaLN, the per-sample shift vectors,
// and the VMAE linears (fc1: GELU-erf).
template <typename OutT, int ACT /*0 none, 1 gelu-erf, 2 gelu-tanh*/>
struct EpiStore {
  struct Params {
    OutT* out;
    const float* bias;   // [N] or nullptr
    int ldo;             // leading dimension of out (elements)
  };
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& g, uint32_t acc, int row0, int n0,
                                             int lane, float* stage) {
    constexpr bool kBf16 = sizeof(OutT) == 2;
    constexpr int kChunk = kBf16 ? 64 : 32;            // accumulator columns per staged block (32 words/row)
    static_assert(BN % kChunk == 0, "EpiStore: tile width must be a multiple of the staged chunk");
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += kChunk) {
      if (n0 + c0 >= g.N) break;
      float v[kChunk];
      tmem_ld32(acc + c0, v);
      if constexpr (kChunk == 64) tmem_ld32(acc + c0 + 32, v + 32);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < kChunk; ++j) {
        const int col = n0 + c0 + j;
        float b = (p.bias != nullptr && col < g.N) ? __ldg(p.bias + col) : 0.f;
        float x = v[j] + b;
        if constexpr (ACT == 1) x = gelu_erf_f(x);
        if constexpr (ACT == 2) x = gelu_tanh_f(x);
        v[j] = x;
      }
      uint32_t* sw = reinterpret_cast<uint32_t*>(stage);
      if constexpr (kBf16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) sw[lane * 33 + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) sw[lane * 33 + j] = __float_as_uint(v[j]);
      }
      __syncwarp();
      // column phase: one row per iteration, lane = word
      const int colw = n0 + c0 + (kBf16 ? 2 * lane : lane);
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const int row = row0 + r;
        if (row < g.M && colw < g.N) {
          uint32_t w = sw[r * 33 + lane];
          if constexpr (kBf16)
            *reinterpret_cast<uint32_t*>(p.out + static_cast<size_t>(row) * p.ldo + colw) = w;
          else
            p.out[static_cast<size_t>(row) * p.ldo + colw] = __uint_as_float(w);
        }
      }
      __syncwarp();
    }
  }
};

// Residual update of the fp32 token stream, fused with the *next* norm's operand preparation:
//   x[row,col] += gate[b,col] * (acc + bias[col])                      (lightningdit.py:248-249)
//   anext[row,col] = bf16( x_new * gnext[b,col] )   with gnext = norm.weight * (1 + scale_b)
//   ssq[row, tile] = sum_{col in tile} x_new^2       (RMSNorm statistics, models/rmsnorm.py:63)
// so that  modulate(RMSNorm(x)) . W^T  ==  rsqrt(ssq/D+eps) * (anext . W^T) + shift_b . W^T  is finished in
// the next GEMM's epilogue without another pass over x.  gate/gnext/ssq/anext are optional (VMAE
// uses the plain residual).  b = row / rows_per_sample.
struct EpiResidual {
  struct Params {
    float* x;              // [M, ldx] fp32, read-modify-write
    const float* bias;     // [N]
    const float* gate;     // [B, gate_ld] or nullptr (=> 1)
    const float* gnext;    // [B, gnext_ld] or nullptr
    __nv_bfloat16* anext;  // [M, ldx] or nullptr
    float* ssq;            // [M, ss_slots] per-row partial sums of squares, or nullptr.  Slot = column/128 of the
                           // producing tile: no atomics, so the statistics are bit-reproducible.
    int ldx, gate_ld, gnext_ld, rows_per_sample, ss_slots;
  };
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& g, uint32_t acc, int row0, int n0,
                                             int lane, float* stage) {
    float ss[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) ss[r] = 0.f;
    const int my_row = row0 + lane;
    const int my_b = (my_row < g.M ? my_row : g.M - 1) / p.rows_per_sample;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= g.N) break;
      float v[32];
      tmem_ld32(acc + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = n0 + c0 + j;
        if (col < g.N) {
          float y = v[j] + __ldg(p.bias + col);
          if (p.gate != nullptr) y *= __ldg(p.gate + static_cast<size_t>(my_b) * p.gate_ld + col);
          v[j] = y;
        }
        stage[lane * 33 + j] = v[j];
      }
      __syncwarp();
      const int col = n0 + c0 + lane;
      const bool col_ok = col < g.N;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const int row = row0 + r;
        if (row >= g.M) break;                                   // warp-uniform
        const size_t off = static_cast<size_t>(row) * p.ldx + (col_ok ? col : 0);
        float xn = 0.f;
        if (col_ok) {
          xn = p.x[off] + stage[r * 33 + lane];
          p.x[off] = xn;
          ss[r] += xn * xn;
        }
        if (p.anext != nullptr) {                                // warp-uniform
          const int b = row / p.rows_per_sample;
          const float a = col_ok ? xn * __ldg(p.gnext + static_cast<size_t>(b) * p.gnext_ld + col) : 0.f;
          const float a_hi = __shfl_down_sync(0xffffffffu, a, 1);
          if (col_ok && (lane & 1) == 0)
            *reinterpret_cast<uint32_t*>(p.anext + off) = pack_bf16x2(a, a_hi);
        }
      }
      __syncwarp();
    }
    if (p.ssq != nullptr) {
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        float s = ss[r];
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (lane == 0 && row0 + r < g.M) {
          float* dst = p.ssq + static_cast<size_t>(row0 + r) * p.ss_slots + n0 / 128;
          dst[0] = s;
          if (BN > 128 && n0 / 128 + 1 < p.ss_slots) dst[1] = 0.f;
        }
      }
    }
  }
};

// rsqrt(mean(x^2) + eps) of a residual-stream row from its per-tile partial sums (fixed summation order)
__device__ __forceinline__ float row_rinv(const float* ssq, int row, int slots, float inv_D, float eps) {
  if (ssq == nullptr) return 1.f;
  float s = 0.f;
  for (int j = 0; j < slots; ++j) s += __ldg(ssq + static_cast<size_t>(row) * slots + j);
  return rsqrtf(s * inv_D + eps);
}

// QKV projection of LightningDiT attention (lightningdit.py:68-74) on the pre-scaled operand:
//   v = acc * rsqrt(ssq[row]/D + eps) + cvec[b,col]        (= Linear(modulate(RMSNorm(x))) incl. bias)
//   q,k heads: RMSNorm over head_dim (fp32) * weight, then 2-D axial RoPE on adjacent pairs
//   (models/pos_embed.py:38-42,135); v heads pass through.  Output bf16 [M, 3D], column = (which, head, d).
// Requires head_dim == 64 (one staged chunk = one head).
struct EpiQKV {
  struct Params {
    __nv_bfloat16* out;     // [M, 3D]
    const float* ssq;       // [M, ss_slots] partial sums of squares of the residual-stream row (nullptr: no row scale)
    const float* cvec;      // [B, 3D]  shift_b . W^T + bias
    const float* qw;        // [64] q_norm.weight or nullptr (no qk-norm)
    const float* kw;        // [64]
    const float* rope_cos;  // [T, 64] or nullptr
    const float* rope_sin;  // [T, 64]
    int D, rows_per_sample, ss_slots;
    float inv_D, eps_row, eps_head;
  };
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& g, uint32_t acc, int row0, int n0,
                                             int lane, float* stage) {
    static_assert(BN % 64 == 0, "QKV epilogue works on whole 64-wide heads");
    const int my_row = min(row0 + lane, g.M - 1);
    const int my_b = my_row / p.rows_per_sample;
    const float rinv = row_rinv(p.ssq, my_row, p.ss_slots, p.inv_D, p.eps_row);
    const float* cv = p.cvec + static_cast<size_t>(my_b) * g.N;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 64) {
      const int colbase = n0 + c0;
      if (colbase >= g.N) break;
      const int which = colbase / p.D;                 // 0 q, 1 k, 2 v  (D % 64 == 0 => uniform per chunk)
      float v[64];
      tmem_ld32(acc + c0, v);
      tmem_ld32(acc + c0 + 32, v + 32);
      tmem_ld_wait();
      float ms = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        v[j] = fmaf(v[j], rinv, __ldg(cv + colbase + j));
        ms = fmaf(v[j], v[j], ms);
      }
      const float* nw = which == 0 ? p.qw : (which == 1 ? p.kw : nullptr);
      const float hs = (nw != nullptr) ? rsqrtf(ms * (1.f / 64.f) + p.eps_head) : 1.f;
#pragma unroll
      for (int j = 0; j < 64; j += 2)
        *reinterpret_cast<float2*>(stage + lane * 66 + j) = make_float2(v[j] * hs, v[j + 1] * hs);
      __syncwarp();
      // column phase: lane owns the adjacent pair (2*lane, 2*lane+1) = one RoPE pair
      float w0 = 1.f, w1 = 1.f;
      if (nw != nullptr) { w0 = __ldg(nw + 2 * lane); w1 = __ldg(nw + 2 * lane + 1); }
      const bool rope = (which < 2) && (p.rope_cos != nullptr);
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const int row = row0 + r;
        if (row >= g.M) break;
        float2 a = *reinterpret_cast<const float2*>(stage + r * 66 + 2 * lane);
        a.x *= w0; a.y *= w1;
        if (rope) {
          const int tok = row % p.rows_per_sample;
          const float2 c = __ldg(reinterpret_cast<const float2*>(p.rope_cos + static_cast<size_t>(tok) * 64) + lane);
          const float2 s = __ldg(reinterpret_cast<const float2*>(p.rope_sin + static_cast<size_t>(tok) * 64) + lane);
          const float ox = a.x * c.x - a.y * s.x;      // t*cos + rotate_half(t)*sin, rotate: (x0,x1)->(-x1,x0)
          const float oy = a.y * c.y + a.x * s.y;
          a.x = ox; a.y = oy;
        }
        *reinterpret_cast<uint32_t*>(p.out + static_cast<size_t>(row) * g.N + colbase + 2 * lane) = pack_bf16x2(a.x, a.y);
      }
      __syncwarp();
    }
  }
};

// SwiGLU first projection (models/swiglu_ffn.py:33-35) on the pre-scaled operand.  The packed weight
// interleaves w12 rows in groups of 64: [32 rows of x1 | the matching 32 rows of x2], so every 64
// accumulator columns yield 32 hidden values h = silu(x1) * x2 without leaving the thread.
struct EpiSwiGLU {
  struct Params {
    __nv_bfloat16* out;   // [M, H]
    const float* ssq;     // [M]
    const float* cvec;    // [B, 2H] in the same interleaved column order
    int H, rows_per_sample, ss_slots;
    float inv_D, eps_row;
  };
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& g, uint32_t acc, int row0, int n0,
                                             int lane, float* stage) {
    static_assert(BN % 128 == 0, "SwiGLU epilogue consumes 128 accumulator columns per staged block");
    const int my_row = min(row0 + lane, g.M - 1);
    const int my_b = my_row / p.rows_per_sample;
    const float rinv = row_rinv(p.ssq, my_row, p.ss_slots, p.inv_D, p.eps_row);
    const float* cv = p.cvec + static_cast<size_t>(my_b) * g.N;
    uint32_t* sw = reinterpret_cast<uint32_t*>(stage);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 128) {
      if (n0 + c0 >= g.N) break;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int cb = c0 + half * 64;
        float v[64];
        tmem_ld32(acc + cb, v);
        tmem_ld32(acc + cb + 32, v + 32);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = fmaf(v[j], rinv, __ldg(cv + n0 + cb + j));
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float h0 = silu_f(v[j]) * v[32 + j];
          const float h1 = silu_f(v[j + 1]) * v[32 + j + 1];
          sw[lane * 33 + half * 16 + (j >> 1)] = pack_bf16x2(h0, h1);
        }
      }
      __syncwarp();
      const int hcol = (n0 + c0) / 2 + 2 * lane;
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const int row = row0 + r;
        if (row < g.M && hcol < p.H)
          *reinterpret_cast<uint32_t*>(p.out + static_cast<size_t>(row) * p.H + hcol) = sw[r * 33 + lane];
      }
      __syncwarp();
    }
  }
};

// Final layer + unpatchify (lightningdit.py:267-272,376-389): out column = (pi*p + qi)*Cout + c,
// scattered to NCHW [B, Cstore, grid*p, grid*p]; learn_sigma keeps only the first Cstore channels.
struct EpiFinal {
  struct Params {
    float* out;          // [B, Cstore, G*p, G*p]
    const float* ssq;    // [M]
    const float* cvec;   // [B, N]
    int grid, patch, cout, cstore, rows_per_sample, ss_slots;
    float inv_D, eps_row;
  };
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, const GemmShape& g, uint32_t acc, int row0, int n0,
                                             int lane, float* stage) {
    (void)stage;
    const int row = row0 + lane;
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
          const int c = col % p.cout;
          const int pq = col / p.cout;
          const int pi = pq / p.patch, qi = pq % p.patch;
          if (c < p.cstore) {
            const float val = fmaf(v[j], rinv, __ldg(p.cvec + static_cast<size_t>(b) * g.N + col));
            p.out[((static_cast<size_t>(b) * p.cstore + c) * HW + (th * p.patch + pi)) * HW + tw * p.patch + qi] = val;
          }
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
template <int BN, int CG, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const GemmShape g, const typename Epi::Params ep) {
  using Cfg = GemmCfg<BN, CG>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  float* smem_epi = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + Cfg::kEpiBytes);
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA   (leader CTA's copy is the live one)
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA   (each CTA's own copy)
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]        MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]      epilogue -> MMA (leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], CG);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], CG * 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<CG>(tmem_slot, Cfg::kTmemCols);
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
    if (lane == 0) {
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
          if constexpr (CG == 1) {
            mbar_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::kBBytes);
            tma_load_2d(&tmap_a, &full_bar[stage], da, kb * kBK, row_a);
            tma_load_2d(&tmap_w, &full_bar[stage], db, kb * kBK, row_w);
          } else {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * (Cfg::kABytes + Cfg::kBBytes));
            else mbar_arrive_cluster(&full_bar[stage], 0);
            tma_load_2d_pair(&tmap_a, &full_bar[stage], da, kb * kBK, row_a);
            tma_load_2d_pair(&tmap_w, &full_bar[stage], db, kb * kBK, row_w);
          }
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
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM * CG, BN);
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
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t ad = umma_smem_desc_sw128(a_addr + k * 32, 1024, 0);
            const uint64_t bd = umma_smem_desc_sw128(b_addr + k * 32, 1024, 0);
            umma_bf16<CG>(tmem_d, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit<CG>(&empty_bar[stage]);                       // frees the smem slot (both CTAs)
          if (kb == num_k - 1) umma_commit<CG>(&tfull_bar[as]);     // accumulator ready (both CTAs)
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int wq = warp & 3;                                         // TMEM lane quarter of this warp
    float* stage_buf = smem_epi + wq * kStageWords;
    int it = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m_blk = (tile / n_tiles) * CG + static_cast<int>(cta_rank);
      const int n_blk = tile % n_tiles;
      mbar_wait(&tfull_bar[as], aphase, 400 + as);
      tc_fence_after();
      const uint32_t acc = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * BN;
      Epi::template run<BN>(ep, g, acc, m_blk * kBM + wq * 32, n_blk * BN, lane, stage_buf);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(&tempty_bar[as]);
        else mbar_arrive_cluster(&tempty_bar[as], 0);
      }
    }
  }

  __syncwarp();      // reconverge the single-lane role warps before the aligned barriers below
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, Cfg::kTmemCols);
}

}  // namespace ldmae
