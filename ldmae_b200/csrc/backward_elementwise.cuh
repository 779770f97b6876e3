// HBM-bound kernels of the training backward (reference: autograd through models/lightningdit.py:239-250, 66-91,
// models/rmsnorm.py:52-77, models/swiglu_ffn.py:31-36, driven by transport.training_losses / train_accum.py:215-230).
// They sit between the tensor-core GEMMs of the backward and carry every reduction the parameter gradients need:
//   * over the columns of a token row (RMSNorm / head-norm Jacobians)      -> warp shuffles
//   * over the tokens of a sample (adaLN shift / scale / gate gradients)   -> shared-memory accumulators per CTA,
//                                                                             then one global atomicAdd per column and CTA
// All global accesses are coalesced and vectorised (a warp walks a token row).
#pragma once
#include "elementwise.cuh"

namespace ldmae {

__device__ __forceinline__ float row_rinv_g(const float* __restrict__ ssq, size_t row, int slots, float inv_D, float eps) {
  float s = 0.f;
  for (int j = 0; j < slots; ++j) s += __ldg(ssq + row * slots + j);
  return rsqrtf(s * inv_D + eps);
}
__device__ __forceinline__ float2 bf2_to_f2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ float dsilu_f(float x) {
  const float s = __fdividef(1.f, 1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}

// ---------------------------------------------------------------------------------------------
// Backward through   a = rsqrt(mean(x^2)+eps) * (x * g_b) + shift_b   (the modulated RMSNorm in front of a Linear) joined
// with the residual add and the gate of the PREVIOUS branch.  With Gp = r * dL/da (the data-gradient GEMM's output; the
// row factor r is already folded into the GEMM's other operand):
//   dx_new  = dx + g_b * Gp - x * (r^2 / D) * sum_col(Gp * x * g_b)                       (dL/dx of this residual point)
//   dg[b]  += sum_t Gp * x                   (-> dscale = dg * w, dw = sum_b dg * (1 + scale))
//   dY      = bf16(dx_new * gate_prev[b])    (output gradient of the previous branch's Linear; gate_prev == nullptr: 1)
//   dgate_prev[b] += sum_t dx_new * m_prev   (m_prev = that branch's output before the gate, saved by the forward)
//   sdx[b] += sum_t dx_new                   (-> bias gradient of the previous branch's Linear, times the gate)
// One CTA = 32 token rows of one sample; one warp = 4 rows.
// ---------------------------------------------------------------------------------------------
struct ResidBwdParams {
  float* dx;                    // [M, D] in/out (dx_in == nullptr on entry of the final layer: treated as 0)
  const __nv_bfloat16* gp;      // [M, D]
  const float* x;               // [M, D] residual stream at this norm's input
  const float* ssq;             // [M, slots]
  const float* g;               // [B, D] norm.weight * (1 + scale)
  float* dg;                    // [B, D] accumulated
  __nv_bfloat16* dy;            // [M, D] out
  const float* gate; int gate_ld;       // [B, gate_ld] or nullptr
  const __nv_bfloat16* m_prev;  // [M, D] or nullptr
  float* dgate; int dgate_ld;   // accumulated, or nullptr
  float* sdx;                   // [B, D] accumulated, or nullptr
  int T, D, slots, dx_zero;
  float eps;
};

// Column sums stay in registers over the 8 rows a warp owns (lane = 4 columns of a 128-column chunk) and reach shared
// memory once per warp and chunk, global memory once per CTA and column.
constexpr int kResidRows = 4;   // rows per warp (8 warps per CTA): 4 keeps the kernel at <= 85 registers, three CTAs per SM
__global__ void __launch_bounds__(256, 3)
resid_bwd_kernel(const ResidBwdParams p) {
  // [8 warps][3][D]: dg, dgate, sdx -- every warp visits every 128-column chunk exactly once, so its partial column sums go
  // into a private row with plain conflict-free 16-byte stores (no shared atomics: 16 M bank-conflict cycles per launch in
  // round 1's layout, ncu) and are summed over the warps at the end
  extern __shared__ float s_all[];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * (8 * kResidRows);
  const int nrows = min(8 * kResidRows, p.T - t0);
  const int D = p.D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_acc = s_all + static_cast<size_t>(warp) * 3 * D;
  const float* g = p.g + static_cast<size_t>(b) * D;
  const float* gate = p.gate ? p.gate + static_cast<size_t>(b) * p.gate_ld : nullptr;
  const size_t row0 = static_cast<size_t>(b) * p.T + t0 + warp * kResidRows;
  const int myrows = max(0, min(kResidRows, nrows - warp * kResidRows));
  // pass A: coef[i] = r^2 / D * sum_col(Gp * x * g) of each of this warp's rows
  float coef[kResidRows];
#pragma unroll
  for (int i = 0; i < kResidRows; ++i) {
    coef[i] = 0.f;
    if (i < myrows) {
      const size_t row = row0 + i;
      const float* xr = p.x + row * D;
      const __nv_bfloat16* gr = p.gp + row * D;
      float t = 0.f;
      for (int c = lane * 4; c < D; c += 128) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + c);
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + c));
        const uint2 gw = *reinterpret_cast<const uint2*>(gr + c);
        const float2 a = bf2_to_f2(gw.x), bq = bf2_to_f2(gw.y);
        t = fmaf(a.x * xv.x, gv.x, t); t = fmaf(a.y * xv.y, gv.y, t);
        t = fmaf(bq.x * xv.z, gv.z, t); t = fmaf(bq.y * xv.w, gv.w, t);
      }
      t = warp_sum(t);
      const float rinv = row_rinv_g(p.ssq, row, p.slots, 1.f / D, p.eps);
      coef[i] = rinv * rinv * t / D;
    }
  }
  // pass B, one 128-column chunk at a time (lane = 4 columns) over the warp's 8 rows: the per-sample vectors g / gate are
  // loaded once per chunk, the 8 rows' loads are independent (deep memory-level parallelism), the column sums of the
  // chunk stay in 12 registers and reach shared memory once per warp and chunk
  for (int c = lane * 4; c < D; c += 128) {
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g + c));
    const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
    float gt[4] = {1.f, 1.f, 1.f, 1.f};
    if (gate) { const float4 q4 = __ldg(reinterpret_cast<const float4*>(gate + c)); gt[0] = q4.x; gt[1] = q4.y; gt[2] = q4.z; gt[3] = q4.w; }
    float a_dg[4] = {0.f, 0.f, 0.f, 0.f}, a_gt[4] = {0.f, 0.f, 0.f, 0.f}, a_sd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < kResidRows; ++i) {
      if (i < myrows) {
        const size_t off = (row0 + i) * D + c;
        const float4 xv = *reinterpret_cast<const float4*>(p.x + off);
        const uint2 gw = *reinterpret_cast<const uint2*>(p.gp + off);
        float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!p.dx_zero) d4 = *reinterpret_cast<const float4*>(p.dx + off);
        uint2 mw = make_uint2(0u, 0u);
        if (p.m_prev) mw = *reinterpret_cast<const uint2*>(p.m_prev + off);
        const float2 ga = bf2_to_f2(gw.x), gb = bf2_to_f2(gw.y), m0 = bf2_to_f2(mw.x), m1 = bf2_to_f2(mw.y);
        const float gpv[4] = {ga.x, ga.y, gb.x, gb.y};
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        const float mv[4] = {m0.x, m0.y, m1.x, m1.y};
        float dn[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          dn[q] = dn[q] + gs[q] * gpv[q] - xs[q] * coef[i];
          a_dg[q] = fmaf(gpv[q], xs[q], a_dg[q]);
          a_gt[q] = fmaf(dn[q], mv[q], a_gt[q]);
          a_sd[q] += dn[q];
        }
        *reinterpret_cast<float4*>(p.dx + off) = make_float4(dn[0], dn[1], dn[2], dn[3]);
        *reinterpret_cast<uint2*>(p.dy + off) = make_uint2(pack_bf16x2(dn[0] * gt[0], dn[1] * gt[1]), pack_bf16x2(dn[2] * gt[2], dn[3] * gt[3]));
      }
    }
    *reinterpret_cast<float4*>(s_acc + c) = make_float4(a_dg[0], a_dg[1], a_dg[2], a_dg[3]);
    *reinterpret_cast<float4*>(s_acc + D + c) = make_float4(a_gt[0], a_gt[1], a_gt[2], a_gt[3]);
    *reinterpret_cast<float4*>(s_acc + 2 * D + c) = make_float4(a_sd[0], a_sd[1], a_sd[2], a_sd[3]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float t0s = 0.f, t1s = 0.f, t2s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float* r = s_all + static_cast<size_t>(w) * 3 * D;
      t0s += r[c]; t1s += r[D + c]; t2s += r[2 * D + c];
    }
    atomicAdd(p.dg + static_cast<size_t>(b) * D + c, t0s);
    if (p.dgate) atomicAdd(p.dgate + static_cast<size_t>(b) * p.dgate_ld + c, t1s);
    if (p.sdx) atomicAdd(p.sdx + static_cast<size_t>(b) * D + c, t2s);
  }
}

// ---------------------------------------------------------------------------------------------
// Backward through the QKV epilogue (lightningdit.py:68-74): RoPE^T, per-head RMSNorm Jacobian, row factor.
//   in : dqkv [M, 3D] bf16 = gradients w.r.t. the rotated / normed q, k and v (attention backward output)
//        raw  [M, 2D] bf16 = q, k before the head norm (saved by the training forward)
//   out: dqkv in place     = r[row] * dL/d(qkv pre-norm)     (operand of the data- and weight-gradient GEMMs)
//        dcvec[b, 3D]     += sum_t dL/d(qkv pre-norm)        (-> bias, shift and shift-path weight gradients)
//        dqw[64], dkw[64] += sum over rows and heads of dy * xhat
// ---------------------------------------------------------------------------------------------
struct QkvBwdParams {
  __nv_bfloat16* dqkv;
  const __nv_bfloat16* raw;
  const float* ssq;
  const float* qw; const float* kw;        // nullptr: no qk-norm
  const float* rope;                        // compact table [2][G][32] or nullptr
  float* dcvec;                             // [B, 3D]
  float* dqw; float* dkw;                   // [64] each
  int T, D, slots, G;
  float eps_row, eps_head;
};

__device__ __forceinline__ float half_warp_sum(float v) {   // sum over the 16 lanes of this half-warp (every lane takes part)
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__device__ __forceinline__ float oct_sum(float v) {   // sum over the 8 lanes that share a head (every lane takes part)
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// One warp = 4 rows x FOUR 64-wide heads at a time: 8 lanes per head, lane = 8 consecutive dims (four RoPE pairs, one
// 16-byte access per row and array: half the memory instructions and 3-step instead of 4-step reductions compared with the
// round-1 layout of 4 dims per lane, which ran at 43 % of the HBM peak -- profiles/r02_c_train_launches.txt).
// One CTA = 32 rows of a sample.
constexpr int kQkvRows = 4;     // rows per warp of qkv_bwd_kernel (32-row slab per CTA)
__global__ void __launch_bounds__(256, 2)
qkv_bwd_kernel(const QkvBwdParams p) {
  // Column sums: every warp walks every column group exactly once, so it keeps a PRIVATE row of partial sums in shared
  // memory, written with plain conflict-free 16-byte stores (two planes of 4 values per lane) and summed over the 8 warps at
  // the end.  (Shared-memory atomics on a common row cost 43 M bank-conflict cycles per launch -- ncu, round 2 -- and were
  // the kernel's bottleneck.)  Layout of a warp's row: [group g of 256 columns][plane 0..1][lane][4], then [128] dqw | dkw.
  extern __shared__ float s_all[];
  const int D = p.D, N = 3 * D;
  const int NG = (N + 255) / 256 * 256;     // columns rounded up to whole groups
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * (8 * kQkvRows);
  const int nrows = min(8 * kQkvRows, p.T - t0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_acc = s_all + static_cast<size_t>(warp) * (NG + 128);
  const int hh = lane >> 3, l8 = lane & 7;
  const int d0 = 8 * l8;                                     // first of this lane's 8 dims inside a head
  const int heads3 = N / 64;
  const int myrows = max(0, min(kQkvRows, nrows - warp * kQkvRows));       // this warp owns kQkvRows consecutive rows of the slab
  const int tok0 = t0 + warp * kQkvRows;
  const size_t row0 = static_cast<size_t>(b) * p.T + tok0;
  float rinv[kQkvRows];
  float4 rc[kQkvRows], rs[kQkvRows];                          // this lane's four RoPE angles per row
#pragma unroll
  for (int i = 0; i < kQkvRows; ++i) {
    rinv[i] = 0.f; rc[i] = make_float4(1.f, 1.f, 1.f, 1.f); rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < myrows) {
      rinv[i] = row_rinv_g(p.ssq, row0 + i, p.slots, 1.f / D, p.eps_row);
      if (p.rope != nullptr) {
        const int axis = d0 >> 5, f = (d0 & 31) >> 1, tok = tok0 + i;     // dims 0..31: token row, 32..63: token column
        const int pos = axis == 0 ? tok / p.G : tok % p.G;
        const float* tab = p.rope + (static_cast<size_t>(axis) * p.G + pos) * 32;
        rc[i] = __ldg(reinterpret_cast<const float4*>(tab + f));
        rs[i] = __ldg(reinterpret_cast<const float4*>(tab + 16 + f));
      }
    }
  }
  float waq[8], wak[8];                                       // dq_norm.weight / dk_norm.weight partial sums (this lane's dims)
#pragma unroll
  for (int q = 0; q < 8; ++q) { waq[q] = 0.f; wak[q] = 0.f; }
  for (int hp = 0; hp < heads3; hp += 4) {
    const int hc = hp + hh;                                  // this lane group's head
    const bool live = hc < heads3;
    const int col = hc * 64 + d0;
    const int which = live ? (hc * 64) / D : 2;              // 0 q, 1 k, 2 v
    const bool normed = which < 2 && p.qw != nullptr;
    float wv[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if (normed) {
      const float* wsrc = (which == 0 ? p.qw : p.kw) + d0;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wsrc)), w1 = __ldg(reinterpret_cast<const float4*>(wsrc + 4));
      wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w; wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
    }
    uint4 dyw[kQkvRows], xw[kQkvRows];
#pragma unroll
    for (int i = 0; i < kQkvRows; ++i) {
      dyw[i] = make_uint4(0u, 0u, 0u, 0u); xw[i] = make_uint4(0u, 0u, 0u, 0u);
      if (i < myrows && live) {
        dyw[i] = *reinterpret_cast<const uint4*>(p.dqkv + (row0 + i) * N + col);
        if (normed) xw[i] = *reinterpret_cast<const uint4*>(p.raw + (row0 + i) * 2 * D + col);
      }
    }
    float cs[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) cs[q] = 0.f;
#pragma unroll
    for (int i = 0; i < kQkvRows; ++i) {
      const uint32_t dwd[4] = {dyw[i].x, dyw[i].y, dyw[i].z, dyw[i].w};
      const uint32_t xwd[4] = {xw[i].x, xw[i].y, xw[i].z, xw[i].w};
      const float cq[4] = {rc[i].x, rc[i].y, rc[i].z, rc[i].w}, sq[4] = {rs[i].x, rs[i].y, rs[i].z, rs[i].w};
      float dy[8], x[8];
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        const float2 a = bf2_to_f2(dwd[pr]), xx = bf2_to_f2(xwd[pr]);
        dy[2 * pr] = a.x; dy[2 * pr + 1] = a.y; x[2 * pr] = xx.x; x[2 * pr + 1] = xx.y;
      }
      if (which < 2 && p.rope != nullptr) {
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {                     // transpose of (a,b) -> (a c - b s, b c + a s)
          const float a = dy[2 * pr] * cq[pr] + dy[2 * pr + 1] * sq[pr];
          const float bb = dy[2 * pr + 1] * cq[pr] - dy[2 * pr] * sq[pr];
          dy[2 * pr] = a; dy[2 * pr + 1] = bb;
        }
      }
      // head-norm Jacobian; the reductions run in every lane (the four lane groups may hold normed and plain heads)
      float ssx = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) ssx = fmaf(x[q], x[q], ssx);
      const float ms = oct_sum(ssx);
      const float hs = rsqrtf(ms * (1.f / 64.f) + p.eps_head);
      float xh[8], u[8], dot = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) { xh[q] = x[q] * hs; u[q] = dy[q] * wv[q]; dot = fmaf(u[q], xh[q], dot); }
      const float mu = oct_sum(dot) * (1.f / 64.f);
      if (normed) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float tq = dy[q] * xh[q];
          waq[q] += which == 0 ? tq : 0.f;                   // (registers: no run-time indexed array)
          wak[q] += which == 0 ? 0.f : tq;
          dy[q] = hs * (u[q] - xh[q] * mu);
        }
      }
      if (i < myrows && live) {
#pragma unroll
        for (int q = 0; q < 8; ++q) cs[q] += dy[q];
        const float r = rinv[i];
        *reinterpret_cast<uint4*>(p.dqkv + (row0 + i) * N + col) =
            make_uint4(pack_bf16x2(dy[0] * r, dy[1] * r), pack_bf16x2(dy[2] * r, dy[3] * r), pack_bf16x2(dy[4] * r, dy[5] * r),
                       pack_bf16x2(dy[6] * r, dy[7] * r));
      }
    }
    {
      float* dst = s_acc + (hp / 4) * 256 + lane * 4;       // (dead lane groups of a ragged last group store their zeros)
      *reinterpret_cast<float4*>(dst) = make_float4(cs[0], cs[1], cs[2], cs[3]);
      *reinterpret_cast<float4*>(dst + 128) = make_float4(cs[4], cs[5], cs[6], cs[7]);
    }
  }
  if (p.qw != nullptr) {
    // the four lane groups hold partial sums for the same 64 dims: fold them, group 0 stores
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      waq[q] += __shfl_xor_sync(0xffffffffu, waq[q], 16); waq[q] += __shfl_xor_sync(0xffffffffu, waq[q], 8);
      wak[q] += __shfl_xor_sync(0xffffffffu, wak[q], 16); wak[q] += __shfl_xor_sync(0xffffffffu, wak[q], 8);
    }
    if (hh == 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { s_acc[NG + d0 + q] = waq[q]; s_acc[NG + 64 + d0 + q] = wak[q]; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NG; i += blockDim.x) {
    const int g = i >> 8, plane = (i >> 7) & 1, ln = (i >> 2) & 31, j = i & 3;
    const int c = g * 256 + (ln >> 3) * 64 + (ln & 7) * 8 + plane * 4 + j;
    if (c < N) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_all[static_cast<size_t>(w) * (NG + 128) + i];
      atomicAdd(p.dcvec + static_cast<size_t>(b) * N + c, t);
    }
  }
  if (p.qw != nullptr && threadIdx.x < 128) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_all[static_cast<size_t>(w) * (NG + 128) + NG + threadIdx.x];
    atomicAdd((threadIdx.x < 64 ? p.dqw : p.dkw - 64) + threadIdx.x, t);
  }
}

// The same for heads in 128-column slots (head_dim hd <= 128, LightningDiT-XL: 72); statistics run over the real head_dim
// (padding columns are zero in dqkv and raw, and are neither read nor written here); RoPE angles come from the reference's
// [T, hd] tables, staged per row slab in shared memory as (cos, sin) pairs.
struct QkvBwdWideParams {
  __nv_bfloat16* dqkv;                      // [M, 3 * QW]
  const __nv_bfloat16* raw;                 // [M, 2 * QW]
  const float* ssq;
  const float* qw; const float* kw;         // [128] zero padded, or nullptr
  const float* rope_cos; const float* rope_sin;   // [T, hd] or nullptr
  float* dcvec;                             // [B, 3 * QW]
  float* dqw; float* dkw;                   // [128] each
  int T, D, QW, hd, slots;
  float eps_row, eps_head;
};

// One CTA = a slab of 32 rows of one sample; the WARPS SPLIT THE HEADS, not the rows: warp w takes head pairs w, w + 8, ...
// (two 128-column heads per pass, 16 lanes per head, lane = 8 consecutive dims, one 16-byte access per row and array) and
// walks all 32 rows of the slab for them, four rows in flight.  A column therefore belongs to one lane of one warp: its sum
// over the slab stays in registers and leaves with one global atomic per column and CTA -- no shared-memory accumulators.
// (Round-2 first version: one warp = one head x one row at a time, 4-byte accesses, two 5-step reductions and scalar RoPE
// loads per row; it took about two thirds of the XL step's element-wise backward time.)
constexpr int kQkvWideRows = 32;
__global__ void __launch_bounds__(256, 2)
qkv_bwd_wide_kernel(const QkvBwdWideParams p) {
  __shared__ float s_rope[kQkvWideRows][128];   // (cos, sin) of the angle of dims (2a, 2a+1) at [row][2a], [row][2a+1]; identity beyond hd
  __shared__ float s_rinv[kQkvWideRows];
  __shared__ float s_w[256];                    // dq_norm.weight | dk_norm.weight partial sums of the CTA
  const int N = 3 * p.QW;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kQkvWideRows;
  const int nrows = min(kQkvWideRows, p.T - t0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row0 = static_cast<size_t>(b) * p.T + t0;
  for (int i = threadIdx.x; i < kQkvWideRows * 64; i += blockDim.x) {
    const int r = i >> 6, a2 = (i & 63) * 2;
    float cs = 1.f, sn = 0.f;
    if (p.rope_cos != nullptr && r < nrows && a2 < p.hd) {
      cs = __ldg(p.rope_cos + static_cast<size_t>(t0 + r) * p.hd + a2);
      sn = __ldg(p.rope_sin + static_cast<size_t>(t0 + r) * p.hd + a2);
    }
    s_rope[r][a2] = cs; s_rope[r][a2 + 1] = sn;
  }
  if (threadIdx.x < kQkvWideRows)
    s_rinv[threadIdx.x] = threadIdx.x < nrows ? row_rinv_g(p.ssq, row0 + threadIdx.x, p.slots, 1.f / p.D, p.eps_row) : 0.f;
  s_w[threadIdx.x] = 0.f;
  __syncthreads();
  const int hh = lane >> 4, l16 = lane & 15;
  const int d0 = 8 * l16;                                    // first of this lane's 8 dims inside a head
  const bool dims_live = d0 < p.hd;
  const int heads3 = N / 128;
  const float inv_hd = 1.f / static_cast<float>(p.hd);
  float waq[8], wak[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { waq[q] = 0.f; wak[q] = 0.f; }
  for (int hp = 2 * warp; hp < heads3; hp += 16) {
    const int hc = hp + hh;                                  // this half-warp's head
    const bool live = hc < heads3 && dims_live;
    const int col = hc * 128 + d0;
    const int which = hc < heads3 ? (hc * 128) / p.QW : 2;   // 0 q, 1 k, 2 v
    const bool normed = which < 2 && p.qw != nullptr;
    const bool roped = which < 2 && p.rope_cos != nullptr;
    float wv[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if (normed) {
      const float* wsrc = (which == 0 ? p.qw : p.kw) + d0;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wsrc)), w1 = __ldg(reinterpret_cast<const float4*>(wsrc + 4));
      wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w; wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
    }
    float cs[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) cs[q] = 0.f;
#pragma unroll 1
    for (int rb = 0; rb < kQkvWideRows; rb += 4) {
      uint4 dyw[4], xw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        dyw[i] = make_uint4(0u, 0u, 0u, 0u); xw[i] = make_uint4(0u, 0u, 0u, 0u);
        if (rb + i < nrows && live) {
          dyw[i] = *reinterpret_cast<const uint4*>(p.dqkv + (row0 + rb + i) * N + col);
          if (normed) xw[i] = *reinterpret_cast<const uint4*>(p.raw + (row0 + rb + i) * 2 * p.QW + col);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t dwd[4] = {dyw[i].x, dyw[i].y, dyw[i].z, dyw[i].w};
        const uint32_t xwd[4] = {xw[i].x, xw[i].y, xw[i].z, xw[i].w};
        float dy[8], x[8];
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {
          const float2 a = bf2_to_f2(dwd[pr]), xx = bf2_to_f2(xwd[pr]);
          dy[2 * pr] = a.x; dy[2 * pr + 1] = a.y; x[2 * pr] = xx.x; x[2 * pr + 1] = xx.y;
        }
        if (roped) {
          const float4 r0 = *reinterpret_cast<const float4*>(&s_rope[rb + i][d0]);
          const float4 r1 = *reinterpret_cast<const float4*>(&s_rope[rb + i][d0 + 4]);
          const float cq[4] = {r0.x, r0.z, r1.x, r1.z}, sq[4] = {r0.y, r0.w, r1.y, r1.w};
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) {                   // transpose of (a,b) -> (a c - b s, b c + a s)
            const float a = dy[2 * pr] * cq[pr] + dy[2 * pr + 1] * sq[pr];
            const float bb = dy[2 * pr + 1] * cq[pr] - dy[2 * pr] * sq[pr];
            dy[2 * pr] = a; dy[2 * pr + 1] = bb;
          }
        }
        // head-norm Jacobian; the reductions run in every lane (the two half-warps may hold normed and plain heads)
        float ssx = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) ssx = fmaf(x[q], x[q], ssx);
        const float ms = half_warp_sum(ssx);
        const float hs = rsqrtf(ms * inv_hd + p.eps_head);
        float xh[8], u[8], dot = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) { xh[q] = x[q] * hs; u[q] = dy[q] * wv[q]; dot = fmaf(u[q], xh[q], dot); }
        const float mu = half_warp_sum(dot) * inv_hd;
        if (normed) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float tq = dy[q] * xh[q];
            waq[q] += which == 0 ? tq : 0.f;                 // (registers: no run-time indexed array)
            wak[q] += which == 0 ? 0.f : tq;
            dy[q] = hs * (u[q] - xh[q] * mu);
          }
        }
        if (rb + i < nrows && live) {
#pragma unroll
          for (int q = 0; q < 8; ++q) cs[q] += dy[q];
          const float r = s_rinv[rb + i];
          *reinterpret_cast<uint4*>(p.dqkv + (row0 + rb + i) * N + col) =
              make_uint4(pack_bf16x2(dy[0] * r, dy[1] * r), pack_bf16x2(dy[2] * r, dy[3] * r), pack_bf16x2(dy[4] * r, dy[5] * r),
                         pack_bf16x2(dy[6] * r, dy[7] * r));
        }
      }
    }
    if (live) {
#pragma unroll
      for (int q = 0; q < 8; ++q) atomicAdd(p.dcvec + static_cast<size_t>(b) * N + col + q, cs[q]);
    }
  }
  if (p.qw != nullptr) {
    // the two half-warps hold partial sums for the same 128 dims: fold them, then across the CTA's warps in shared memory
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      waq[q] += __shfl_xor_sync(0xffffffffu, waq[q], 16);
      wak[q] += __shfl_xor_sync(0xffffffffu, wak[q], 16);
    }
    if (hh == 0 && dims_live) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { atomicAdd(&s_w[d0 + q], waq[q]); atomicAdd(&s_w[128 + d0 + q], wak[q]); }
    }
    __syncthreads();
    if ((threadIdx.x & 127) < p.hd) atomicAdd((threadIdx.x < 128 ? p.dqw : p.dkw - 128) + threadIdx.x, s_w[threadIdx.x]);
  }
}

// dst[sec*nh*hd + h*hd + d, :] = src[sec*nh*hp + h*hp + d, :]  (gradient of a head-padded weight back to the reference rows)
__global__ void unpad_heads_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int nh, int hd, int hp, int K) {
  const int r = blockIdx.x;                 // reference row in [0, 3*nh*hd)
  const int sec = r / (nh * hd), h = (r / hd) % nh, d = r % hd;
  const size_t sr = static_cast<size_t>(sec) * nh * hp + h * hp + d;
  for (int k = threadIdx.x; k < K; k += blockDim.x) dst[static_cast<size_t>(r) * K + k] = src[sr * K + k];
}

// ---------------------------------------------------------------------------------------------
// Backward through SwiGLU (swiglu_ffn.py:33-35) in the interleaved column layout of the packed w12
// (64-column groups = [32 x1 | 32 x2]):   dx1 = dh * x2 * silu'(x1),  dx2 = dh * silu(x1)
//   out: dh12 [M, 2H] bf16 = r[row] * (dx1 | dx2);   dcvec[b, 2H] += sum_t (dx1 | dx2)
// Thread = two adjacent hidden units, loops over the 64 rows of the CTA's slab (column sums stay in registers).
// grid = (ceil(H/512), ceil(T/64), B)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(__nv_bfloat16* __restrict__ dh12, float* __restrict__ dcvec, const __nv_bfloat16* __restrict__ dh,
                  const __nv_bfloat16* __restrict__ h12, const float* __restrict__ ssq, int T, int H, int D, int slots, float eps) {
  __shared__ float s_r[64];
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * 64;
  const int nrows = min(64, T - t0);
  if (threadIdx.x < nrows) s_r[threadIdx.x] = row_rinv_g(ssq, static_cast<size_t>(b) * T + t0 + threadIdx.x, slots, 1.f / D, eps);
  __syncthreads();
  const int u = blockIdx.x * 512 + threadIdx.x * 2;          // first hidden unit of this thread
  if (u >= H) return;
  const int c1 = (u >> 5) * 64 + (u & 31);                   // column of x1 (x2 at +32)
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  for (int r = 0; r < nrows; ++r) {
    const size_t row = static_cast<size_t>(b) * T + t0 + r;
    const float2 g = bf2_to_f2(*reinterpret_cast<const uint32_t*>(dh + row * H + u));
    const float2 x1 = bf2_to_f2(*reinterpret_cast<const uint32_t*>(h12 + row * 2 * H + c1));
    const float2 x2 = bf2_to_f2(*reinterpret_cast<const uint32_t*>(h12 + row * 2 * H + c1 + 32));
    const float d10 = g.x * x2.x * dsilu_f(x1.x), d11 = g.y * x2.y * dsilu_f(x1.y);
    const float d20 = g.x * silu_f(x1.x), d21 = g.y * silu_f(x1.y);
    a0 += d10; a1 += d11; b0 += d20; b1 += d21;
    const float rr = s_r[r];
    *reinterpret_cast<uint32_t*>(dh12 + row * 2 * H + c1) = pack_bf16x2(d10 * rr, d11 * rr);
    *reinterpret_cast<uint32_t*>(dh12 + row * 2 * H + c1 + 32) = pack_bf16x2(d20 * rr, d21 * rr);
  }
  float* dc = dcvec + static_cast<size_t>(b) * 2 * H;
  atomicAdd(dc + c1, a0); atomicAdd(dc + c1 + 1, a1);
  atomicAdd(dc + c1 + 32, b0); atomicAdd(dc + c1 + 33, b1);
}

// ---------------------------------------------------------------------------------------------
// Final layer (lightningdit.py:267-272 + unpatchify :376-389) backward entry:
//   dyf[row, col] = bf16(r[row] * dout[b, ch, th*p+pi, tw*p+qi])   (col = (pi*p+qi)*cout + ch; 0 for ch >= cstore)
//   dcvec[b, col] += sum_t dout(...)
// ---------------------------------------------------------------------------------------------
__global__ void final_bwd_prep_kernel(__nv_bfloat16* __restrict__ dyf, float* __restrict__ dcvec, const float* __restrict__ dout,
                                      const float* __restrict__ ssq, int B, int T, int G, int patch, int cout, int cstore, int Nf,
                                      int D, int slots, float eps) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * T * Nf) return;
  const int col = i % Nf;
  const size_t row = i / Nf;
  const int b = row / T, tok = row % T;
  const int ch = col % cout, pq = col / cout, pi = pq / patch, qi = pq % patch;
  const int th = tok / G, tw = tok % G, HW = G * patch;
  float v = 0.f;
  if (ch < cstore) v = dout[((static_cast<size_t>(b) * cstore + ch) * HW + th * patch + pi) * HW + tw * patch + qi];
  dyf[i] = __float2bfloat16(v * row_rinv_g(ssq, row, slots, 1.f / D, eps));
  if (v != 0.f) atomicAdd(dcvec + static_cast<size_t>(b) * Nf + col, v);
}

// latent [B, C, S, S] fp32 -> patch rows [B*T, C*p*p] bf16 in the Conv2d weight's (c, pi, qi) column order
__global__ void patchify_bf16_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ lat, int B, int C, int S, int p) {
  const int G = S / p, T = G * G, Kp = C * p * p;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * T * Kp) return;
  const int k = i % Kp;
  const size_t row = i / Kp;
  const int b = row / T, tok = row % T;
  const int c = k / (p * p), pi = (k / p) % p, qi = k % p;
  out[i] = __float2bfloat16(lat[((static_cast<size_t>(b) * C + c) * S + (tok / G) * p + pi) * S + (tok % G) * p + qi]);
}

// ---------------------------------------------------------------------------------------------
// Small reductions / fp32 matrix products of the conditioning backward (batch-sized: negligible work)
// ---------------------------------------------------------------------------------------------
// out[n] (+)= sum_b in[b*ld_in + n] * (mul ? mul[b*ld_mul + n] : 1)
__global__ void colsum_kernel(float* __restrict__ out, const float* __restrict__ in, int ld_in, const float* __restrict__ mul,
                              int ld_mul, int B, int N, int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += in[static_cast<size_t>(b) * ld_in + n] * (mul ? mul[static_cast<size_t>(b) * ld_mul + n] : 1.f);
  out[n] = accumulate ? out[n] + s : s;
}
// adaLN slot gradients: dmods[b, scale_off[s] + d] = dg[s][b][d] * norm_w[s][d];   dnorm_w[s][d] = sum_b dg * (1 + scale)
__global__ void adaln_bwd_kernel(float* __restrict__ dmods, float* __restrict__ dnorm_w, const float* __restrict__ dg,
                                 const float* __restrict__ mods, const float* __restrict__ norm_w,
                                 const int* __restrict__ slot_scale_off, int B, int D, int Ntot, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * D) return;
  const int s = i / D, d = i % D;
  const int so = slot_scale_off[s];
  const float w = norm_w[i];
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const float v = dg[(static_cast<size_t>(s) * B + b) * D + d];
    dmods[static_cast<size_t>(b) * Ntot + so + d] = v * w;
    acc = fmaf(v, 1.f + mods[static_cast<size_t>(b) * Ntot + so + d], acc);
  }
  dnorm_w[i] = acc;
}
// out[i] = in[i] * silu'(pre[i])
__global__ void dsilu_mul_kernel(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ pre, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * dsilu_f(pre[i]);
}
__global__ void silu_f32_kernel(float* __restrict__ out, const float* __restrict__ in, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = silu_f(in[i]);
}
// sinusoidal timestep features (lightningdit.py:108-128): out[b, 0:half] = cos(t f), [half:] = sin(t f)
__global__ void timestep_features_kernel(float* __restrict__ out, const float* __restrict__ tvals, float tscalar, int B, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  const int b = i / K, k = i % K, half = K / 2;
  const float t = tvals ? tvals[b] : tscalar;
  const float a = t * expf(-9.210340371976184f * static_cast<float>(k % half) / static_cast<float>(half));
  out[i] = (k < half) ? cosf(a) : sinf(a);
}
// C[n, k] += sum_b A[b, n] * Bm[b, k]   (fp32; one thread per output)
__global__ void small_wgrad_kernel(float* __restrict__ C, const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                                   int B, int N, int K) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(N) * K) return;
  const int n = i / K, k = i % K;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s = fmaf(A[static_cast<size_t>(b) * lda + n], Bm[static_cast<size_t>(b) * ldb + k], s);
  C[i] += s;
}
// out[b, k] (row pitch ld_out) = sum_n A[b, n] * W[n, k]   (fp32)
__global__ void small_dgrad_kernel(float* __restrict__ out, int ld_out, const float* __restrict__ A, int lda,
                                   const float* __restrict__ W, int B, int N, int K) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * K) return;
  const int b = i / K, k = i % K;
  float s = 0.f;
  for (int n = 0; n < N; ++n) s = fmaf(A[static_cast<size_t>(b) * lda + n], W[static_cast<size_t>(n) * K + k], s);
  out[static_cast<size_t>(b) * ld_out + k] = s;
}
// table[idx[b], :] += v[b, :]   (rows outside the table were flagged by the forward's gather and are skipped here)
__global__ void embed_bwd_kernel(float* __restrict__ table, const long long* __restrict__ idx, const float* __restrict__ v, int B, int D,
                                 int table_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const long long r = idx[i / D];
  if (r < 0 || r >= table_rows) return;
  atomicAdd(table + static_cast<size_t>(r) * D + (i % D), v[i]);
}
// bf16 matrix transpose: dst[c, r] = src[r, c]  (src [R, C]); 32 x 32 tiles through shared memory
__global__ void transpose_bf16_kernel(__nv_bfloat16* __restrict__ dst, const __nv_bfloat16* __restrict__ src, int R, int C) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < R && c < C) ? src[static_cast<size_t>(r) * C + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < C && r < R) dst[static_cast<size_t>(c) * R + r] = tile[threadIdx.x][j];
  }
}
__global__ void f32_to_bf16_ld_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ in, int rows, int cols, int ld_in) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(rows) * cols) return;
  out[i] = __float2bfloat16(in[(i / cols) * ld_in + (i % cols)]);
}

// ---------------------------------------------------------------------------------------------
// Fused AdamW + EMA over a flat parameter buffer (reference train_accum.py:121,240-246,337-347:
// torch.optim.AdamW(lr, betas, weight_decay) followed by update_ema(ema, model, decay)).  One pass:
// 5 reads + 4 writes of 4 bytes per parameter.  grad_scale folds the 1/world_size of the gradient all-reduce.
// ---------------------------------------------------------------------------------------------
__global__ void adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 float* __restrict__ ema, size_t n, float lr, float beta1, float beta2, float eps, float wd,
                                 float bc1, float bc2, float ema_decay, float grad_scale) {
  const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    float4 pv = *reinterpret_cast<float4*>(p + i);
    const float4 gv = *reinterpret_cast<const float4*>(g + i);
    float4 mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
    float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w}, v2[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float gq = gg[q] * grad_scale;
      pp[q] *= 1.f - lr * wd;
      mm[q] = beta1 * mm[q] + (1.f - beta1) * gq;
      v2[q] = beta2 * v2[q] + (1.f - beta2) * gq * gq;
      pp[q] -= lr / bc1 * mm[q] / (sqrtf(v2[q]) / sqrtf(bc2) + eps);
    }
    *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(v2[0], v2[1], v2[2], v2[3]);
    if (ema) {
      float4 ev = *reinterpret_cast<float4*>(ema + i);
      ev.x = ev.x * ema_decay + pp[0] * (1.f - ema_decay); ev.y = ev.y * ema_decay + pp[1] * (1.f - ema_decay);
      ev.z = ev.z * ema_decay + pp[2] * (1.f - ema_decay); ev.w = ev.w * ema_decay + pp[3] * (1.f - ema_decay);
      *reinterpret_cast<float4*>(ema + i) = ev;
    }
  } else {
    for (size_t j = i; j < n; ++j) {
      const float gq = g[j] * grad_scale;
      float pj = p[j] * (1.f - lr * wd);
      m[j] = beta1 * m[j] + (1.f - beta1) * gq;
      v[j] = beta2 * v[j] + (1.f - beta2) * gq * gq;
      pj -= lr / bc1 * m[j] / (sqrtf(v[j]) / sqrtf(bc2) + eps);
      p[j] = pj;
      if (ema) ema[j] = ema[j] * ema_decay + pj * (1.f - ema_decay);
    }
  }
}

}  // namespace ldmae
