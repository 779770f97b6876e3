// HBM-bound kernels either side of the training forward/backward: the trainer's input pipeline and the loss.
//   flow_prepare_kernel   posterior sample -> per-channel normalise -> flow-matching pair in ONE pass
//                         (reference datasets/img_latent_dataset.py:76-94 + tokenizer/util/misc.py:74-96 +
//                          transport/transport.py:136-166 + transport/path.py:114-136)
//   flow_loss_kernel      loss[b] = mean_flat((out - ut)^2) and dout = d(mean_b(loss) * loss_scale) / d out
//                         (transport.py:195, train_accum.py:220-223)
// 128-bit vectorised, coalesced; the loss reduction is a fixed-order block reduction (bit-reproducible).
#pragma once
#include "elementwise.cuh"

namespace ldmae {

// One thread = 4 consecutive pixels of one (sample, channel) plane.  HW must be a multiple of 4.
//   moments != nullptr: z = mu + exp(0.5 * clamp(logvar, -30, 20)) * eps   (eps == nullptr: z = mu, the posterior mode),
//                       rows taken from `moments_flip` where flip[b] != 0 (the dataset's coin flip);
//                       x1 = (z - mean[c]) / std[c] * multiplier            (mean == nullptr: no normalisation)
//   moments == nullptr: x1 = x1_in (already normalised latents)
//   xt = t[b] * x1 + (1 - t[b]) * x0 ;  ut = x1 - x0
__global__ void __launch_bounds__(256)
flow_prepare_kernel(const float* __restrict__ moments, const float* __restrict__ moments_flip, const unsigned char* __restrict__ flip,
                    const float* __restrict__ eps_post, const float* __restrict__ mean, const float* __restrict__ stdv, float multiplier,
                    const float* __restrict__ x1_in, const float* __restrict__ x0, const float* __restrict__ t,
                    float* __restrict__ x1_out, float* __restrict__ xt, float* __restrict__ ut, int B, int C, int HW) {
  const size_t i4 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total4 = static_cast<size_t>(B) * C * HW / 4;
  if (i4 >= total4) return;
  const size_t i = i4 * 4;
  const int p = static_cast<int>(i % HW);
  const int c = static_cast<int>((i / HW) % C);
  const size_t b = i / (static_cast<size_t>(HW) * C);
  float4 x1;
  if (moments != nullptr) {
    const float* src = (flip != nullptr && flip[b] != 0 && moments_flip != nullptr) ? moments_flip : moments;
    const float* row = src + (b * 2 * C + c) * HW + p;
    const float4 mu = *reinterpret_cast<const float4*>(row);
    x1 = mu;
    if (eps_post != nullptr) {
      const float4 lv = *reinterpret_cast<const float4*>(row + static_cast<size_t>(C) * HW);
      const float4 e = *reinterpret_cast<const float4*>(eps_post + i);
      // two roundings (multiply, then add) as the reference's eager `mean + std * sample`
      x1.x = __fadd_rn(mu.x, __fmul_rn(expf(0.5f * fminf(fmaxf(lv.x, -30.f), 20.f)), e.x));
      x1.y = __fadd_rn(mu.y, __fmul_rn(expf(0.5f * fminf(fmaxf(lv.y, -30.f), 20.f)), e.y));
      x1.z = __fadd_rn(mu.z, __fmul_rn(expf(0.5f * fminf(fmaxf(lv.z, -30.f), 20.f)), e.z));
      x1.w = __fadd_rn(mu.w, __fmul_rn(expf(0.5f * fminf(fmaxf(lv.w, -30.f), 20.f)), e.w));
    }
    if (mean != nullptr) {
      const float m = mean[c], s = stdv[c];
      x1.x = __fdiv_rn(x1.x - m, s); x1.y = __fdiv_rn(x1.y - m, s); x1.z = __fdiv_rn(x1.z - m, s); x1.w = __fdiv_rn(x1.w - m, s);
    }
    x1.x = __fmul_rn(x1.x, multiplier); x1.y = __fmul_rn(x1.y, multiplier);
    x1.z = __fmul_rn(x1.z, multiplier); x1.w = __fmul_rn(x1.w, multiplier);
  } else {
    x1 = *reinterpret_cast<const float4*>(x1_in + i);
  }
  const float4 n0 = *reinterpret_cast<const float4*>(x0 + i);
  const float tb = t[b], omt = 1.f - tb;
  float4 a, u;
  // path.py:133-136: alpha_t * x1 + sigma_t * x0 with alpha = t, sigma = 1 - t  (products rounded, then summed)
  a.x = __fadd_rn(__fmul_rn(tb, x1.x), __fmul_rn(omt, n0.x)); a.y = __fadd_rn(__fmul_rn(tb, x1.y), __fmul_rn(omt, n0.y));
  a.z = __fadd_rn(__fmul_rn(tb, x1.z), __fmul_rn(omt, n0.z)); a.w = __fadd_rn(__fmul_rn(tb, x1.w), __fmul_rn(omt, n0.w));
  u.x = x1.x - n0.x; u.y = x1.y - n0.y; u.z = x1.z - n0.z; u.w = x1.w - n0.w;
  *reinterpret_cast<float4*>(xt + i) = a;
  *reinterpret_cast<float4*>(ut + i) = u;
  if (x1_out != nullptr) *reinterpret_cast<float4*>(x1_out + i) = x1;
}

// One CTA = one sample (n = C*H*W elements, multiple of 4).  dout may be nullptr (evaluation).
__global__ void __launch_bounds__(256)
flow_loss_kernel(const float* __restrict__ out, const float* __restrict__ ut, float* __restrict__ loss, float* __restrict__ dout,
                 float dscale /* 2 * loss_scale / (n * B) */, int n) {
  __shared__ float part[8];
  const size_t base = static_cast<size_t>(blockIdx.x) * n;
  float s = 0.f;
  for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) {
    const float4 o = *reinterpret_cast<const float4*>(out + base + i);
    const float4 u = *reinterpret_cast<const float4*>(ut + base + i);
    const float4 d = make_float4(o.x - u.x, o.y - u.y, o.z - u.z, o.w - u.w);
    s = fmaf(d.x, d.x, s); s = fmaf(d.y, d.y, s); s = fmaf(d.z, d.z, s); s = fmaf(d.w, d.w, s);
    if (dout != nullptr)
      *reinterpret_cast<float4*>(dout + base + i) = make_float4(d.x * dscale, d.y * dscale, d.z * dscale, d.w * dscale);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += part[w];
    loss[blockIdx.x] = tot / static_cast<float>(n);
  }
}

}  // namespace ldmae
