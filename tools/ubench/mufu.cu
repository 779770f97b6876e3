// Micro-benchmark: MUFU.EX2 throughput per SM (fp32, f16x2, bf16x2) and FMA-pipe polynomial exp2.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  unsigned int h[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xBC00BC00u + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) {  // FMA-pipe emulation: 2^x = 2^floor(x) * p(frac), degree-3 polynomial
        float x = a[i];
        float t = x + 12582912.f;
        float fl = t - 12582912.f;
        float f = x - fl;
        float p = fmaf(fmaf(fmaf(0.0555041f, f, 0.2402265f), f, 0.6931472f), f, 1.0f);
        a[i] = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23)) - 1.0f;
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  if (s == 12345.f) out[0] = s;
}
template <int MODE>
void run(const char* name, int threads) {
  float* d; cudaMalloc(&d, 4);
  int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, threads>>>(d, 16);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ops = (double)threads * iters * 8;           // per SM (lane-ops; x2 elements for packed modes)
  printf("%-10s threads/SM %4d: %.3f ms  -> %.2f lane-instr/clk/SM (at %d MHz nominal)\n", name, threads, ms, ops / (ms * 1e-3) / (clk * 1e3), clk / 1000);
  cudaFree(d);
}
int main() {
  for (int th : {32, 128, 256, 512, 1024}) {
    run<0>("ex2.f32", th); run<1>("ex2.f16x2", th); run<2>("ex2.bf16x2", th); run<3>("poly3", th);
  }
  return 0;
}
