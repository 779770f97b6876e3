// Host-side plumbing shared by the library translation units: error reporting across the C ABI,
// device buffers, TMA tensor-map construction and the GEMM launcher.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>

#include "../../include/ldmae_b200.h"
#include "gemm_sm100.cuh"

namespace ldmae {

std::string& last_error();
int set_error(int code, const char* fmt, ...);

#define LDMAE_CUDA(expr)                                                                              \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return ::ldmae::set_error(LDMAE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define LDMAE_TRY(expr)          \
  do {                           \
    int _r = (expr);             \
    if (_r != LDMAE_OK) return _r; \
  } while (0)
#define LDMAE_REQUIRE(cond, ...)                                    \
  do {                                                              \
    if (!(cond)) return ::ldmae::set_error(LDMAE_ERR_INVALID, __VA_ARGS__); \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  int alloc(size_t count, bool zero = false) {
    release();
    if (count == 0) return LDMAE_OK;
    LDMAE_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
    n = count;
    if (zero) LDMAE_CUDA(cudaMemset(p, 0, count * sizeof(T)));
    return LDMAE_OK;
  }
};

int device_sm_count();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: one flag per (call site, device ordinal), so a second
// GPU driven from the same process gets its own opt-in.
struct PerDeviceOnce {
  bool done_[64] = {};
  static int dev() { int d = 0; cudaGetDevice(&d); return d & 63; }
  bool pending() const { return !done_[dev()]; }
  void mark() { done_[dev()] = true; }
};

// 2-D row-major [rows, cols] (leading dimension ld elements) tensor map with a {box_cols x box_rows} box.
// elem_bytes 2 = bf16, 4 = fp32; swizzle_bytes 128 / 64 (box_cols * elem_bytes must not exceed it).  Cached.
int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, int rows, int cols, int ld, int box_cols, int box_rows,
                 int swizzle_bytes);
// bf16 operand tiles: {64 x box_rows} box, 128-byte swizzle
inline int make_tmap_bf16(CUtensorMap* out, const void* ptr, int rows, int cols, int ld, int box_rows) {
  return make_tmap_2d(out, ptr, 2, rows, cols, ld, 64, box_rows, 128);
}
// epilogue staging tiles (32 rows x 128 bytes, 128-byte swizzle)
inline int make_tmap_out_bf16(CUtensorMap* out, const void* ptr, int rows, int cols, int ld) {
  return make_tmap_2d(out, ptr, 2, rows, cols, ld, 64, 32, 128);
}
inline int make_tmap_out_f32(CUtensorMap* out, const void* ptr, int rows, int cols, int ld) {
  return make_tmap_2d(out, ptr, 4, rows, cols, ld, 32, 32, 128);
}

template <int BN, int CG, class Epi>
int launch_gemm(const void* a, int lda, const void* w, int ldw, GemmShape g, const typename Epi::Params& ep,
                cudaStream_t st) {
  using Cfg = GemmCfg<BN, CG, Epi>;
  LDMAE_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty shape %d %d %d", g.M, g.N, g.K);
  LDMAE_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "gemm: leading dimensions must be multiples of 8 (16-byte TMA strides)");
  CUtensorMap ta, tw;
  LDMAE_TRY(make_tmap_bf16(&ta, a, g.M, g.K, lda, kBM));
  LDMAE_REQUIRE(g.n_live == 0 || (g.n_live % 16 == 0 && g.n_live >= 16 && g.n_live <= Cfg::kLoadBN),
                "gemm: n_live = %d must be a multiple of 16 up to %d", g.n_live, Cfg::kLoadBN);
  LDMAE_TRY(make_tmap_bf16(&tw, w, g.N, g.K, ldw, g.n_live > 0 ? g.n_live : Cfg::kLoadBN));
  auto kern = gemm_tn_kernel<BN, CG, Epi>;
  static PerDeviceOnce attr_set;
  if (attr_set.pending()) {
    LDMAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set.mark();
  }
  const int n_tiles = (g.N + BN - 1) / BN;
  const int m_tiles = (g.M + kBM * CG - 1) / (kBM * CG);
  const long long total = static_cast<long long>(n_tiles) * m_tiles;
  int sms = device_sm_count();
  long long clusters = sms / CG;
  if (total < clusters) clusters = total;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * CG));
  cfg.blockDim = dim3(gemm_threads<Epi>());
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LDMAE_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tw, g, ep));
  return LDMAE_OK;
}

// counts kernel launches issued by this library (bench.py reports it as gpu_launches)
extern long long g_launch_count;

}  // namespace ldmae
