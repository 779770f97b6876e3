// libldmae_b200: C ABI (include/ldmae_b200.h) over the sm_100a kernels.
//   * ldmae_dit_*      LightningDiT forward / forward_with_cfg (reference models/lightningdit.py:391-442)
//   * ldmae_sample_ode transport ODE sampler loop (reference transport/integrators.py:77-126)
//   * ldmae_vmae_*     VMAE ViT decoder (reference tokenizer/models_mae.py:865-887,963-973)
// Host orchestration only: every FLOP and byte of the path runs in the CUDA kernels of this directory.
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "attention_bwd_sm100.cuh"
#include "attention_hd128_sm100.cuh"
#include "attention_persist_sm100.cuh"
#include "attention_sm100.cuh"
#include "common.cuh"
#include "gemm_wgrad_sm100.cuh"
#include "backward_elementwise.cuh"
#include "elementwise.cuh"
#include "train_io.cuh"

namespace ldmae {

// ------------------------------------------------------------------------------------------- plumbing
std::string& last_error() {
  static thread_local std::string e;
  return e;
}
int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}
long long g_launch_count = 0;
static long long* g_gemm_trace = nullptr;   // debug builds (-DLDMAE_GEMM_TRACE): residual-epilogue clock stamps

// Optional per-kernel-class device timing (CUDA events on the launching stream), used by bench.py for the
// roofline of the dominant kernel.  Classes: 0 qkv GEMM, 1 attention, 2 proj GEMM, 3 w12 (SwiGLU) GEMM,
// 4 w3 GEMM, 5 adaLN / shift-vector GEMMs, 6 final-layer GEMM, 7 conditioning + patch embed + ODE update, 8 VMAE decode.

struct ProfRec { int cls; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
struct ProfScope {
  cudaStream_t st; int cls; cudaEvent_t a = nullptr, b = nullptr; bool on;
  ProfScope(int c, cudaStream_t s) : st(s), cls(c), on(g_prof_on) {
    if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
  }
  ~ProfScope() {
    if (on) { cudaEventRecord(b, st); g_prof.push_back(ProfRec{cls, a, b}); }
  }
};

int device_sm_count() {
  static int sms[64] = {};       // per device ordinal
  int dev = 0;
  cudaGetDevice(&dev);
  int& s = sms[dev & 63];
  if (s == 0) {
    cudaDeviceGetAttribute(&s, cudaDevAttrMultiProcessorCount, dev);
    if (s <= 0) s = 148;
  }
  return s;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, int rows, int cols, int ld, int box_cols, int box_rows,
                 int swizzle_bytes) {
  typedef std::tuple<const void*, int, int, int, int, int, int, int> Key;
  static thread_local std::map<Key, CUtensorMap> cache;
  Key key(ptr, elem_bytes, rows, cols, ld, box_cols, box_rows, swizzle_bytes);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return LDMAE_OK;
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(LDMAE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  LDMAE_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "tensor map: base pointer must be 16-byte aligned");
  LDMAE_REQUIRE((static_cast<long long>(ld) * elem_bytes) % 16 == 0, "tensor map: row pitch %d x %d bytes is not a multiple of 16", ld, elem_bytes);
  LDMAE_REQUIRE(box_cols * elem_bytes <= swizzle_bytes, "tensor map: box of %d bytes exceeds the %d-byte swizzle span", box_cols * elem_bytes, swizzle_bytes);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LDMAE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box=%dx%d elem=%d", (int)r, rows,
                     cols, ld, box_cols, box_rows, elem_bytes);
  if (cache.size() > 8192) cache.clear();
  cache[key] = *out;
  return LDMAE_OK;
}

static inline unsigned cdiv(size_t a, size_t b) { return static_cast<unsigned>((a + b - 1) / b); }
#define LDMAE_LAUNCH_CHECK()                                                                      \
  do {                                                                                            \
    ++::ldmae::g_launch_count;                                                                    \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess)                                                                        \
      return ::ldmae::set_error(LDMAE_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// GEMM tile configuration: big token GEMMs use CTA pairs with 256x256 cluster tiles; LDMAE_GEMM_CG=1
// switches the whole library to single-CTA 128x128 tiles (debug / A-B measurements).
static int gemm_cg() {
  static int cg = 0;
  if (cg == 0) {
    const char* e = getenv("LDMAE_GEMM_CG");
    cg = (e && atoi(e) == 1) ? 1 : 2;
  }
  return cg;
}
template <class Epi>
static int gemm_auto(const void* a, int lda, const void* w, int ldw, GemmShape g, const typename Epi::Params& ep,
                     cudaStream_t st) {
  ++g_launch_count;
  if (gemm_cg() == 2 && g.M > 128) return launch_gemm<256, 2, Epi>(a, lda, w, ldw, g, ep, st);
  return launch_gemm<128, 1, Epi>(a, lda, w, ldw, g, ep, st);
}

// out[M,N] = act(a . w^T + bias) through the TMA-store epilogue (out leading dimension ldo)
template <typename OutT, int ACT>
static int gemm_store(const void* a, int lda, const void* w, int ldw, GemmShape g, OutT* out, int ldo, const float* bias,
                      cudaStream_t st) {
  typename EpiStore<OutT, ACT>::Params ep;
  if (sizeof(OutT) == 2) LDMAE_TRY(make_tmap_out_bf16(&ep.omap, out, g.M, g.N, ldo));
  else LDMAE_TRY(make_tmap_out_f32(&ep.omap, out, g.M, g.N, ldo));
  ep.bias = bias;
  return gemm_auto<EpiStore<OutT, ACT>>(a, lda, w, ldw, g, ep, st);
}
// x[M,N] (fp32, ld ldx) += gate * (a . w^T + bias); optionally anext = bf16(x * gnext) and row statistics (see EpiResidual)
static int gemm_residual(const void* a, int lda, const void* w, int ldw, GemmShape g, float* x, int ldx, const float* bias,
                         const float* gate, int gate_ld, const float* gnext, int gnext_ld, __nv_bfloat16* anext, float* ssq,
                         int ss_slots, int rows_per_sample, cudaStream_t st) {
  typename EpiResidual::Params ep;
  LDMAE_REQUIRE(g.N % 64 == 0, "residual GEMM: N = %d must be a multiple of 64 (the epilogue walks the tile in 64-column pairs)", g.N);
  LDMAE_TRY(make_tmap_out_f32(&ep.xmap, x, g.M, g.N, ldx));
  ep.xmap_out = ep.xmap;
  ep.mmap = ep.xmap;
  ep.has_anext = anext != nullptr;
  if (anext) LDMAE_TRY(make_tmap_out_bf16(&ep.amap, anext, g.M, g.N, ldx));
  else ep.amap = ep.xmap;
  ep.bias = bias; ep.gate = gate; ep.gnext = gnext; ep.ssq = ssq;
  ep.gate_ld = gate_ld; ep.gnext_ld = gnext_ld; ep.rows_per_sample = rows_per_sample; ep.ss_slots = ss_slots;
  ep.x_row_mod = 0;
  ep.trace = g_gemm_trace;
  static int deep = -1;
  if (deep < 0) { const char* e = getenv("LDMAE_RESID_DEEP"); deep = e ? atoi(e) : 1; }
  if (((deep == 1 && g.K <= 1024) || deep == 2) && g.M > 128) {
    // short K: the epilogue's residual traffic, not the tensor core, bounds the kernel
    EpiResidualDeep::Params ed;
    memcpy(&ed, &ep, sizeof ed);
    return gemm_auto<EpiResidualDeep>(a, lda, w, ldw, g, ed, st);
  }
  return gemm_auto<EpiResidual>(a, lda, w, ldw, g, ep, st);
}

// Training forward of the same epilogue: x_out (a different buffer) = x_in + gate * m, and m = a . w^T + bias is kept (bf16).
static int gemm_residual_train(const void* a, int lda, const void* w, int ldw, GemmShape g, const float* x_in, float* x_out,
                               int ldx, const float* bias, const float* gate, int gate_ld, const float* gnext, int gnext_ld,
                               __nv_bfloat16* anext, float* ssq, int ss_slots, __nv_bfloat16* m_out, int rows_per_sample,
                               cudaStream_t st) {
  typename EpiResidualTrain::Params ep;
  LDMAE_REQUIRE(g.N % 64 == 0, "residual GEMM: N = %d must be a multiple of 64 (the epilogue walks the tile in 64-column pairs)", g.N);
  LDMAE_TRY(make_tmap_out_f32(&ep.xmap, x_in, g.M, g.N, ldx));
  LDMAE_TRY(make_tmap_out_f32(&ep.xmap_out, x_out, g.M, g.N, ldx));
  LDMAE_TRY(make_tmap_out_bf16(&ep.mmap, m_out, g.M, g.N, ldx));
  ep.has_anext = anext != nullptr;
  if (anext) LDMAE_TRY(make_tmap_out_bf16(&ep.amap, anext, g.M, g.N, ldx));
  else ep.amap = ep.mmap;
  ep.bias = bias; ep.gate = gate; ep.gnext = gnext; ep.ssq = ssq;
  ep.gate_ld = gate_ld; ep.gnext_ld = gnext_ld; ep.rows_per_sample = rows_per_sample; ep.ss_slots = ss_slots;
  ep.x_row_mod = 0;
  ep.trace = nullptr;
  return gemm_auto<EpiResidualTrain>(a, lda, w, ldw, g, ep, st);
}

extern "C" int ldmae_gemm_trace(long long* dev_buf) { g_gemm_trace = dev_buf; return LDMAE_OK; }
static long long* g_attn_trace = nullptr;   // debug builds (-DLDMAE_ATTN_TRACE): device buffer [2][64][8] of clock64 stamps
extern "C" int ldmae_attention_trace(long long* dev_buf) { g_attn_trace = dev_buf; return LDMAE_OK; }

// m0_log2 > 0: the scores are bounded (|s * scale * log2e| <= m0_log2): constant-offset softmax instantiation
static int run_attention(const void* qkv, int ldq, void* out, int ldo, int B, int T, int H, int q_col, int k_col,
                         int v_col, float scale, cudaStream_t st, float* lse2 = nullptr, float m0_log2 = -1.f,
                         bool prescaled = false) {
  CUtensorMap tm, tmo;
  LDMAE_TRY(make_tmap_bf16(&tm, qkv, B * T, ldq, ldq, 128));
  LDMAE_TRY(make_tmap_out_bf16(&tmo, out, B * T, H * 64, ldo));
  static PerDeviceOnce attr;
  if (attr.pending()) {
    LDMAE_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    LDMAE_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    LDMAE_CUDA((cudaFuncSetAttribute(attn_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes)));
    LDMAE_CUDA(cudaFuncSetAttribute(attn_fwd_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnPSmemBytes));
    LDMAE_CUDA(cudaFuncSetAttribute(attn_fwd_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnPSmemBytes));
    LDMAE_CUDA((cudaFuncSetAttribute(attn_fwd_persist_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnPSmemBytes)));
    LDMAE_CUDA((cudaFuncSetAttribute(attn_fwd_persist_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnPSmemBytes)));
    attr.mark();
  }
  AttnParams p;
  p.trace = g_attn_trace;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse2 = lse2;
  p.T = T; p.H = H; p.ldo = ldo;
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  // prescaled: q already carries scale * log2(e) (EpiQKV::Params::q_mul), every kernel sees a unit scale
  p.scale_log2 = prescaled ? 1.f : scale * 1.4426950408889634f;
  p.m0_log2 = m0_log2;
  p.raw = prescaled ? 1 : 0;
  // measured (B=128, T=1024, H=12; tools/bench_attn.py): taking turns helps the tracking instantiation (561 -> 594 TFLOP/s)
  // and is neutral-to-negative for the constant-offset one (605 vs 599); LDMAE_ATTN_ALTERNATE=0/1 forces either
  static int alt = -2;
  if (alt == -2) { const char* e = getenv("LDMAE_ATTN_ALTERNATE"); alt = e ? atoi(e) : -1; }
  p.alternate = alt >= 0 ? alt : (m0_log2 > 0.f ? 0 : 1);
  dim3 grid(cdiv(T, 256), H, B);
  static int wide = -1;
  if (wide < 0) { const char* e = getenv("LDMAE_ATTN_WIDE"); wide = e ? atoi(e) : 0; }
  // persistent scheduling: measured +6 % for the constant-offset instantiation (606 -> 644 TFLOP/s, 49.5 -> 46.2 ms per
  // sampling step of batch 64) and -2.5 % for the tracking one, which therefore keeps one CTA per work item
  // (LDMAE_ATTN_PERSIST=0 / 2 forces never / always)
  static int persist = -1;
  if (persist < 0) { const char* e = getenv("LDMAE_ATTN_PERSIST"); persist = e ? atoi(e) : 1; }
  if (!wide && (persist == 2 || (persist == 1 && m0_log2 > 0.f))) {
    const int n_qpairs = static_cast<int>(cdiv(T, 256));
    const long long items = static_cast<long long>(n_qpairs) * H * B;
    LDMAE_REQUIRE(items < (1ll << 31), "attention: too many work items");
    const unsigned ctas = static_cast<unsigned>(std::min<long long>(items, device_sm_count()));
    // software-pipelined softmax warps (attention_persist_sm100.cuh, kPipe): 644 -> 765 TFLOP/s standalone;
    // LDMAE_ATTN_PIPE=0 keeps the phase-by-phase loop.  Needs whole key blocks and at least two per item: with a single block
    // per item the softmax warps could run a whole item ahead of the epilogue warps and overrun the l_full barrier's parity.
    static int pipe = -1;
    if (pipe < 0) { const char* e = getenv("LDMAE_ATTN_PIPE"); pipe = e ? atoi(e) : 1; }
    if (m0_log2 > 0.f && pipe && T % 128 == 0 && T >= 256 && prescaled) attn_fwd_persist_kernel<true, true, true><<<ctas, kAttnPipeThreads, kAttnPSmemBytes, st>>>(tm, tmo, p, n_qpairs, static_cast<int>(items));
    else if (m0_log2 > 0.f && pipe && T % 128 == 0 && T >= 256) attn_fwd_persist_kernel<true, true><<<ctas, kAttnPipeThreads, kAttnPSmemBytes, st>>>(tm, tmo, p, n_qpairs, static_cast<int>(items));
    else if (m0_log2 > 0.f) attn_fwd_persist_kernel<true><<<ctas, kAttnThreads, kAttnPSmemBytes, st>>>(tm, tmo, p, n_qpairs, static_cast<int>(items));
    else attn_fwd_persist_kernel<false><<<ctas, kAttnThreads, kAttnPSmemBytes, st>>>(tm, tmo, p, n_qpairs, static_cast<int>(items));
  }
  else if (m0_log2 > 0.f && wide) attn_fwd_kernel<true, true><<<grid, 640, kAttnSmemBytes, st>>>(tm, tmo, p);
  else if (m0_log2 > 0.f) attn_fwd_kernel<true><<<grid, kAttnThreads, kAttnSmemBytes, st>>>(tm, tmo, p);
  else attn_fwd_kernel<false><<<grid, kAttnThreads, kAttnSmemBytes, st>>>(tm, tmo, p);
  LDMAE_LAUNCH_CHECK();
  return LDMAE_OK;
}

// Heads stored with a 128-column stride (real head_dim hd <= 128): attention_hd128_sm100.cuh
static int run_attention_hd128(const void* qkv, int ldq, void* out, int ldo, int B, int T, int H, int hd, int q_col, int k_col,
                               int v_col, float scale, cudaStream_t st, float* lse2 = nullptr, float m0_log2 = -1.f) {
  LDMAE_REQUIRE(hd > 0 && hd <= 128 && hd % 8 == 0, "wide attention: head_dim %d must be a multiple of 8 up to 128", hd);
  CUtensorMap tm;
  LDMAE_TRY(make_tmap_bf16(&tm, qkv, B * T, ldq, ldq, 128));
  static PerDeviceOnce attr;
  if (attr.pending()) {
    LDMAE_CUDA(cudaFuncSetAttribute(attn_fwd_hd128_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA128SmemBytes));
    LDMAE_CUDA(cudaFuncSetAttribute(attn_fwd_hd128_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA128SmemBytes));
    attr.mark();
  }
  Attn128Params p;
  p.out = static_cast<__nv_bfloat16*>(out); p.lse2 = lse2;
  p.T = T; p.H = H; p.ldo = ldo; p.hd = hd; p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.m0_log2 = m0_log2;
  static int poly = -1;
  // measured on XL/1 @512 (round 2): 741 vs 737 TFLOP/s with / without the polynomial quarter -- inside the noise, so off
  if (poly < 0) { const char* e = getenv("LDMAE_ATTN_WIDE_POLY"); poly = e ? atoi(e) : 0; }
  p.poly = poly;
  dim3 grid(cdiv(T, 128), H, B);
  static int half = -1;
  if (half < 0) { const char* e = getenv("LDMAE_ATTN_WIDE_HALF"); half = e ? atoi(e) : 1; }
  if (m0_log2 > 0.f && half) attn_fwd_hd128_kernel<true><<<grid, kA128HalfThreads, kA128SmemBytes, st>>>(tm, p);
  else attn_fwd_hd128_kernel<false><<<grid, kA128Threads, kA128SmemBytes, st>>>(tm, p);
  LDMAE_LAUNCH_CHECK();
  return LDMAE_OK;
}

// Backward of run_attention / run_attention_hd128: dqkv [B*T, ldq] (same column layout as qkv, head slots of HW columns)
// from dO [B*T, H*hd] (dense), the forward output o, the forward's lse2 [B,H,T] (padded by 64 floats) and a statistics
// workspace of 2 * (B*H*T + 64) floats.
template <int HW>
static int run_attention_bwd_t(const void* qkv, int ldq, const void* o, const void* d_o, int ldo, const float* lse2, float* delta,
                               void* dqkv, int B, int T, int H, int hd, int q_col, int k_col, int v_col, float scale,
                               cudaStream_t st) {
  LDMAE_REQUIRE(ldo == H * hd, "attention backward expects a dense [B*T, H*head_dim] output gradient");
  LDMAE_REQUIRE(T % 4 == 0, "attention backward: T must be a multiple of 4 (16-byte aligned statistics rows)");
  LDMAE_REQUIRE(hd <= HW && hd % 8 == 0, "attention backward: head_dim %d does not fit the %d-column head slot", hd, HW);
  CUtensorMap tqr, tqc, tdr, tdc;
  LDMAE_TRY(make_tmap_bf16(&tqr, qkv, B * T, ldq, ldq, 128));
  LDMAE_TRY(make_tmap_bf16(&tqc, qkv, B * T, ldq, ldq, 64));
  LDMAE_TRY(make_tmap_bf16(&tdr, d_o, B * T, ldo, ldo, 128));
  LDMAE_TRY(make_tmap_bf16(&tdc, d_o, B * T, ldo, ldo, 64));
  constexpr int kSmem = AbGeo<HW>::kSmemBytes;
  static PerDeviceOnce attr;
  if (attr.pending()) {
    LDMAE_CUDA((cudaFuncSetAttribute(attn_bwd_kernel<false, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)));
    LDMAE_CUDA((cudaFuncSetAttribute(attn_bwd_kernel<true, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)));
    attr.mark();
  }
  float* nlse2 = delta + static_cast<size_t>(B) * H * T + 64;       // second half of the workspace
  attn_delta_kernel<<<cdiv(static_cast<size_t>(B) * T * H, 256), 256, 0, st>>>(delta, nlse2, lse2, static_cast<const __nv_bfloat16*>(d_o),
                                                                         static_cast<const __nv_bfloat16*>(o), B, T, H, hd, scale);
  LDMAE_LAUNCH_CHECK();
  AttnBwdParams p;
  p.trace = g_attn_trace;
  p.nlse2 = nlse2; p.delta = delta; p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.T = T; p.H = H; p.ld = ldq; p.q_col = q_col; p.k_col = k_col; p.v_col = v_col; p.hd = hd;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  // work items = (row block, head, sample); one CTA per SM walks them (persistent) unless LDMAE_ATTN_BWD_PERSIST=0
  const int n_rblk = static_cast<int>(cdiv(T, 128));
  const long long items = static_cast<long long>(n_rblk) * H * B;
  LDMAE_REQUIRE(items < (1ll << 31), "attention backward: too many work items");
  static int persist = -1;
  if (persist < 0) { const char* e = getenv("LDMAE_ATTN_BWD_PERSIST"); persist = e ? atoi(e) : 1; }
  const unsigned ctas = static_cast<unsigned>(persist ? std::min<long long>(items, device_sm_count()) : items);
  attn_bwd_kernel<true, HW><<<ctas, kAbThreads, kSmem, st>>>(tqr, tqc, tdr, tdc, p, n_rblk, static_cast<int>(items));
  LDMAE_LAUNCH_CHECK();
  attn_bwd_kernel<false, HW><<<ctas, kAbThreads, kSmem, st>>>(tqr, tqc, tdr, tdc, p, n_rblk, static_cast<int>(items));
  LDMAE_LAUNCH_CHECK();
  return LDMAE_OK;
}
static int run_attention_bwd(const void* qkv, int ldq, const void* o, const void* d_o, int ldo, const float* lse2, float* delta,
                             void* dqkv, int B, int T, int H, int q_col, int k_col, int v_col, float scale, cudaStream_t st,
                             int hd = 64, int HW = 64) {
  if (HW == 64) return run_attention_bwd_t<64>(qkv, ldq, o, d_o, ldo, lse2, delta, dqkv, B, T, H, hd, q_col, k_col, v_col, scale, st);
  return run_attention_bwd_t<128>(qkv, ldq, o, d_o, ldo, lse2, delta, dqkv, B, T, H, hd, q_col, k_col, v_col, scale, st);
}

// C[N1,N2] (fp32, leading dimension ldc) += alpha * sum_m p[m,N1] * q[m,N2]   (weight gradients; see gemm_wgrad_sm100.cuh)
static int gemm_wgrad(const void* pm, int ldp, const void* qm, int ldq, float* c, int ldc, int N1, int N2, int M, float alpha,
                      cudaStream_t st) {
  LDMAE_REQUIRE(N1 > 0 && N2 > 0 && M > 0, "wgrad: empty shape %d %d %d", N1, N2, M);
  LDMAE_REQUIRE(ldp % 8 == 0 && ldq % 8 == 0 && ldc % 4 == 0, "wgrad: leading dimensions must give 16-byte row pitches");
  constexpr int BN = 256, CG = 2;
  using Cfg = WgradCfg<BN, CG>;
  CUtensorMap tp, tq;
  LDMAE_TRY(make_tmap_2d(&tp, pm, 2, M, N1, ldp, 64, 64, 128));
  LDMAE_TRY(make_tmap_2d(&tq, qm, 2, M, N2, ldq, 64, 64, 128));
  EpiAccum::Params ep;
  LDMAE_TRY(make_tmap_out_f32(&ep.cmap, c, N1, N2, ldc));
  ep.alpha = alpha;
  auto kern = gemm_wgrad_kernel<BN, CG>;
  static PerDeviceOnce attr;
  if (attr.pending()) {
    LDMAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr.mark();
  }
  const int tiles = ((N1 + kBM * CG - 1) / (kBM * CG)) * ((N2 + BN - 1) / BN);
  const int num_kb = (M + kBK - 1) / kBK;
  const int clusters = device_sm_count() / CG;
  // split the contraction so that the units fill whole waves of clusters: among the split counts that keep at least 8
  // k-blocks per unit and give about two to four waves, take the one with the best last-wave occupancy
  const int max_splits = std::max(1, num_kb / 8);
  int splits = 1;
  double best = -1.0;
  for (int s = 1; s <= max_splits && static_cast<long long>(tiles) * s <= 4ll * clusters + tiles; ++s) {
    const long long units = static_cast<long long>(tiles) * s;
    const long long waves = (units + clusters - 1) / clusters;
    const double eff = static_cast<double>(units) / static_cast<double>(waves * clusters);
    const double score = eff - (units < 2ll * clusters ? 0.15 : 0.0) - 1e-4 * s;     // prefer >= 2 waves (tail hiding), then fewer splits
    if (score > best) { best = score; splits = s; }
  }
  int kb_per = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per - 1) / kb_per;            // no empty unit
  WgradShape g{N1, N2, M, splits, kb_per};
  const long long units = static_cast<long long>(tiles) * splits;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(std::min<long long>(units, clusters) * CG));
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  LDMAE_CUDA(cudaLaunchKernelEx(&cfg, kern, tp, tq, g, ep));
  ++g_launch_count;
  return LDMAE_OK;
}

// ------------------------------------------------------------------------------------------- LightningDiT
struct DitBlockW {
  DevBuf<__nv_bfloat16> w_qkv, w_proj, w12, w3;
  DevBuf<float> b_qkv, b_proj, b12, b3, qw, kw;
  DevBuf<float> qb, kb;        // q_norm / k_norm bias: nn.LayerNorm head norm of the use_rmsnorm=False variant
  // use_swiglu=False variant (timm Mlp, tanh GELU): fc1 lives in w12 / b12 (plain row order), fc2 in w3 / b3
  float attn_m0_log2 = -1.f;   // > 0: bound of |q.k| * scale * log2(e) from the q/k norm weights (constant-offset softmax)
};

}  // namespace ldmae

using namespace ldmae;

struct DitTrain;
static void dit_train_free(DitTrain* t);

struct ldmae_dit {
  ldmae_dit_config c;
  int D, T, G, Kp, H, Hp, nmod, Ntot, S /*norm slots*/, Nf, maxB, SS /*ssq partial slots*/;
  int K2 /*(hi | lo) patch row pitch of the tensor-core patch embedding: 2*Kp rounded up to 64*/;
  int hd /*real head_dim*/, HW /*head stride in the qkv buffer: 64, or 128 for wider heads*/, QW /*heads * HW*/;
  // weights
  DevBuf<float> pos, patch_w, patch_b, t_w0, t_b0, t_w2, t_b2, emb, rope_cos, rope_sin, rope_tab, norm_w, b_ada, b_f, w_f32;
  bool rope_wide_ok = false;   // rope_tab holds the compact table of the wide-head QKV epilogue (checked in finalize)
  DevBuf<__nv_bfloat16> w_ada, w_f;
  std::vector<DitBlockW> blk;
  DevBuf<int> slot_shift_off, slot_scale_off;
  std::vector<int> so_host, sco_host;   // host copies of the slot offsets (backward orchestration)
  std::vector<std::string> loaded;
  bool finalized = false;
  int debug_stop = -1;   // >= 0: dit_forward_impl returns after that many launch groups (ldmae_dit_debug_stop)
  // workspace (sized for maxB)
  DevBuf<float> xres, ssq, cvec_c, th1, mods, gmul, cvec_qkv, cvec_12, cvec_f, vbuf, k1buf, xtmp, cvec_all;
  // inference: W_fused = W_linear . W_adaLN[shift slot] for every modulated Linear, so ONE GEMM on silu(c) yields all per-sample
  // vectors shift_b . W^T + bias (dit_build_fused_shift); the training forward keeps the two-step route the backward differentiates
  DevBuf<__nv_bfloat16> w_shf, shf_tmp, shf_btmp;
  DevBuf<__nv_bfloat16> patch_w2, tok2;     // tensor-core patch embedding: [D, 2Kp] doubled weight, [maxB*T, 2Kp] (hi | lo) patches
  DevBuf<float> b_shf;
  int Nsh = 0;
  bool shf_valid = false;
  bool fused = true;       // shipped recipe (RMSNorm + SwiGLU): norms folded into the GEMM epilogues; else the generic path
  DevBuf<__nv_bfloat16> abuf, qkv, obuf, hbuf, sc, shift_bf16;
  DitTrain* tr = nullptr;   // activations kept by the training forward, backward workspace, gradients (dit_train.cuh)
  bool wT_valid = false;    // transposed bf16 weight copies (data-gradient GEMM operands) are current
  long long ws_gen = 0;     // bumped by every forward and workspace re-allocation: a backward checks its forward is the latest
  DevBuf<int> label_err;    // device flag: a class label outside y_embedder.embedding_table was seen (clamped, never read OOB)
  int* label_err_host = nullptr;   // pinned mirror, refreshed after every conditioning kernel (read without a sync)
  ~ldmae_dit() {
    dit_train_free(tr);
    if (label_err_host) cudaFreeHost(label_err_host);
  }
};

// Reports (once) a label that was out of range in an EARLIER call on this handle: the flag travels through a pinned
// mirror that the stream refreshes after each conditioning kernel, so no call waits for the device here.
static int dit_label_check(ldmae_dit* h) {
  if (h->label_err_host && *h->label_err_host != 0) {
    *h->label_err_host = 0;
    cudaMemset(h->label_err.p, 0, sizeof(int));
    return set_error(LDMAE_ERR_INVALID, "class label outside y_embedder.embedding_table (%d rows) in an earlier call on this handle "
                     "(labels must lie in [0, %d]; the null class needs class_dropout_prob > 0)", h->c.num_embeddings, h->c.num_embeddings - 1);
  }
  return LDMAE_OK;
}

static int dit_alloc_ws(ldmae_dit* h, int B) {
  const size_t M = static_cast<size_t>(B) * h->T;
  const int D = h->D;
  LDMAE_TRY(h->xres.alloc(M * D));
  LDMAE_TRY(h->abuf.alloc(M * D));
  LDMAE_TRY(h->qkv.alloc(M * 3 * h->QW));
  LDMAE_TRY(h->obuf.alloc(M * D));
  LDMAE_TRY(h->hbuf.alloc(M * h->Hp));
  LDMAE_TRY(h->ssq.alloc(M * h->SS));
  LDMAE_TRY(h->cvec_c.alloc(static_cast<size_t>(B) * D));
  LDMAE_TRY(h->th1.alloc(static_cast<size_t>(B) * D));
  LDMAE_TRY(h->sc.alloc(static_cast<size_t>(B) * D));
  LDMAE_TRY(h->mods.alloc(static_cast<size_t>(B) * h->Ntot));
  LDMAE_TRY(h->shift_bf16.alloc(static_cast<size_t>(h->S) * B * D));
  LDMAE_TRY(h->gmul.alloc(static_cast<size_t>(h->S) * B * D));
  LDMAE_TRY(h->cvec_qkv.alloc(static_cast<size_t>(h->c.depth) * B * 3 * h->QW));
  LDMAE_TRY(h->cvec_12.alloc(static_cast<size_t>(h->c.depth) * B * 2 * h->Hp));
  LDMAE_TRY(h->cvec_f.alloc(static_cast<size_t>(B) * h->Nf));
  LDMAE_TRY(h->cvec_all.alloc(static_cast<size_t>(B) * h->Nsh));
  LDMAE_TRY(h->tok2.alloc(M * h->K2, true));
  const size_t lat = static_cast<size_t>(B) * h->c.in_channels * h->c.input_size * h->c.input_size;
  LDMAE_TRY(h->vbuf.alloc(lat));
  LDMAE_TRY(h->k1buf.alloc(lat));
  LDMAE_TRY(h->xtmp.alloc(lat));
  h->maxB = B;
  ++h->ws_gen;
  return LDMAE_OK;
}

extern "C" const char* ldmae_last_error(void) { return last_error().c_str(); }
extern "C" int ldmae_version(void) { return 100; }
extern "C" long long ldmae_launch_count(void) { return g_launch_count; }

extern "C" int ldmae_profile_begin(void) {
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = true;
  return LDMAE_OK;
}
// Stops profiling; ms[c] = summed device time of class c, launches[c] = number of timed scopes.
extern "C" int ldmae_profile_end(double* ms, long long* launches, int32_t nclasses) {
  g_prof_on = false;
  LDMAE_CUDA(cudaDeviceSynchronize());
  for (int c = 0; c < nclasses; ++c) { ms[c] = 0.0; launches[c] = 0; }
  for (auto& r : g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.cls < nclasses) { ms[r.cls] += t; launches[r.cls] += 1; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return LDMAE_OK;
}

extern "C" int ldmae_device_info(int* sm_count, int* cc) {
  int dev = 0, major = 0, minor = 0, sms = 0;
  LDMAE_CUDA(cudaGetDevice(&dev));
  LDMAE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  LDMAE_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  LDMAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count) *sm_count = sms;
  if (cc) *cc = major * 10 + minor;
  return LDMAE_OK;
}

static int require_sm100() {
  int cc = 0;
  LDMAE_TRY(ldmae_device_info(nullptr, &cc));
  if (cc / 10 != 10) return set_error(LDMAE_ERR_INVALID, "ldmae_b200 needs an sm_100 (B200) device, found sm_%d", cc);
  return LDMAE_OK;
}

extern "C" int ldmae_dit_create(const ldmae_dit_config* cfg, ldmae_dit** out) {
  LDMAE_REQUIRE(cfg && out, "null argument");
  LDMAE_TRY(require_sm100());
  const ldmae_dit_config& c = *cfg;
  if (!(c.use_rmsnorm == 1 && c.use_swiglu == 1))
    LDMAE_REQUIRE(c.hidden_size / std::max(1, c.num_heads) == 64,
                  "the LayerNorm / GELU-Mlp variants (use_rmsnorm=False / use_swiglu=False) are built for head_dim 64");
  LDMAE_REQUIRE(c.num_heads > 0 && c.hidden_size % c.num_heads == 0, "hidden_size %% num_heads != 0");
  {
    const int hd = c.hidden_size / c.num_heads;
    LDMAE_REQUIRE(hd == 64 || (hd > 64 && hd <= 128 && hd % 8 == 0),
                  "head_dim %d: built for 64 (tuned path) and for multiples of 8 in (64, 128] (XL: 72, inference only)", hd);
  }
  LDMAE_REQUIRE(c.hidden_size % 64 == 0, "hidden_size %d must be a multiple of 64", c.hidden_size);
  LDMAE_REQUIRE(c.input_size % c.patch_size == 0, "input_size %% patch_size != 0");
  LDMAE_REQUIRE((c.in_channels * c.patch_size * c.patch_size) % 4 == 0, "C*p*p must be a multiple of 4");
  ldmae_dit* h = new ldmae_dit();
  h->c = c;
  h->D = c.hidden_size;
  h->G = c.input_size / c.patch_size;
  h->T = h->G * h->G;
  h->Kp = c.in_channels * c.patch_size * c.patch_size;
  h->K2 = (2 * h->Kp + 63) / 64 * 64;
  h->H = c.mlp_hidden;
  h->Hp = (c.mlp_hidden + 31) / 32 * 32;
  h->nmod = c.wo_shift ? 4 : 6;
  h->Ntot = (c.depth * h->nmod + 2) * h->D;
  h->S = 2 * c.depth + 1;
  h->Nf = c.patch_size * c.patch_size * c.in_channels * (c.learn_sigma ? 2 : 1);
  h->SS = (c.hidden_size + 127) / 128;
  h->hd = c.hidden_size / c.num_heads;
  h->HW = h->hd == 64 ? 64 : 128;
  h->QW = c.num_heads * h->HW;
  h->fused = c.use_rmsnorm == 1 && c.use_swiglu == 1;
  h->Nsh = h->fused ? c.depth * (3 * h->QW + 2 * ((c.mlp_hidden + 31) / 32 * 32)) + c.patch_size * c.patch_size * c.in_channels * (c.learn_sigma ? 2 : 1) : 4;
  const int D = h->D;
  h->blk.resize(c.depth);
  int r = LDMAE_OK;
  auto A = [&](int rc) { if (r == LDMAE_OK) r = rc; };
  A(h->pos.alloc(static_cast<size_t>(h->T) * D));
  A(h->patch_w.alloc(static_cast<size_t>(D) * h->Kp));
  A(h->patch_w2.alloc(static_cast<size_t>(D) * h->K2, true));
  A(h->patch_b.alloc(D));
  A(h->t_w0.alloc(static_cast<size_t>(D) * 256)); A(h->t_b0.alloc(D));
  A(h->t_w2.alloc(static_cast<size_t>(D) * D)); A(h->t_b2.alloc(D));
  A(h->emb.alloc(static_cast<size_t>(c.num_embeddings) * D));
  A(h->rope_cos.alloc(static_cast<size_t>(h->T) * h->hd)); A(h->rope_sin.alloc(static_cast<size_t>(h->T) * h->hd));
  A(h->rope_tab.alloc(static_cast<size_t>(2) * h->G * (h->hd > 64 ? h->hd / 2 : 32) + 1));
  A(h->norm_w.alloc(static_cast<size_t>(h->S) * D));
  A(h->w_ada.alloc(static_cast<size_t>(h->Ntot) * D)); A(h->b_ada.alloc(h->Ntot));
  A(h->w_f.alloc(static_cast<size_t>(h->Nf) * D)); A(h->b_f.alloc(h->Nf)); A(h->w_f32.alloc(static_cast<size_t>(h->Nf) * D));
  A(h->w_shf.alloc(static_cast<size_t>(h->Nsh) * D)); A(h->b_shf.alloc(h->Nsh)); A(h->shf_tmp.alloc(static_cast<size_t>(D) * D)); A(h->shf_btmp.alloc(D));
  for (auto& b : h->blk) {
    A(b.w_qkv.alloc(static_cast<size_t>(3 * h->QW) * D)); A(b.b_qkv.alloc(3 * h->QW, true));
    A(b.w_proj.alloc(static_cast<size_t>(D) * D)); A(b.b_proj.alloc(D));
    const int w12_rows = c.use_swiglu ? 2 * h->Hp : h->Hp;
    A(b.w12.alloc(static_cast<size_t>(w12_rows) * D)); A(b.b12.alloc(w12_rows));
    A(b.w3.alloc(static_cast<size_t>(D) * h->Hp)); A(b.b3.alloc(D));
    A(b.qw.alloc(h->HW, true)); A(b.kw.alloc(h->HW, true));
    A(b.qb.alloc(h->HW, true)); A(b.kb.alloc(h->HW, true));
  }
  // adaLN slot -> column offsets inside a mods row
  std::vector<int> so(h->S), sco(h->S);
  for (int i = 0; i < c.depth; ++i) {
    const int base = i * h->nmod * D;
    if (c.wo_shift) {
      so[2 * i] = -1; sco[2 * i] = base;                // scale_msa, gate_msa, scale_mlp, gate_mlp
      so[2 * i + 1] = -1; sco[2 * i + 1] = base + 2 * D;
    } else {
      so[2 * i] = base; sco[2 * i] = base + D;          // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
      so[2 * i + 1] = base + 3 * D; sco[2 * i + 1] = base + 4 * D;
    }
  }
  so[2 * c.depth] = c.depth * h->nmod * D;
  sco[2 * c.depth] = c.depth * h->nmod * D + D;
  h->so_host = so; h->sco_host = sco;
  A(h->slot_shift_off.alloc(h->S)); A(h->slot_scale_off.alloc(h->S));
  if (r == LDMAE_OK && cudaMemcpy(h->slot_shift_off.p, so.data(), h->S * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
    r = set_error(LDMAE_ERR_CUDA, "memcpy slot offsets");
  if (r == LDMAE_OK && cudaMemcpy(h->slot_scale_off.p, sco.data(), h->S * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
    r = set_error(LDMAE_ERR_CUDA, "memcpy slot offsets");
  if (r == LDMAE_OK) r = dit_alloc_ws(h, std::max(1, c.max_batch));
  A(h->label_err.alloc(1, true));
  if (r == LDMAE_OK && cudaHostAlloc(reinterpret_cast<void**>(&h->label_err_host), sizeof(int), cudaHostAllocDefault) != cudaSuccess)
    r = set_error(LDMAE_ERR_CUDA, "cudaHostAlloc failed");
  if (r == LDMAE_OK) *h->label_err_host = 0;
  if (r != LDMAE_OK) { delete h; return r; }
  *out = h;
  return LDMAE_OK;
}

extern "C" void ldmae_dit_destroy(ldmae_dit* h) { delete h; }

static int copy_f32(float* dst, const float* src, int64_t n, int64_t expect, const char* name, cudaStream_t st) {
  LDMAE_REQUIRE(n == expect, "%s: expected %lld elements, got %lld", name, (long long)expect, (long long)n);
  LDMAE_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return LDMAE_OK;
}
static int pack_bf16(__nv_bfloat16* dst, const float* src, int rows, int Ks, int Kd, int64_t n, const char* name,
                     cudaStream_t st, int mode = 0, int H = 0, int src_rows = -1) {
  const int64_t expect = static_cast<int64_t>(src_rows < 0 ? rows : src_rows) * Ks;
  LDMAE_REQUIRE(n == expect, "%s: expected %lld elements, got %lld", name, (long long)expect, (long long)n);
  pack_rows_bf16_kernel<<<rows, 256, 0, st>>>(dst, src, rows, Ks, Kd, mode, H);
  LDMAE_LAUNCH_CHECK();
  return LDMAE_OK;
}

extern "C" int ldmae_dit_load_tensor(ldmae_dit* h, const char* name, const float* data, int64_t numel, void* stream) {
  LDMAE_REQUIRE(h && name && data, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = h->D;
  const std::string k(name);
  int rc = LDMAE_OK;
  int bi = -1;
  std::string sub;
  if (k.rfind("blocks.", 0) == 0) {
    const size_t dot = k.find('.', 7);
    LDMAE_REQUIRE(dot != std::string::npos, "bad key %s", name);
    bi = atoi(k.substr(7, dot - 7).c_str());
    sub = k.substr(dot + 1);
    LDMAE_REQUIRE(bi >= 0 && bi < h->c.depth, "block index out of range in %s", name);
  }
  if (k == "pos_embed") rc = copy_f32(h->pos.p, data, numel, (int64_t)h->T * D, name, st);
  else if (k == "x_embedder.proj.weight") {
    rc = copy_f32(h->patch_w.p, data, numel, (int64_t)D * h->Kp, name, st);
    if (rc == LDMAE_OK) {
      pack_dup_bf16_kernel<<<cdiv(static_cast<size_t>(D) * h->Kp, 256), 256, 0, st>>>(h->patch_w2.p, data, D, h->Kp, h->K2);
      LDMAE_LAUNCH_CHECK();
    }
  }
  else if (k == "x_embedder.proj.bias") rc = copy_f32(h->patch_b.p, data, numel, D, name, st);
  else if (k == "t_embedder.mlp.0.weight") rc = copy_f32(h->t_w0.p, data, numel, (int64_t)D * 256, name, st);
  else if (k == "t_embedder.mlp.0.bias") rc = copy_f32(h->t_b0.p, data, numel, D, name, st);
  else if (k == "t_embedder.mlp.2.weight") rc = copy_f32(h->t_w2.p, data, numel, (int64_t)D * D, name, st);
  else if (k == "t_embedder.mlp.2.bias") rc = copy_f32(h->t_b2.p, data, numel, D, name, st);
  else if (k == "y_embedder.embedding_table.weight") rc = copy_f32(h->emb.p, data, numel, (int64_t)h->c.num_embeddings * D, name, st);
  else if (k == "feat_rope.freqs_cos") rc = copy_f32(h->rope_cos.p, data, numel, (int64_t)h->T * h->hd, name, st);
  else if (k == "feat_rope.freqs_sin") rc = copy_f32(h->rope_sin.p, data, numel, (int64_t)h->T * h->hd, name, st);
  else if (k == "final_layer.norm_final.weight") rc = copy_f32(h->norm_w.p + (size_t)(2 * h->c.depth) * D, data, numel, D, name, st);
  else if (k == "final_layer.linear.weight") {
    rc = pack_bf16(h->w_f.p, data, h->Nf, D, D, numel, name, st);
    if (rc == LDMAE_OK) rc = copy_f32(h->w_f32.p, data, numel, (int64_t)h->Nf * D, name, st);
  }
  else if (k == "final_layer.linear.bias") rc = copy_f32(h->b_f.p, data, numel, h->Nf, name, st);
  else if (k == "final_layer.adaLN_modulation.1.weight")
    rc = pack_bf16(h->w_ada.p + (size_t)h->c.depth * h->nmod * D * D, data, 2 * D, D, D, numel, name, st);
  else if (k == "final_layer.adaLN_modulation.1.bias")
    rc = copy_f32(h->b_ada.p + (size_t)h->c.depth * h->nmod * D, data, numel, 2 * D, name, st);
  else if (bi >= 0) {
    DitBlockW& b = h->blk[bi];
    if (sub == "norm1.weight") rc = copy_f32(h->norm_w.p + (size_t)(2 * bi) * D, data, numel, D, name, st);
    else if (sub == "norm2.weight") rc = copy_f32(h->norm_w.p + (size_t)(2 * bi + 1) * D, data, numel, D, name, st);
    else if (sub == "attn.qkv.weight" && h->HW == 64) rc = pack_bf16(b.w_qkv.p, data, 3 * D, D, D, numel, name, st);
    else if (sub == "attn.qkv.bias" && h->HW == 64) rc = copy_f32(b.b_qkv.p, data, numel, 3 * D, name, st);
    else if (sub == "attn.qkv.weight") {
      // wider heads: every head gets a 128-row slot, rows beyond head_dim stay zero
      LDMAE_REQUIRE(numel == (int64_t)3 * D * D, "%s: bad size", name);
      pad_heads_rows_kernel<<<3 * h->QW, 128, 0, st>>>(b.w_qkv.p, nullptr, data, nullptr, h->c.num_heads, h->hd, D, h->HW);
      LDMAE_LAUNCH_CHECK();
    } else if (sub == "attn.qkv.bias") {
      LDMAE_REQUIRE(numel == 3 * D, "%s: bad size", name);
      LDMAE_CUDA(cudaMemsetAsync(b.b_qkv.p, 0, 3 * h->QW * sizeof(float), st));
      LDMAE_CUDA(cudaMemcpy2DAsync(b.b_qkv.p, h->HW * sizeof(float), data, h->hd * sizeof(float), h->hd * sizeof(float),
                                   3 * h->c.num_heads, cudaMemcpyDeviceToDevice, st));
    }
    else if (sub == "attn.q_norm.weight") rc = copy_f32(b.qw.p, data, numel, h->hd, name, st);
    else if (sub == "attn.k_norm.weight") rc = copy_f32(b.kw.p, data, numel, h->hd, name, st);
    else if (sub == "attn.q_norm.bias" && !h->c.use_rmsnorm) rc = copy_f32(b.qb.p, data, numel, h->hd, name, st);
    else if (sub == "attn.k_norm.bias" && !h->c.use_rmsnorm) rc = copy_f32(b.kb.p, data, numel, h->hd, name, st);
    else if (sub == "mlp.fc1.weight" && !h->c.use_swiglu) rc = pack_bf16(b.w12.p, data, h->Hp, D, D, numel, name, st, 0, 0, h->H);
    else if (sub == "mlp.fc1.bias" && !h->c.use_swiglu) {
      LDMAE_CUDA(cudaMemsetAsync(b.b12.p, 0, h->Hp * sizeof(float), st));
      rc = copy_f32(b.b12.p, data, numel, h->H, name, st);
    }
    else if (sub == "mlp.fc2.weight" && !h->c.use_swiglu) rc = pack_bf16(b.w3.p, data, D, h->H, h->Hp, numel, name, st);
    else if (sub == "mlp.fc2.bias" && !h->c.use_swiglu) rc = copy_f32(b.b3.p, data, numel, D, name, st);
    else if (sub == "attn.proj.weight") rc = pack_bf16(b.w_proj.p, data, D, D, D, numel, name, st);
    else if (sub == "attn.proj.bias") rc = copy_f32(b.b_proj.p, data, numel, D, name, st);
    else if (sub == "mlp.w12.weight") rc = pack_bf16(b.w12.p, data, 2 * h->Hp, D, D, numel, name, st, 1, h->H, 2 * h->H);
    else if (sub == "mlp.w12.bias") {
      LDMAE_REQUIRE(numel == 2 * h->H, "%s: expected %d elements", name, 2 * h->H);
      pack_vec_f32_kernel<<<cdiv(2 * h->Hp, 256), 256, 0, st>>>(b.b12.p, data, 2 * h->Hp, 1, h->H);
      LDMAE_LAUNCH_CHECK();
    } else if (sub == "mlp.w3.weight") rc = pack_bf16(b.w3.p, data, D, h->H, h->Hp, numel, name, st);
    else if (sub == "mlp.w3.bias") rc = copy_f32(b.b3.p, data, numel, D, name, st);
    else if (sub == "adaLN_modulation.1.weight")
      rc = pack_bf16(h->w_ada.p + (size_t)bi * h->nmod * D * D, data, h->nmod * D, D, D, numel, name, st);
    else if (sub == "adaLN_modulation.1.bias") rc = copy_f32(h->b_ada.p + (size_t)bi * h->nmod * D, data, numel, h->nmod * D, name, st);
    else return set_error(LDMAE_ERR_INVALID, "unknown LightningDiT state_dict key %s", name);
  } else {
    return set_error(LDMAE_ERR_INVALID, "unknown LightningDiT state_dict key %s", name);
  }
  if (rc == LDMAE_OK && std::find(h->loaded.begin(), h->loaded.end(), k) == h->loaded.end()) h->loaded.push_back(k);
  h->wT_valid = false;
  h->shf_valid = false;
  return rc;
}

// n tensors in one call (weight re-upload after every fused optimizer step)
extern "C" int ldmae_dit_load_tensors(ldmae_dit* h, const char* const* names, const float* const* data, const int64_t* numels,
                                      int32_t n, void* stream) {
  LDMAE_REQUIRE(h && names && data && numels && n >= 0, "bad argument");
  for (int i = 0; i < n; ++i) LDMAE_TRY(ldmae_dit_load_tensor(h, names[i], data[i], numels[i], stream));
  return LDMAE_OK;
}

extern "C" int ldmae_dit_finalize(ldmae_dit* h, void* stream) {
  LDMAE_REQUIRE(h, "null handle");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int expect = 8 + 5 + 12 * h->c.depth;   // without qk-norm / rope keys
  if (!h->c.use_rmsnorm) expect -= 1 + 2 * h->c.depth;      // LayerNorm(elementwise_affine=False): no norm weights
  if (h->c.use_qknorm) expect += (h->c.use_rmsnorm ? 2 : 4) * h->c.depth;   // nn.LayerNorm head norms carry a bias too
  if (h->c.use_rope) expect += 2;
  if ((int)h->loaded.size() != expect)
    return set_error(LDMAE_ERR_STATE, "LightningDiT weights incomplete: %d of %d tensors loaded", (int)h->loaded.size(), expect);
  if (h->c.use_rope && h->HW == 64) {
    // compact axial table for the QKV epilogue + a check that the loaded [T, 64] buffers really have that structure
    float* maxdiff = h->rope_tab.p + static_cast<size_t>(2) * h->G * 32;
    float md = 0.f;
    LDMAE_CUDA(cudaMemsetAsync(maxdiff, 0, sizeof(float), st));
    rope_compact_kernel<<<cdiv(2 * h->G * 32, 128), 128, 0, st>>>(h->rope_tab.p, h->rope_cos.p, h->rope_sin.p, h->G);
    LDMAE_LAUNCH_CHECK();
    rope_check_kernel<<<cdiv(static_cast<size_t>(h->T) * 64, 256), 256, 0, st>>>(maxdiff, h->rope_tab.p, h->rope_cos.p, h->rope_sin.p, h->G);
    LDMAE_LAUNCH_CHECK();
    LDMAE_CUDA(cudaMemcpyAsync(&md, maxdiff, sizeof(float), cudaMemcpyDeviceToHost, st));
    LDMAE_CUDA(cudaStreamSynchronize(st));
    // Tolerance: the buffers are computed by the HOST's cos / sin at model construction (models/pos_embed.py:96-133).  On most
    // hosts equal angles give bit-equal values and the deviation is exactly 0; some GPU boxes of the pool produce tables whose
    // repeated entries differ by up to 1.5e-4 (deterministically, round 2: bench.py aborted there with the former 1e-6 bound).
    // That is far below the bf16 resolution of q and k (4e-3); a table with another layout deviates by O(1).
    if (!(md <= 1e-3f))
      return set_error(LDMAE_ERR_INVALID, "feat_rope.freqs_cos/sin are not the 2-D axial table of models/pos_embed.py:96-133 "
                       "(max deviation %g): unsupported RoPE buffers", md);
  }
  h->rope_wide_ok = false;
  if (h->c.use_rope && h->HW == 128 && h->hd % 8 == 0) {
    // wide heads: the same compact table (interleaved cos / sin); buffers without the axial structure keep the element-wise
    // global reads of the epilogue instead of failing
    const int w = h->hd / 2;
    float* maxdiff = h->rope_tab.p + static_cast<size_t>(2) * h->G * w;
    float md = 0.f;
    LDMAE_CUDA(cudaMemsetAsync(maxdiff, 0, sizeof(float), st));
    rope_compact_wide_kernel<<<cdiv(2 * h->G * w, 128), 128, 0, st>>>(h->rope_tab.p, h->rope_cos.p, h->rope_sin.p, h->G, h->hd);
    LDMAE_LAUNCH_CHECK();
    rope_check_wide_kernel<<<cdiv(static_cast<size_t>(h->T) * h->hd, 256), 256, 0, st>>>(maxdiff, h->rope_tab.p, h->rope_cos.p,
                                                                                       h->rope_sin.p, h->G, h->hd);
    LDMAE_LAUNCH_CHECK();
    LDMAE_CUDA(cudaMemcpyAsync(&md, maxdiff, sizeof(float), cudaMemcpyDeviceToHost, st));
    LDMAE_CUDA(cudaStreamSynchronize(st));
    static int tab_on = -1;
    if (tab_on < 0) { const char* e = getenv("LDMAE_ROPE_WIDE_TABLE"); tab_on = e ? atoi(e) : 1; }
    h->rope_wide_ok = md <= 1e-3f && tab_on;
  }
  // Score bound of the qk-normed attention: |q| <= 8 max|q_norm.w|, |k| <= 8 max|k_norm.w| (RMSNorm over 64, RoPE is a
  // rotation), so |q.k| / 8 * log2(e) <= 8 * log2(e) * max|wq| * max|wk|.  Re-evaluated whenever weights are (re)loaded.
  for (auto& b : h->blk) b.attn_m0_log2 = -1.f;
  static int fixed = -1;
  if (fixed < 0) { const char* e = getenv("LDMAE_ATTN_FIXED_MAX"); fixed = e ? atoi(e) : 1; }
  if (h->c.use_qknorm && h->c.use_rmsnorm && fixed) {
    // general head_dim: |q| <= sqrt(hd) max|q_norm.w| (RMS over hd, RoPE is a rotation), scale = 1 / sqrt(hd)
    //   => |q.k| * scale * log2(e) <= sqrt(hd) * log2(e) * max|wq| * max|wk|     (8 * log2(e) ... for head_dim 64)
    const int depth = h->c.depth, hdr = h->hd;
    std::vector<float> w(static_cast<size_t>(depth) * 256);
    for (int i = 0; i < depth; ++i) {
      LDMAE_CUDA(cudaMemcpyAsync(w.data() + i * 256, h->blk[i].qw.p, hdr * sizeof(float), cudaMemcpyDeviceToHost, st));
      LDMAE_CUDA(cudaMemcpyAsync(w.data() + i * 256 + 128, h->blk[i].kw.p, hdr * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    LDMAE_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < depth; ++i) {
      float mq = 0.f, mk = 0.f;
      for (int j = 0; j < hdr; ++j) { mq = std::max(mq, std::fabs(w[i * 256 + j])); mk = std::max(mk, std::fabs(w[i * 256 + 128 + j])); }
      // 2 % head-room for the bf16 rounding of q and k; beyond 48 the exponent range [-2 m0, 0] gets uncomfortable
      const float m0 = std::sqrt(static_cast<float>(hdr)) * 1.4426950408889634f * mq * mk * 1.02f;
      if (m0 > 0.f && m0 <= 48.f) h->blk[i].attn_m0_log2 = m0;
    }
  }
  h->finalized = true;
  return LDMAE_OK;
}

// Views of the buffers one forward writes: the shared inference workspace (in place, one block at a time) or, for the
// training forward, the per-block slices of the activation store (dit_train.cuh).
struct DitTrain {
  int B = 0;
  // kept by the forward
  DevBuf<float> X, SSQ, LSE, th1pre, tfreq;
  DevBuf<__nv_bfloat16> A, QKV, QKR, O, MA, MM, H12, HB, XTOK;
  DevBuf<long long> ykeep;
  // backward workspace
  DevBuf<float> dx, delta, dg, sdx, dcv_qkv, dcv_12, dcv_f, dmods, dsc, dcond, dth1, dth1pre, grads, gscratch;
  DevBuf<__nv_bfloat16> dY, dO, G, dQKV, dH, dH12, dYf, cv_bf16, dmods_bf16;
  // transposed weights (operands of the data-gradient GEMMs)
  std::vector<DevBuf<__nv_bfloat16>> wT_qkv, wT_proj, wT_12, wT_3;
  DevBuf<__nv_bfloat16> wT_ada, wT_f;
  // gradient layout inside `grads` (internal packed layouts; ldmae_dit_grad_read converts to the reference's)
  std::map<std::string, std::pair<size_t, size_t>> goff;
  size_t gtotal = 0;
  bool have_forward = false;
  int lastB = 0;
  long long fwd_gen = -1;   // ldmae_dit::ws_gen when the kept forward ran
};
static void dit_train_free(DitTrain* t) { delete t; }

// Pre-multiplied shift path of the inference forward.  For a modulated Linear  y = Linear(modulate(RMSNorm(x), shift_b, scale_b))
// the per-sample vector is  cvec_b = shift_b . W^T + bias  with  shift_b = silu(c_b) . W_ada[slot]^T + b_ada[slot]
// (lightningdit.py:239-250), i.e.  cvec_b = silu(c_b) . (W . W_ada[slot])^T + (W . b_ada[slot] + bias):  W_fused [Nsh, D] (bf16)
// and b_fused [Nsh] are built once per weight upload with the library's own GEMM; rows follow the packed layouts of w_qkv / w12.
static int transpose_bf16(__nv_bfloat16* dst, const __nv_bfloat16* src, int R, int C, cudaStream_t st);
__global__ void f32_to_bf16_kernel(__nv_bfloat16* out, const float* in, size_t n);
static int dit_build_fused_shift(ldmae_dit* h, cudaStream_t st) {
  if (h->shf_valid) return LDMAE_OK;
  const int D = h->D, per = 3 * h->QW + 2 * h->Hp;
  for (int i = 0; i < h->c.depth; ++i) {
    DitBlockW& b = h->blk[i];
    for (int which = 0; which < 2; ++which) {
      const int slot = 2 * i + which;
      const int rows = which == 0 ? 3 * h->QW : 2 * h->Hp;
      const __nv_bfloat16* W = which == 0 ? b.w_qkv.p : b.w12.p;
      const float* bias = which == 0 ? b.b_qkv.p : b.b12.p;
      const size_t r0 = static_cast<size_t>(i) * per + (which == 0 ? 0 : 3 * h->QW);
      const int so = h->so_host[slot];
      if (so < 0) {      // wo_shift: no shift term, the vector is the bias
        LDMAE_CUDA(cudaMemsetAsync(h->w_shf.p + r0 * D, 0, static_cast<size_t>(rows) * D * sizeof(__nv_bfloat16), st));
        LDMAE_CUDA(cudaMemcpyAsync(h->b_shf.p + r0, bias, rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
        continue;
      }
      // F[n, k'] = sum_j W[n, j] * W_ada[so + j, k']  ==  TN GEMM of W against the transposed D x D slot
      LDMAE_TRY(transpose_bf16(h->shf_tmp.p, h->w_ada.p + static_cast<size_t>(so) * D, D, D, st));
      LDMAE_TRY((gemm_store<__nv_bfloat16, 0>(W, D, h->shf_tmp.p, D, GemmShape{rows, D, D}, h->w_shf.p + r0 * D, D, nullptr, st)));
      // b_fused[n] = sum_j W[n, j] * bf16(b_ada[so + j]) + bias[n]   (the two-step route rounds the whole shift to bf16 as well)
      f32_to_bf16_kernel<<<cdiv(D, 256), 256, 0, st>>>(h->shf_btmp.p, h->b_ada.p + so, static_cast<size_t>(D));
      LDMAE_LAUNCH_CHECK();
      LDMAE_TRY((gemm_store<float, 0>(h->shf_btmp.p, D, W, D, GemmShape{1, rows, D}, h->b_shf.p + r0, rows, bias, st)));
    }
  }
  {
    // final layer (lightningdit.py:267-272): rows of final_layer.linear against the final adaLN's shift slot
    const size_t r0 = static_cast<size_t>(h->c.depth) * per;
    const int so = h->so_host[2 * h->c.depth];
    LDMAE_TRY(transpose_bf16(h->shf_tmp.p, h->w_ada.p + static_cast<size_t>(so) * D, D, D, st));
    LDMAE_TRY((gemm_store<__nv_bfloat16, 0>(h->w_f.p, D, h->shf_tmp.p, D, GemmShape{h->Nf, D, D}, h->w_shf.p + r0 * D, D, nullptr, st)));
    f32_to_bf16_kernel<<<cdiv(D, 256), 256, 0, st>>>(h->shf_btmp.p, h->b_ada.p + so, static_cast<size_t>(D));
    LDMAE_LAUNCH_CHECK();
    LDMAE_TRY((gemm_store<float, 0>(h->shf_btmp.p, D, h->w_f.p, D, GemmShape{1, h->Nf, D}, h->b_shf.p + r0, h->Nf, h->b_f.p, st)));
  }
  h->shf_valid = true;
  return LDMAE_OK;
}

// Patch embedding launch: the weight goes through shared memory (transposed) whenever it fits next to two resident blocks.
static int launch_patch_embed(ldmae_dit* h, float* xs, __nv_bfloat16* as, float* ss, const float* x, const float* gmul, int B,
                              int src_mod, cudaStream_t st, bool tensor_ok = true) {
  const ldmae_dit_config& c = h->c;
  static int tc = -1;
  if (tc < 0) { const char* e = getenv("LDMAE_PATCH_TC"); tc = e ? atoi(e) : 1; }
  if (tc && tensor_ok && h->T % 128 == 0 && h->Kp % 4 == 0 && h->D % 64 == 0) {
    // tensor-core route: (hi | lo) bf16 patches . [W | W]^T through the residual epilogue, whose "x" tile is the pos_embed row
    // of the token (x_row_mod = T): HBM-bound at 6 bytes per output element instead of CUDA-core bound
    const int M = B * h->T, K2 = h->K2, D = h->D;
    patchify_hilo_kernel<<<cdiv(static_cast<size_t>(M) * h->Kp, 256), 256, 0, st>>>(h->tok2.p, x, B, c.in_channels, c.input_size,
                                                                                   c.patch_size, src_mod, K2);
    LDMAE_LAUNCH_CHECK();
    EpiResidualDeep::Params ep;
    LDMAE_TRY(make_tmap_out_f32(&ep.xmap, h->pos.p, h->T, D, D));
    LDMAE_TRY(make_tmap_out_f32(&ep.xmap_out, xs, M, D, D));
    ep.mmap = ep.xmap;
    ep.has_anext = as != nullptr;
    if (as) LDMAE_TRY(make_tmap_out_bf16(&ep.amap, as, M, D, D));
    else ep.amap = ep.xmap;
    ep.bias = h->patch_b.p; ep.gate = nullptr; ep.gnext = gmul; ep.ssq = ss;
    ep.gate_ld = 0; ep.gnext_ld = D; ep.rows_per_sample = h->T; ep.ss_slots = h->SS; ep.x_row_mod = h->T;
    ep.trace = nullptr;
    return gemm_auto<EpiResidualDeep>(h->tok2.p, K2, h->patch_w2.p, K2, GemmShape{M, D, K2}, ep, st);
  }
  const bool w_smem = patch_embed_smem(h->Kp, h->D, true) <= 100 * 1024;
  const size_t sm = patch_embed_smem(h->Kp, h->D, w_smem);
  static PerDeviceOnce attr;
  if (attr.pending()) {
    LDMAE_CUDA(cudaFuncSetAttribute(patch_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr.mark();
  }
  LDMAE_REQUIRE(sm <= 200 * 1024, "patch embed: patch vector of %d values does not fit shared memory", h->Kp);
  dim3 grid(cdiv(h->T, kPeTokens), B);
  patch_embed_kernel<<<grid, 256, sm, st>>>(xs, as, ss, x, h->patch_w.p, h->patch_b.p, h->pos.p, gmul, c.in_channels, c.input_size,
                                            c.patch_size, h->D, src_mod, h->SS, w_smem ? 1 : 0);
  LDMAE_LAUNCH_CHECK();
  return LDMAE_OK;
}

// The LightningDiT variants without the shipped RMSNorm + SwiGLU pair (reference lightningdit.py:195-224,257-272:
// nn.LayerNorm(elementwise_affine=False, eps=1e-6) block / final norms, nn.LayerNorm(head_dim) q/k norms, timm Mlp with
// tanh-GELU) -- inference only.  The norm is one HBM pass per modulated Linear (norm_modulate_bf16_kernel) instead of being
// folded into the neighbouring GEMM epilogues; everything else runs on the same tcgen05 GEMM / attention kernels with the
// row factor switched off (ssq = nullptr) and the Linear bias in the per-sample-vector slot (cvec_ld = 0).
// Called by dit_forward_impl after the conditioning and the adaLN GEMM (h->mods is ready).
static int dit_forward_generic(ldmae_dit* h, const float* x, float* out, int B, int src_mod, cudaStream_t st) {
  const ldmae_dit_config& c = h->c;
  const int D = h->D, T = h->T, depth = c.depth, M = B * T;
  const float eps = 1e-6f;
  LDMAE_REQUIRE(h->HW == 64, "LayerNorm / GELU variants need head_dim 64");
  {
    ProfScope ps(7, st);
    LDMAE_TRY(launch_patch_embed(h, h->xres.p, nullptr, nullptr, x, nullptr, B, src_mod, st));
  }
  auto norm_mod = [&](int slot) -> int {
    ProfScope ps(7, st);
    const float* w = c.use_rmsnorm ? h->norm_w.p + static_cast<size_t>(slot) * D : nullptr;
    const float* sh = h->so_host[slot] >= 0 ? h->mods.p + h->so_host[slot] : nullptr;
    norm_modulate_bf16_kernel<<<cdiv(M, 8), 256, 0, st>>>(h->abuf.p, h->xres.p, w, h->mods.p + h->sco_host[slot], sh, h->Ntot, T, M, D, eps);
    LDMAE_LAUNCH_CHECK();
    return LDMAE_OK;
  };
  for (int i = 0; i < depth; ++i) {
    DitBlockW& b = h->blk[i];
    const float* mods_i = h->mods.p + static_cast<size_t>(i) * h->nmod * D;
    const float* gate_msa = mods_i + (c.wo_shift ? 1 : 2) * D;
    const float* gate_mlp = mods_i + (c.wo_shift ? 3 : 5) * D;
    LDMAE_TRY(norm_mod(2 * i));
    EpiQKV::Params eq;
    LDMAE_TRY(make_tmap_out_bf16(&eq.omap, h->qkv.p, M, 3 * D, 3 * D));
    eq.has_raw = 0; eq.rawmap = eq.omap;
    eq.ssq = nullptr; eq.cvec = b.b_qkv.p; eq.cvec_ld = 0;
    eq.qw = c.use_qknorm ? b.qw.p : nullptr; eq.kw = c.use_qknorm ? b.kw.p : nullptr;
    const bool ln_heads = c.use_qknorm && !c.use_rmsnorm;
    eq.qb = ln_heads ? b.qb.p : nullptr; eq.kb = ln_heads ? b.kb.p : nullptr;
    eq.rope = c.use_rope ? h->rope_tab.p : nullptr; eq.grid = h->G;
    eq.D = D; eq.rows_per_sample = T; eq.ss_slots = h->SS; eq.inv_D = 1.f / D; eq.eps_row = eps;
    eq.eps_head = ln_heads ? 1e-5f : eps;              // nn.LayerNorm default eps vs models/rmsnorm.py
    eq.q_mul = 1.f;
    { ProfScope ps(0, st); LDMAE_TRY((gemm_auto<EpiQKV>(h->abuf.p, D, b.w_qkv.p, D, GemmShape{M, 3 * D, D}, eq, st))); }
    {
      ProfScope ps(1, st);
      // the score bound of finalize() holds for RMS-normed heads only; LayerNorm heads use the running-maximum kernel
      const float m0 = (c.use_qknorm && c.use_rmsnorm) ? b.attn_m0_log2 : -1.f;
      LDMAE_TRY(run_attention(h->qkv.p, 3 * D, h->obuf.p, D, B, T, c.num_heads, 0, D, 2 * D, 0.125f, st, nullptr, m0, false));
    }
    { ProfScope ps(2, st);
      LDMAE_TRY(gemm_residual(h->obuf.p, D, b.w_proj.p, D, GemmShape{M, D, D}, h->xres.p, D, b.b_proj.p, gate_msa, h->Ntot, nullptr, 0,
                              nullptr, nullptr, 0, T, st)); }
    LDMAE_TRY(norm_mod(2 * i + 1));
    if (c.use_swiglu) {
      EpiSwiGLU::Params es;
      LDMAE_TRY(make_tmap_out_bf16(&es.omap, h->hbuf.p, M, h->Hp, h->Hp));
      es.has_pre = 0; es.premap = es.omap;
      es.ssq = nullptr; es.cvec = b.b12.p; es.cvec_ld = 0;
      es.rows_per_sample = T; es.ss_slots = h->SS; es.inv_D = 1.f / D; es.eps_row = eps;
      ProfScope ps(3, st);
      LDMAE_TRY((gemm_auto<EpiSwiGLU>(h->abuf.p, D, b.w12.p, D, GemmShape{M, 2 * h->Hp, D}, es, st)));
    } else {
      ProfScope ps(3, st);     // timm Mlp: fc1 + GELU(approximate="tanh")
      LDMAE_TRY((gemm_store<__nv_bfloat16, 2>(h->abuf.p, D, b.w12.p, D, GemmShape{M, h->Hp, D}, h->hbuf.p, h->Hp, b.b12.p, st)));
    }
    { ProfScope ps(4, st);
      LDMAE_TRY(gemm_residual(h->hbuf.p, h->Hp, b.w3.p, h->Hp, GemmShape{M, D, h->Hp}, h->xres.p, D, b.b3.p, gate_mlp, h->Ntot, nullptr, 0,
                              nullptr, nullptr, 0, T, st)); }
  }
  LDMAE_TRY(norm_mod(2 * depth));
  EpiFinal::Params ef;
  ef.out = out; ef.ssq = nullptr; ef.cvec = h->b_f.p; ef.cvec_ld = 0; ef.grid = h->G; ef.patch = c.patch_size;
  ef.cout = c.in_channels * (c.learn_sigma ? 2 : 1); ef.cstore = c.in_channels; ef.rows_per_sample = T; ef.ss_slots = h->SS;
  ef.inv_D = 1.f / D; ef.eps_row = eps;
  ProfScope ps(6, st);
  LDMAE_TRY((launch_gemm<16, 1, EpiFinal>(h->abuf.p, D, h->w_f.p, D, GemmShape{M, h->Nf, D}, ef, st)));
  ++g_launch_count;
  return LDMAE_OK;
}

// One forward pass of cat-batch B (see header).  out: [B, Cstore, S, S].  tr != nullptr: training forward (keeps activations).
static int dit_forward_impl(ldmae_dit* h, const float* x, const float* t, float t_scalar, const int64_t* y, float* out,
                            int B, int src_mod, cudaStream_t st, DitTrain* tr = nullptr) {
  LDMAE_REQUIRE(h && h->finalized, "LightningDiT handle not finalized (load all weights, then ldmae_dit_finalize)");
  LDMAE_REQUIRE(B >= 1 && src_mod >= 1, "bad batch");
  LDMAE_TRY(dit_label_check(h));
  ++h->ws_gen;
  if (B > h->maxB) {
    LDMAE_CUDA(cudaStreamSynchronize(st));
    LDMAE_TRY(dit_alloc_ws(h, B));
  }
  const ldmae_dit_config& c = h->c;
  const int D = h->D, T = h->T, depth = c.depth;
  const int M = B * T;
  const size_t MD = static_cast<size_t>(M) * D;
  const float eps = 1e-6f;
  // slot s = norm index (2 per block + final): stream copy, norm operand and row statistics of that norm's input
  auto Xs = [&](int s) { return tr ? tr->X.p + s * MD : h->xres.p; };
  auto As = [&](int s) { return tr ? tr->A.p + s * MD : h->abuf.p; };
  auto Ss = [&](int s) { return tr ? tr->SSQ.p + static_cast<size_t>(s) * M * h->SS : h->ssq.p; };
  int dbg_stage = 0;
#define LDMAE_DBG_STAGE() do { if (h->debug_stop >= 0 && ++dbg_stage > h->debug_stop) return LDMAE_OK; } while (0)
  // 1. conditioning c = t_emb + y_emb ; sc = bf16(silu(c))
  {
    ProfScope ps(7, st);
    dim3 grid(cdiv(D, 64), cdiv(B, 16));
    const size_t sm0 = 16 * 256 * sizeof(float), sm2 = 16 * static_cast<size_t>(D) * sizeof(float);
    static PerDeviceOnce attr;
    if (attr.pending()) {
      LDMAE_CUDA(cudaFuncSetAttribute(small_linear_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      LDMAE_CUDA(cudaFuncSetAttribute(small_linear_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr.mark();
    }
    if (tr) {
      // the backward needs the pre-activation of the timestep MLP and its input features
      timestep_features_kernel<<<cdiv(static_cast<size_t>(B) * 256, 256), 256, 0, st>>>(tr->tfreq.p, t, t_scalar, B, 256);
      LDMAE_LAUNCH_CHECK();
      small_linear_kernel<0><<<grid, 256, sm0, st>>>(tr->th1pre.p, nullptr, t, t_scalar, h->t_w0.p, h->t_b0.p, nullptr, nullptr, B, 256, D, 1, 256);
      LDMAE_LAUNCH_CHECK();
      silu_f32_kernel<<<cdiv(static_cast<size_t>(B) * D, 256), 256, 0, st>>>(h->th1.p, tr->th1pre.p, static_cast<size_t>(B) * D);
      LDMAE_LAUNCH_CHECK();
      LDMAE_CUDA(cudaMemcpyAsync(tr->ykeep.p, y, static_cast<size_t>(B) * sizeof(long long), cudaMemcpyDeviceToDevice, st));
    } else {
      small_linear_kernel<1><<<grid, 256, sm0, st>>>(h->th1.p, nullptr, t, t_scalar, h->t_w0.p, h->t_b0.p, nullptr, nullptr, B, 256, D, 1, 256);
      LDMAE_LAUNCH_CHECK();
    }
    small_linear_kernel<0><<<grid, 256, sm2, st>>>(h->cvec_c.p, h->th1.p, nullptr, 0.f, h->t_w2.p, h->t_b2.p, h->emb.p,
                                                   reinterpret_cast<const long long*>(y), B, D, D, 0, D, c.num_embeddings,
                                                   h->label_err.p);
    LDMAE_LAUNCH_CHECK();
    LDMAE_CUDA(cudaMemcpyAsync(h->label_err_host, h->label_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    silu_to_bf16_kernel<<<cdiv(static_cast<size_t>(B) * D, 256), 256, 0, st>>>(h->sc.p, h->cvec_c.p, static_cast<size_t>(B) * D);
    LDMAE_LAUNCH_CHECK();
  }
  LDMAE_DBG_STAGE();   // 1
  // 2. all adaLN modulations at once: mods[B, Ntot] = sc . W_ada^T + b_ada
  {
    ProfScope ps(5, st);
    LDMAE_TRY((gemm_store<float, 0>(h->sc.p, D, h->w_ada.p, D, GemmShape{B, h->Ntot, D}, h->mods.p, h->Ntot, h->b_ada.p, st)));
    if (!h->fused) return dit_forward_generic(h, x, out, B, src_mod, st);
    const size_t tot = static_cast<size_t>(h->S) * B * D;
    adaln_prep_kernel<<<cdiv(tot, 256), 256, 0, st>>>(tr ? h->shift_bf16.p : nullptr, h->gmul.p, h->mods.p, h->norm_w.p,
                                                      h->slot_shift_off.p, h->slot_scale_off.p, B, D, h->Ntot, h->S);
    LDMAE_LAUNCH_CHECK();
  }
  LDMAE_DBG_STAGE();   // 2
  // 3. per-sample vectors  shift_b . W^T + bias  for every modulated Linear
  const bool fused_shift = tr == nullptr;
  if (fused_shift) {
    // inference: one GEMM on silu(c) against the pre-multiplied matrices (dit_build_fused_shift)
    ProfScope ps(5, st);
    LDMAE_TRY(dit_build_fused_shift(h, st));
    LDMAE_TRY((gemm_store<float, 0>(h->sc.p, D, h->w_shf.p, D, GemmShape{B, h->Nsh, D}, h->cvec_all.p, h->Nsh, h->b_shf.p, st)));
  } else {
    for (int i = 0; i < depth; ++i) {
      ProfScope ps(5, st);
      DitBlockW& b = h->blk[i];
      float* cq = h->cvec_qkv.p + static_cast<size_t>(i) * B * 3 * h->QW;
      float* c12 = h->cvec_12.p + static_cast<size_t>(i) * B * 2 * h->Hp;
      LDMAE_TRY((gemm_store<float, 0>(h->shift_bf16.p + static_cast<size_t>(2 * i) * B * D, D, b.w_qkv.p, D, GemmShape{B, 3 * h->QW, D},
                                      cq, 3 * h->QW, b.b_qkv.p, st)));
      LDMAE_TRY((gemm_store<float, 0>(h->shift_bf16.p + static_cast<size_t>(2 * i + 1) * B * D, D, b.w12.p, D,
                                      GemmShape{B, 2 * h->Hp, D}, c12, 2 * h->Hp, b.b12.p, st)));
    }
  }
  const int sh_per_blk = 3 * h->QW + 2 * h->Hp;
  auto cvec_q = [&](int i) { return fused_shift ? h->cvec_all.p + static_cast<size_t>(i) * sh_per_blk : h->cvec_qkv.p + static_cast<size_t>(i) * B * 3 * h->QW; };
  auto cvec_m = [&](int i) { return fused_shift ? h->cvec_all.p + static_cast<size_t>(i) * sh_per_blk + 3 * h->QW : h->cvec_12.p + static_cast<size_t>(i) * B * 2 * h->Hp; };
  const int cvq_ld = fused_shift ? h->Nsh : 3 * h->QW, cvm_ld = fused_shift ? h->Nsh : 2 * h->Hp;
  if (!fused_shift) {
    ProfScope ps(7, st);
    // final linear (N = p*p*C_out, e.g. 16): fp32 on CUDA cores straight from the shift columns of mods
    dim3 grid(cdiv(h->Nf, 64), cdiv(B, 16));
    const size_t sm2 = 16 * static_cast<size_t>(D) * sizeof(float);
    small_linear_kernel<0><<<grid, 256, sm2, st>>>(h->cvec_f.p, h->mods.p + static_cast<size_t>(depth) * h->nmod * D, nullptr, 0.f,
                                                   h->w_f32.p, h->b_f.p, nullptr, nullptr, B, D, h->Nf, 0, h->Ntot);
    LDMAE_LAUNCH_CHECK();
  }
  LDMAE_DBG_STAGE();   // 3
  // 4. patch embed + pos embed -> residual stream, first operand, first row statistics
  {
    ProfScope ps(7, st);
    LDMAE_TRY(launch_patch_embed(h, Xs(0), As(0), Ss(0), x, h->gmul.p, B, src_mod, st, tr == nullptr));
    if (tr) {
      LDMAE_REQUIRE(src_mod == B, "training forward takes a plain batch");
      patchify_bf16_kernel<<<cdiv(static_cast<size_t>(M) * h->Kp, 256), 256, 0, st>>>(tr->XTOK.p, x, B, c.in_channels, c.input_size, c.patch_size);
      LDMAE_LAUNCH_CHECK();
    }
  }
  LDMAE_DBG_STAGE();   // 4
  // 5. blocks  (debug stages 5 + 5*i + {0 qkv, 1 attention, 2 proj, 3 w12, 4 w3})
  for (int i = 0; i < depth; ++i) {
    DitBlockW& b = h->blk[i];
    const float* mods_i = h->mods.p + static_cast<size_t>(i) * h->nmod * D;
    const float* gate_msa = mods_i + (c.wo_shift ? 1 : 2) * D;
    const float* gate_mlp = mods_i + (c.wo_shift ? 3 : 5) * D;
    __nv_bfloat16* qkv_i = tr ? tr->QKV.p + static_cast<size_t>(i) * M * 3 * h->QW : h->qkv.p;
    __nv_bfloat16* o_i = tr ? tr->O.p + i * MD : h->obuf.p;
    __nv_bfloat16* hb_i = tr ? tr->HB.p + static_cast<size_t>(i) * M * h->Hp : h->hbuf.p;
    if (h->HW == 64) {
      EpiQKV::Params eq;
      LDMAE_TRY(make_tmap_out_bf16(&eq.omap, qkv_i, M, 3 * D, 3 * D));
      eq.has_raw = 0; eq.rawmap = eq.omap;
      if (tr) {
        eq.has_raw = 1;
        LDMAE_TRY(make_tmap_out_bf16(&eq.rawmap, tr->QKR.p + static_cast<size_t>(i) * M * 2 * D, M, 2 * D, 2 * D));
      }
      eq.ssq = Ss(2 * i); eq.cvec = cvec_q(i); eq.cvec_ld = cvq_ld;
      eq.qw = c.use_qknorm ? b.qw.p : nullptr; eq.kw = c.use_qknorm ? b.kw.p : nullptr; eq.qb = nullptr; eq.kb = nullptr;
      eq.rope = c.use_rope ? h->rope_tab.p : nullptr; eq.grid = h->G;
      eq.D = D; eq.rows_per_sample = T; eq.ss_slots = h->SS; eq.inv_D = 1.f / D; eq.eps_row = eps; eq.eps_head = eps;
      // inference forward of qk-normed heads: the softmax scale and log2(e) ride on q_norm.weight, the attention kernel takes
      // 2^score directly (no per-score multiply-add); the training forward keeps the plain q the backward kernels expect
      const bool presc = tr == nullptr && c.use_qknorm && b.attn_m0_log2 > 0.f;
      eq.q_mul = presc ? 0.125f * 1.4426950408889634f : 1.f;
      { ProfScope ps(0, st); LDMAE_TRY((gemm_auto<EpiQKV>(As(2 * i), D, b.w_qkv.p, D, GemmShape{M, 3 * D, D}, eq, st))); }
      LDMAE_DBG_STAGE();
      {
        ProfScope ps(1, st);
        float* lse = tr ? tr->LSE.p + static_cast<size_t>(i) * B * c.num_heads * T : nullptr;
        LDMAE_TRY(run_attention(qkv_i, 3 * D, o_i, D, B, T, c.num_heads, 0, D, 2 * D, 0.125f, st, lse, b.attn_m0_log2, presc));
      }
    } else {
      // wider heads (XL, head_dim 72): 128-column head slots, generic epilogue and the one-tile attention kernel
      EpiQKVWide::Params eq;
      LDMAE_TRY(make_tmap_out_bf16(&eq.omap, qkv_i, M, 3 * h->QW, 3 * h->QW));
      eq.has_raw = 0; eq.rawmap = eq.omap;
      if (tr) {
        eq.has_raw = 1;
        LDMAE_TRY(make_tmap_out_bf16(&eq.rawmap, tr->QKR.p + static_cast<size_t>(i) * M * 2 * h->QW, M, 2 * h->QW, 2 * h->QW));
      }
      eq.ssq = Ss(2 * i); eq.cvec = cvec_q(i); eq.cvec_ld = cvq_ld;
      eq.qw = c.use_qknorm ? b.qw.p : nullptr; eq.kw = c.use_qknorm ? b.kw.p : nullptr;
      eq.rope_cos = c.use_rope ? h->rope_cos.p : nullptr; eq.rope_sin = c.use_rope ? h->rope_sin.p : nullptr;
      eq.rope_tab = (c.use_rope && h->rope_wide_ok) ? h->rope_tab.p : nullptr; eq.grid = h->G;
      eq.section = h->QW; eq.hd = h->hd; eq.rows_per_sample = T; eq.ss_slots = h->SS;
      eq.inv_D = 1.f / D; eq.eps_row = eps; eq.eps_head = eps;
      {
        // multiply only the live rows of every 128-row head slot of the packed weight (the rest are zeros): 16 * ceil(hd / 16)
        // of 128 -- the two-CTA tile (one head slot per CTA) then has N = 2 * n_live; LDMAE_QKV_WIDE_LIVE=0 keeps N = 256
        static int live_on = -1;
        if (live_on < 0) { const char* e = getenv("LDMAE_QKV_WIDE_LIVE"); live_on = e ? atoi(e) : 1; }
        GemmShape gs{M, 3 * h->QW, D};
        const bool two_cta = gemm_cg() == 2 && M > 128;
        if (live_on && two_cta) gs.n_live = 16 * ((h->hd + 15) / 16);
        eq.acc_stride = gs.n_live > 0 ? gs.n_live : 128;
        ProfScope ps(0, st);
        LDMAE_TRY((gemm_auto<EpiQKVWide>(As(2 * i), D, b.w_qkv.p, D, gs, eq, st)));
      }
      LDMAE_DBG_STAGE();
      {
        ProfScope ps(1, st);
        float* lse = tr ? tr->LSE.p + static_cast<size_t>(i) * B * c.num_heads * T : nullptr;
        LDMAE_TRY(run_attention_hd128(qkv_i, 3 * h->QW, o_i, D, B, T, c.num_heads, h->hd, 0, h->QW, 2 * h->QW,
                                      1.0f / sqrtf(static_cast<float>(h->hd)), st, lse, c.use_qknorm ? b.attn_m0_log2 : -1.f));
      }
    }
    LDMAE_DBG_STAGE();
    {
      ProfScope ps(2, st);
      if (tr)
        LDMAE_TRY(gemm_residual_train(o_i, D, b.w_proj.p, D, GemmShape{M, D, D}, Xs(2 * i), Xs(2 * i + 1), D, b.b_proj.p, gate_msa,
                                      h->Ntot, h->gmul.p + static_cast<size_t>(2 * i + 1) * B * D, D, As(2 * i + 1), Ss(2 * i + 1),
                                      h->SS, tr->MA.p + i * MD, T, st));
      else
        LDMAE_TRY(gemm_residual(o_i, D, b.w_proj.p, D, GemmShape{M, D, D}, h->xres.p, D, b.b_proj.p, gate_msa, h->Ntot,
                                h->gmul.p + static_cast<size_t>(2 * i + 1) * B * D, D, h->abuf.p, h->ssq.p, h->SS, T, st));
    }
    LDMAE_DBG_STAGE();
    EpiSwiGLU::Params es;
    LDMAE_TRY(make_tmap_out_bf16(&es.omap, hb_i, M, h->Hp, h->Hp));
    es.has_pre = 0; es.premap = es.omap;
    if (tr) {
      es.has_pre = 1;
      LDMAE_TRY(make_tmap_out_bf16(&es.premap, tr->H12.p + static_cast<size_t>(i) * M * 2 * h->Hp, M, 2 * h->Hp, 2 * h->Hp));
    }
    es.ssq = Ss(2 * i + 1); es.cvec = cvec_m(i); es.cvec_ld = cvm_ld;
    es.rows_per_sample = T; es.ss_slots = h->SS; es.inv_D = 1.f / D; es.eps_row = eps;
    { ProfScope ps(3, st); LDMAE_TRY((gemm_auto<EpiSwiGLU>(As(2 * i + 1), D, b.w12.p, D, GemmShape{M, 2 * h->Hp, D}, es, st))); }
    LDMAE_DBG_STAGE();
    {
      ProfScope ps(4, st);
      if (tr)
        LDMAE_TRY(gemm_residual_train(hb_i, h->Hp, b.w3.p, h->Hp, GemmShape{M, D, h->Hp}, Xs(2 * i + 1), Xs(2 * i + 2), D, b.b3.p,
                                      gate_mlp, h->Ntot, h->gmul.p + static_cast<size_t>(2 * i + 2) * B * D, D, As(2 * i + 2),
                                      Ss(2 * i + 2), h->SS, tr->MM.p + i * MD, T, st));
      else
        LDMAE_TRY(gemm_residual(h->hbuf.p, h->Hp, b.w3.p, h->Hp, GemmShape{M, D, h->Hp}, h->xres.p, D, b.b3.p, gate_mlp, h->Ntot,
                                h->gmul.p + static_cast<size_t>(2 * i + 2) * B * D, D, h->abuf.p, h->ssq.p, h->SS, T, st));
    }
    LDMAE_DBG_STAGE();
  }
  // 6. final layer + unpatchify
  {
    EpiFinal::Params ef;
    ef.out = out; ef.ssq = Ss(2 * depth); ef.grid = h->G; ef.patch = c.patch_size;
    ef.cvec = fused_shift ? h->cvec_all.p + static_cast<size_t>(depth) * sh_per_blk : h->cvec_f.p;
    ef.cvec_ld = fused_shift ? h->Nsh : h->Nf;
    ef.cout = c.in_channels * (c.learn_sigma ? 2 : 1); ef.cstore = c.in_channels; ef.rows_per_sample = T; ef.ss_slots = h->SS;
    ef.inv_D = 1.f / D; ef.eps_row = eps;
    ProfScope ps(6, st);
    LDMAE_TRY((launch_gemm<16, 1, EpiFinal>(As(2 * depth), D, h->w_f.p, D, GemmShape{M, h->Nf, D}, ef, st)));
    ++g_launch_count;
  }
  return LDMAE_OK;
}

// Debug hooks for the parity tests: stop the forward after `stages` launch groups / read a workspace buffer.
extern "C" int ldmae_dit_debug_stop(ldmae_dit* h, int32_t stages) {
  LDMAE_REQUIRE(h, "null handle");
  h->debug_stop = stages;
  return LDMAE_OK;
}
// Fills every workspace buffer with `byte` (0xFF = NaN patterns): a read-before-write then shows up as NaNs.
extern "C" int ldmae_dit_debug_poison(ldmae_dit* h, int32_t byte, void* stream) {
  LDMAE_REQUIRE(h, "null handle");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LDMAE_POISON(BUF) if (h->BUF.p) LDMAE_CUDA(cudaMemsetAsync(h->BUF.p, byte, h->BUF.n * sizeof(*h->BUF.p), st));
  LDMAE_POISON(xres) LDMAE_POISON(ssq) LDMAE_POISON(cvec_c) LDMAE_POISON(th1) LDMAE_POISON(mods) LDMAE_POISON(gmul)
  LDMAE_POISON(cvec_qkv) LDMAE_POISON(cvec_12) LDMAE_POISON(cvec_f) LDMAE_POISON(vbuf) LDMAE_POISON(k1buf) LDMAE_POISON(xtmp)
  LDMAE_POISON(abuf) LDMAE_POISON(qkv) LDMAE_POISON(obuf) LDMAE_POISON(hbuf) LDMAE_POISON(sc) LDMAE_POISON(shift_bf16)
  LDMAE_POISON(cvec_all)
#undef LDMAE_POISON
  return LDMAE_OK;
}
extern "C" int ldmae_dit_debug_read(ldmae_dit* h, const char* name, void* dst, int64_t nbytes, void* stream) {
  LDMAE_REQUIRE(h && name && dst, "null argument");
  const std::string k(name);
  const void* src = nullptr;
  size_t have = 0;
#define LDMAE_DBG_BUF(NAME, BUF) if (k == NAME) { src = h->BUF.p; have = h->BUF.n * sizeof(*h->BUF.p); }
  LDMAE_DBG_BUF("xres", xres) LDMAE_DBG_BUF("abuf", abuf) LDMAE_DBG_BUF("qkv", qkv) LDMAE_DBG_BUF("obuf", obuf)
  LDMAE_DBG_BUF("hbuf", hbuf) LDMAE_DBG_BUF("ssq", ssq) LDMAE_DBG_BUF("mods", mods) LDMAE_DBG_BUF("cvec_c", cvec_c)
  LDMAE_DBG_BUF("cvec_qkv", cvec_qkv) LDMAE_DBG_BUF("cvec_12", cvec_12) LDMAE_DBG_BUF("cvec_f", cvec_f)
  LDMAE_DBG_BUF("gmul", gmul) LDMAE_DBG_BUF("shift_bf16", shift_bf16) LDMAE_DBG_BUF("sc", sc) LDMAE_DBG_BUF("th1", th1)
#undef LDMAE_DBG_BUF
  LDMAE_REQUIRE(src != nullptr, "unknown debug buffer %s", name);
  LDMAE_REQUIRE(nbytes >= 0 && static_cast<size_t>(nbytes) <= have, "debug buffer %s holds %zu bytes, asked for %lld", name, have, (long long)nbytes);
  LDMAE_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return LDMAE_OK;
}

// Waits for the stream and reports a class label that fell outside the embedding table in any call since the last check.
extern "C" int ldmae_dit_check_labels(ldmae_dit* h, void* stream) {
  LDMAE_REQUIRE(h, "null handle");
  LDMAE_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return dit_label_check(h);
}
// Generation of the handle's shared workspace: bumped by every forward (inference or training) and every workspace
// re-allocation.  A caller that keeps a training forward pending (autograd) records it and compares before the backward.
extern "C" long long ldmae_dit_generation(ldmae_dit* h) { return h ? h->ws_gen : -1; }

extern "C" int ldmae_dit_forward(ldmae_dit* h, const float* x, const float* t, float t_scalar, const int64_t* y,
                                 float* out, int32_t B, int32_t src_mod, void* stream) {
  LDMAE_REQUIRE(x && y && out, "null tensor");
  return dit_forward_impl(h, x, t, t_scalar, y, out, B, src_mod, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------- LightningDiT training
// Backward of LightningDiT.forward for transport.training_losses / train_accum.py:215-230.
#include "dit_train.inc"

// Trainer input pipeline + loss (see train_io.cuh).  All pointers are device pointers; HW = S*S, n = C*HW.
extern "C" int ldmae_flow_prepare(const float* moments, const float* moments_flip, const uint8_t* flip, const float* eps_post,
                                  const float* mean, const float* stdv, float multiplier, const float* x1_in, const float* x0,
                                  const float* t, float* x1_out, float* xt, float* ut, int32_t B, int32_t C, int32_t HW, void* stream) {
  LDMAE_REQUIRE((moments != nullptr) != (x1_in != nullptr), "give either the posterior moments or ready latents x1_in");
  LDMAE_REQUIRE(x0 && t && xt && ut && B >= 1 && C >= 1 && HW >= 4 && HW % 4 == 0, "bad argument");
  LDMAE_REQUIRE((mean == nullptr) == (stdv == nullptr), "mean/std must be given together");
  LDMAE_REQUIRE(!(flip != nullptr && moments_flip == nullptr), "a flip mask needs the flipped moments");
  const size_t total4 = static_cast<size_t>(B) * C * HW / 4;
  flow_prepare_kernel<<<cdiv(total4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      moments, moments_flip, flip, eps_post, mean, stdv, multiplier, x1_in, x0, t, x1_out, xt, ut, B, C, HW);
  LDMAE_LAUNCH_CHECK();
  return LDMAE_OK;
}
extern "C" int ldmae_flow_loss(const float* out, const float* ut, float* loss, float* dout, float loss_scale, int32_t B, int32_t n,
                               void* stream) {
  LDMAE_REQUIRE(out && ut && loss && B >= 1 && n >= 4 && n % 4 == 0, "bad argument");
  flow_loss_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, ut, loss, dout,
                                                                    2.0f * loss_scale / (static_cast<float>(n) * static_cast<float>(B)), n);
  LDMAE_LAUNCH_CHECK();
  return LDMAE_OK;
}

static int launch_update(float* xout, const float* xin, const float* v, const float* kprev, float* gout, int n_half,
                         int C, int HW, float cfg_scale, int use_guidance, float a, float bcoef, size_t total,
                         cudaStream_t st) {
  ProfScope ps(7, st);
  cfg_ode_update_kernel<<<cdiv(total, 256), 256, 0, st>>>(xout, xin, v, kprev, gout, n_half, C, HW, 3, cfg_scale,
                                                          use_guidance, a, bcoef, total);
  LDMAE_LAUNCH_CHECK();
  return LDMAE_OK;
}

extern "C" int ldmae_dit_forward_with_cfg(ldmae_dit* h, const float* x, const float* t, float t_scalar,
                                          const int64_t* y, float* out, int32_t n, float cfg_scale,
                                          int32_t use_guidance, void* stream) {
  LDMAE_REQUIRE(h && x && y && out && n >= 1, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = h->c.in_channels, HW = h->c.input_size * h->c.input_size;
  const size_t total = static_cast<size_t>(2 * n) * C * HW;
  if (2 * n > h->maxB) { LDMAE_CUDA(cudaStreamSynchronize(st)); LDMAE_TRY(dit_alloc_ws(h, 2 * n)); }
  LDMAE_TRY(dit_forward_impl(h, x, t, t_scalar, y, h->vbuf.p, 2 * n, n, st));
  return launch_update(nullptr, nullptr, h->vbuf.p, nullptr, out, n, C, HW, cfg_scale, use_guidance, 0.f, 0.f, total, st);
}

extern "C" int ldmae_sample_ode(ldmae_dit* h, float* x, const int64_t* y, int32_t n, int32_t use_cfg, float cfg_scale,
                                float cfg_interval_start, const float* tgrid, int32_t npts, int32_t method, float* traj,
                                int32_t flags, void* stream) {
  LDMAE_REQUIRE(h && x && y && tgrid && n >= 1 && npts >= 2, "bad argument");
  LDMAE_REQUIRE(method == 0 || method == 1, "method must be 0 (euler) or 1 (heun2)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = h->c.in_channels, HW = h->c.input_size * h->c.input_size;
  const int Btot = use_cfg ? 2 * n : n;
  const int src_mod = n;
  const int n_half = use_cfg ? n : 0;
  const size_t total = static_cast<size_t>(Btot) * C * HW;
  const size_t half_total = static_cast<size_t>(n) * C * HW;
  const bool cond_only = use_cfg && (flags & LDMAE_ODE_COND_ONLY_WHEN_UNGUIDED) != 0;
  LDMAE_REQUIRE(!(cond_only && traj), "the cond-only shortcut does not keep the unconditional half: no trajectory output");
  if (Btot > h->maxB) { LDMAE_CUDA(cudaStreamSynchronize(st)); LDMAE_TRY(dit_alloc_ws(h, Btot)); }
  if (traj) LDMAE_CUDA(cudaMemcpyAsync(traj, x, total * sizeof(float), cudaMemcpyDeviceToDevice, st));
  auto guided = [&](float tt) { return (use_cfg && !(cfg_interval_start >= 0.f && tt < cfg_interval_start)) ? 1 : 0; };
  // One drift evaluation at time tt on state xs, then xout = xin + a*g + b*kprev (g = guided velocity, optionally kept in gout).
  // Below the guidance interval the guided velocity of the conditional half is its own prediction (lightningdit.py:436-439),
  // so with LDMAE_ODE_COND_ONLY_WHEN_UNGUIDED only the n conditional samples are evaluated and advanced.
  auto stage = [&](const float* xs, float tt, float* xout, const float* xin, const float* kprev, float* gout, float a,
                   float bcoef) -> int {
    if (cond_only && !guided(tt)) {
      LDMAE_TRY(dit_forward_impl(h, xs, nullptr, tt, y, h->vbuf.p, n, src_mod, st));
      return launch_update(xout, xin, h->vbuf.p, kprev, gout, 0, C, HW, cfg_scale, 0, a, bcoef, half_total, st);
    }
    LDMAE_TRY(dit_forward_impl(h, xs, nullptr, tt, y, h->vbuf.p, Btot, src_mod, st));
    return launch_update(xout, xin, h->vbuf.p, kprev, gout, n_half, C, HW, cfg_scale, guided(tt), a, bcoef, total, st);
  };
  for (int k = 0; k + 1 < npts; ++k) {
    const float t0 = tgrid[k], t1 = tgrid[k + 1];
    const float dt = t1 - t0;
    if (method == 0) {
      LDMAE_TRY(stage(x, t0, x, x, nullptr, nullptr, dt, 0.f));
    } else {
      // k1 = f(t0, x); xtmp = x + dt*k1; k2 = f(t0+dt, xtmp); x += dt*(k1/2 + k2/2)
      LDMAE_TRY(stage(x, t0, h->xtmp.p, x, nullptr, h->k1buf.p, dt, 0.f));
      LDMAE_TRY(stage(h->xtmp.p, t0 + dt, x, x, h->k1buf.p, nullptr, 0.5f * dt, 0.5f * dt));
    }
    if (traj)
      LDMAE_CUDA(cudaMemcpyAsync(traj + static_cast<size_t>(k + 1) * total, x, total * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return LDMAE_OK;
}

// ------------------------------------------------------------------------------------------- VMAE decoder
struct VmaeBlockW {
  DevBuf<__nv_bfloat16> w_qkv, w_proj, w_fc1, w_fc2;
  DevBuf<float> b_qkv, b_proj, b_fc1, b_fc2, n1w, n1b, n2w, n2b;
};
struct ldmae_vmae {
  ldmae_vmae_config c;
  int D, L, G, Hm, HP /*padded heads width nh*64*/, PP, maxB;
  DevBuf<__nv_bfloat16> w_from, w_embed, w_pred;
  DevBuf<float> b_from, b_embed, pos, nfw, nfb, b_pred, conv_w, conv_b;
  std::vector<VmaeBlockW> blk;
  std::vector<std::string> loaded;
  bool finalized = false;
  // encoder (tokenizer/models_mae.py:819-836): patch embed, pos embed, ViT blocks, norm, to_latent
  int HPe = 0, NL = 0;                       // padded heads width of the encoder, to_latent outputs (2 * latent with KL)
  DevBuf<__nv_bfloat16> w_patch, w_tolat, etok;
  DevBuf<float> b_patch, epos, enw, enb, b_tolat;
  std::vector<VmaeBlockW> eblk;
  std::vector<std::string> enc_loaded;
  DevBuf<float> x, pred;
  DevBuf<__nv_bfloat16> tok, t1, a, qkv, o, hid;
};

static int vmae_alloc_ws(ldmae_vmae* h, int B) {
  const size_t M = static_cast<size_t>(B) * h->L;
  LDMAE_TRY(h->x.alloc(M * h->D));
  LDMAE_TRY(h->pred.alloc(M * h->PP));
  LDMAE_TRY(h->tok.alloc(M * 64));
  LDMAE_TRY(h->t1.alloc(M * h->c.embed_dim));
  LDMAE_TRY(h->a.alloc(M * h->D));
  LDMAE_TRY(h->qkv.alloc(M * 3 * std::max(h->HP, h->HPe)));
  LDMAE_TRY(h->o.alloc(M * std::max(h->HP, h->HPe)));
  LDMAE_TRY(h->hid.alloc(M * h->Hm));
  if (h->HPe > 0) LDMAE_TRY(h->etok.alloc(M * h->PP));
  h->maxB = B;
  return LDMAE_OK;
}

extern "C" int ldmae_vmae_create(const ldmae_vmae_config* cfg, ldmae_vmae** out) {
  LDMAE_REQUIRE(cfg && out, "null argument");
  LDMAE_TRY(require_sm100());
  const ldmae_vmae_config& c = *cfg;
  LDMAE_REQUIRE(c.decoder_embed_dim % c.decoder_num_heads == 0, "decoder dim %% heads");
  const int hd = c.decoder_embed_dim / c.decoder_num_heads;
  LDMAE_REQUIRE(hd <= 64, "VMAE decoder head_dim %d > 64 not supported", hd);
  LDMAE_REQUIRE(c.latent_dim <= 64, "latent_dim > 64 not supported");
  LDMAE_REQUIRE(c.decoder_embed_dim % 8 == 0 && c.embed_dim % 8 == 0 && c.mlp_hidden % 8 == 0, "dims must be multiples of 8");
  ldmae_vmae* h = new ldmae_vmae();
  h->c = c;
  h->D = c.decoder_embed_dim;
  h->G = c.img_size / c.patch_size;
  h->L = h->G * h->G;
  h->Hm = c.mlp_hidden;
  h->HP = c.decoder_num_heads * 64;
  h->PP = c.patch_size * c.patch_size * 3;
  const int D = h->D;
  int r = LDMAE_OK;
  auto A = [&](int rc) { if (r == LDMAE_OK) r = rc; };
  if (c.depth > 0) {
    // encoder side; it shares the decoder's workspace, so the widths must agree (they do in every shipped factory)
    LDMAE_REQUIRE(c.embed_dim == D && c.num_heads > 0 && c.embed_dim % c.num_heads == 0 && c.embed_dim / c.num_heads <= 64,
                  "VMAE encoder: embed_dim %d must equal decoder_embed_dim %d with head_dim <= 64", c.embed_dim, D);
    LDMAE_REQUIRE(h->PP % 8 == 0, "VMAE encoder: patch_size^2 * 3 must be a multiple of 8");
    h->HPe = c.num_heads * 64;
    h->NL = c.to_latent_dim;
    LDMAE_REQUIRE(h->NL > 0 && h->NL % 4 == 0 && h->NL <= h->PP, "VMAE encoder: bad to_latent width %d", h->NL);
    A(h->w_patch.alloc(static_cast<size_t>(D) * h->PP)); A(h->b_patch.alloc(D));
    A(h->epos.alloc(static_cast<size_t>(h->L) * D));
    A(h->enw.alloc(D)); A(h->enb.alloc(D));
    A(h->w_tolat.alloc(static_cast<size_t>(h->NL) * D)); A(h->b_tolat.alloc(h->NL));
    h->eblk.resize(c.depth);
    for (auto& b : h->eblk) {
      A(b.w_qkv.alloc(static_cast<size_t>(3 * h->HPe) * D)); A(b.b_qkv.alloc(3 * h->HPe, true));
      A(b.w_proj.alloc(static_cast<size_t>(D) * h->HPe)); A(b.b_proj.alloc(D));
      A(b.w_fc1.alloc(static_cast<size_t>(h->Hm) * D)); A(b.b_fc1.alloc(h->Hm));
      A(b.w_fc2.alloc(static_cast<size_t>(D) * h->Hm)); A(b.b_fc2.alloc(D));
      A(b.n1w.alloc(D)); A(b.n1b.alloc(D)); A(b.n2w.alloc(D)); A(b.n2b.alloc(D));
    }
  }
  A(h->w_from.alloc(static_cast<size_t>(c.embed_dim) * 64)); A(h->b_from.alloc(c.embed_dim));
  A(h->w_embed.alloc(static_cast<size_t>(D) * c.embed_dim)); A(h->b_embed.alloc(D));
  A(h->pos.alloc(static_cast<size_t>(h->L) * D));
  A(h->nfw.alloc(D)); A(h->nfb.alloc(D));
  A(h->w_pred.alloc(static_cast<size_t>(h->PP) * D)); A(h->b_pred.alloc(h->PP));
  A(h->conv_w.alloc(81)); A(h->conv_b.alloc(3));
  h->blk.resize(c.decoder_depth);
  for (auto& b : h->blk) {
    A(b.w_qkv.alloc(static_cast<size_t>(3 * h->HP) * D)); A(b.b_qkv.alloc(3 * h->HP, true));
    A(b.w_proj.alloc(static_cast<size_t>(D) * h->HP)); A(b.b_proj.alloc(D));
    A(b.w_fc1.alloc(static_cast<size_t>(h->Hm) * D)); A(b.b_fc1.alloc(h->Hm));
    A(b.w_fc2.alloc(static_cast<size_t>(D) * h->Hm)); A(b.b_fc2.alloc(D));
    A(b.n1w.alloc(D)); A(b.n1b.alloc(D)); A(b.n2w.alloc(D)); A(b.n2b.alloc(D));
  }
  if (r == LDMAE_OK) r = vmae_alloc_ws(h, std::max(1, c.max_batch));
  if (r != LDMAE_OK) { delete h; return r; }
  *out = h;
  return LDMAE_OK;
}
extern "C" void ldmae_vmae_destroy(ldmae_vmae* h) { delete h; }

extern "C" int ldmae_vmae_load_tensor(ldmae_vmae* h, const char* name, const float* data, int64_t numel, void* stream) {
  LDMAE_REQUIRE(h && name && data, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const std::string k(name);
  const int D = h->D, E = h->c.embed_dim, nh = h->c.decoder_num_heads;
  int rc = LDMAE_OK;
  int bi = -1;
  bool enc = false;
  std::string sub;
  int nhb = nh, HPb = h->HP;                 // head geometry of the block being loaded
  if (k.rfind("decoder_blocks.", 0) == 0) {
    const size_t dot = k.find('.', 15);
    LDMAE_REQUIRE(dot != std::string::npos, "bad key %s", name);
    bi = atoi(k.substr(15, dot - 15).c_str());
    sub = k.substr(dot + 1);
    LDMAE_REQUIRE(bi >= 0 && bi < h->c.decoder_depth, "block index out of range in %s", name);
  } else if (k.rfind("blocks.", 0) == 0) {
    LDMAE_REQUIRE(h->HPe > 0, "VMAE handle was created without an encoder (key %s)", name);
    const size_t dot = k.find('.', 7);
    LDMAE_REQUIRE(dot != std::string::npos, "bad key %s", name);
    bi = atoi(k.substr(7, dot - 7).c_str());
    sub = k.substr(dot + 1);
    LDMAE_REQUIRE(bi >= 0 && bi < h->c.depth, "block index out of range in %s", name);
    enc = true; nhb = h->c.num_heads; HPb = h->HPe;
  }
  const int hdb = D / nhb;
  const bool enc_key = enc || k == "patch_embed.proj.weight" || k == "patch_embed.proj.bias" || k == "pos_embed" ||
                       k == "norm.weight" || k == "norm.bias" || k == "to_latent.weight" || k == "to_latent.bias";
  if (enc_key && !enc) LDMAE_REQUIRE(h->HPe > 0, "VMAE handle was created without an encoder (key %s)", name);
  if (k == "patch_embed.proj.weight") rc = pack_bf16(h->w_patch.p, data, D, h->PP, h->PP, numel, name, st);
  else if (k == "patch_embed.proj.bias") rc = copy_f32(h->b_patch.p, data, numel, D, name, st);
  else if (k == "pos_embed") rc = copy_f32(h->epos.p, data, numel, (int64_t)h->L * D, name, st);
  else if (k == "norm.weight") rc = copy_f32(h->enw.p, data, numel, D, name, st);
  else if (k == "norm.bias") rc = copy_f32(h->enb.p, data, numel, D, name, st);
  else if (k == "to_latent.weight") rc = pack_bf16(h->w_tolat.p, data, h->NL, D, D, numel, name, st);
  else if (k == "to_latent.bias") rc = copy_f32(h->b_tolat.p, data, numel, h->NL, name, st);
  else if (k == "from_latent.weight") rc = pack_bf16(h->w_from.p, data, E, h->c.latent_dim, 64, numel, name, st);
  else if (k == "from_latent.bias") rc = copy_f32(h->b_from.p, data, numel, E, name, st);
  else if (k == "decoder_embed.weight") rc = pack_bf16(h->w_embed.p, data, D, E, E, numel, name, st);
  else if (k == "decoder_embed.bias") rc = copy_f32(h->b_embed.p, data, numel, D, name, st);
  else if (k == "decoder_pos_embed") rc = copy_f32(h->pos.p, data, numel, (int64_t)h->L * D, name, st);
  else if (k == "decoder_norm.weight") rc = copy_f32(h->nfw.p, data, numel, D, name, st);
  else if (k == "decoder_norm.bias") rc = copy_f32(h->nfb.p, data, numel, D, name, st);
  else if (k == "decoder_pred.linear_pred.weight") rc = pack_bf16(h->w_pred.p, data, h->PP, D, D, numel, name, st);
  else if (k == "decoder_pred.linear_pred.bias") rc = copy_f32(h->b_pred.p, data, numel, h->PP, name, st);
  else if (k == "decoder_pred.conv_smoother.weight") rc = copy_f32(h->conv_w.p, data, numel, 81, name, st);
  else if (k == "decoder_pred.conv_smoother.bias") rc = copy_f32(h->conv_b.p, data, numel, 3, name, st);
  else if (bi >= 0) {
    VmaeBlockW& b = enc ? h->eblk[bi] : h->blk[bi];
    if (sub == "norm1.weight") rc = copy_f32(b.n1w.p, data, numel, D, name, st);
    else if (sub == "norm1.bias") rc = copy_f32(b.n1b.p, data, numel, D, name, st);
    else if (sub == "norm2.weight") rc = copy_f32(b.n2w.p, data, numel, D, name, st);
    else if (sub == "norm2.bias") rc = copy_f32(b.n2b.p, data, numel, D, name, st);
    else if (sub == "attn.qkv.weight") {
      LDMAE_REQUIRE(numel == (int64_t)3 * D * D, "%s: bad size", name);
      pad_heads_rows_kernel<<<3 * HPb, 128, 0, st>>>(b.w_qkv.p, nullptr, data, nullptr, nhb, hdb, D);
      LDMAE_LAUNCH_CHECK();
    } else if (sub == "attn.qkv.bias") {
      LDMAE_REQUIRE(numel == 3 * D, "%s: bad size", name);
      // scatter bias into the padded layout with a strided 2-D copy: [3*nh, hd] -> [3*nh, 64]
      LDMAE_CUDA(cudaMemsetAsync(b.b_qkv.p, 0, 3 * HPb * sizeof(float), st));
      LDMAE_CUDA(cudaMemcpy2DAsync(b.b_qkv.p, 64 * sizeof(float), data, hdb * sizeof(float), hdb * sizeof(float), 3 * nhb,
                                   cudaMemcpyDeviceToDevice, st));
    } else if (sub == "attn.proj.weight") {
      LDMAE_REQUIRE(numel == (int64_t)D * D, "%s: bad size", name);
      pad_heads_cols_kernel<<<D, 256, 0, st>>>(b.w_proj.p, data, nhb, hdb);
      LDMAE_LAUNCH_CHECK();
    } else if (sub == "attn.proj.bias") rc = copy_f32(b.b_proj.p, data, numel, D, name, st);
    else if (sub == "mlp.fc1.weight") rc = pack_bf16(b.w_fc1.p, data, h->Hm, D, D, numel, name, st);
    else if (sub == "mlp.fc1.bias") rc = copy_f32(b.b_fc1.p, data, numel, h->Hm, name, st);
    else if (sub == "mlp.fc2.weight") rc = pack_bf16(b.w_fc2.p, data, D, h->Hm, h->Hm, numel, name, st);
    else if (sub == "mlp.fc2.bias") rc = copy_f32(b.b_fc2.p, data, numel, D, name, st);
    else return set_error(LDMAE_ERR_INVALID, "unknown VMAE decoder key %s", name);
  } else {
    return set_error(LDMAE_ERR_INVALID, "unknown VMAE decoder key %s", name);
  }
  std::vector<std::string>& lst = enc_key ? h->enc_loaded : h->loaded;
  if (rc == LDMAE_OK && std::find(lst.begin(), lst.end(), k) == lst.end()) lst.push_back(k);
  return rc;
}

extern "C" int ldmae_vmae_finalize(ldmae_vmae* h, void* stream) {
  LDMAE_REQUIRE(h, "null handle");
  (void)stream;
  const int expect = 11 + 12 * h->c.decoder_depth;
  if ((int)h->loaded.size() != expect)
    return set_error(LDMAE_ERR_STATE, "VMAE decoder weights incomplete: %d of %d tensors loaded", (int)h->loaded.size(), expect);
  h->finalized = true;
  return LDMAE_OK;
}

// One pre-norm ViT block (tokenizer/models_mae.py:176-187) on the fp32 stream h->x: LN -> qkv -> attention (heads padded
// to 64) -> proj (+residual) -> LN -> fc1 + exact GELU -> fc2 (+residual).  Shared by the decoder and the encoder.
static int vmae_vit_block(ldmae_vmae* h, VmaeBlockW& b, int B, int nh, int HP, cudaStream_t st) {
  const int D = h->D, L = h->L, M = B * L;
  const float scale = 1.0f / sqrtf(static_cast<float>(D / nh));
  const unsigned ln_grid = cdiv(M, 8);
  auto resid = [&](const void* a, int lda, const void* w, int ldw, int K, const float* bias) {
    return gemm_residual(a, lda, w, ldw, GemmShape{M, D, K}, h->x.p, D, bias, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, L, st);
  };
  layernorm_bf16_kernel<<<ln_grid, 256, 0, st>>>(h->a.p, h->x.p, b.n1w.p, b.n1b.p, M, D, h->c.ln_eps);
  LDMAE_LAUNCH_CHECK();
  LDMAE_TRY((gemm_store<__nv_bfloat16, 0>(h->a.p, D, b.w_qkv.p, D, GemmShape{M, 3 * HP, D}, h->qkv.p, 3 * HP, b.b_qkv.p, st)));
  LDMAE_TRY(run_attention(h->qkv.p, 3 * HP, h->o.p, HP, B, L, nh, 0, HP, 2 * HP, scale, st));
  LDMAE_TRY(resid(h->o.p, HP, b.w_proj.p, HP, HP, b.b_proj.p));
  layernorm_bf16_kernel<<<ln_grid, 256, 0, st>>>(h->a.p, h->x.p, b.n2w.p, b.n2b.p, M, D, h->c.ln_eps);
  LDMAE_LAUNCH_CHECK();
  LDMAE_TRY((gemm_store<__nv_bfloat16, 1>(h->a.p, D, b.w_fc1.p, D, GemmShape{M, h->Hm, D}, h->hid.p, h->Hm, b.b_fc1.p, st)));
  LDMAE_TRY(resid(h->hid.p, h->Hm, b.w_fc2.p, h->Hm, h->Hm, b.b_fc2.p));
  return LDMAE_OK;
}

// out[b, c, t] = in[b*L + t, c]   ('b (h w) c -> b c h w', models_mae.py:835)
__global__ void tokens_to_nchw_kernel(float* __restrict__ out, const float* __restrict__ in, int B, int L, int C) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * L * C) return;
  const int t = i % L;
  const int c = (i / L) % C;
  const size_t b = i / (static_cast<size_t>(L) * C);
  out[i] = in[(b * L + t) * C + c];
}

// MaskedAutoencoderViT._encode (tokenizer/models_mae.py:819-836): img [B,3,H,W] fp32 -> moments [B, NL, g, g] fp32
// (NL = 2 * latent_dim with the KL bottleneck: mean || logvar, consumed by DiagonalGaussianDistribution).
extern "C" int ldmae_vmae_encode(ldmae_vmae* h, const float* img, float* moments, int32_t B, void* stream) {
  LDMAE_REQUIRE(h && h->HPe > 0, "VMAE handle was created without an encoder");
  LDMAE_REQUIRE(img && moments && B >= 1, "bad argument");
  const int expect = 7 + 12 * h->c.depth;
  if ((int)h->enc_loaded.size() != expect)
    return set_error(LDMAE_ERR_STATE, "VMAE encoder weights incomplete: %d of %d tensors loaded", (int)h->enc_loaded.size(), expect);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B > h->maxB) { LDMAE_CUDA(cudaStreamSynchronize(st)); LDMAE_TRY(vmae_alloc_ws(h, B)); }
  ProfScope ps(8, st);
  const int D = h->D, L = h->L, M = B * L;
  // timm PatchEmbed: Conv2d(3, E, k = stride = p) == GEMM over the (c, pi, qi) patch rows
  patchify_bf16_kernel<<<cdiv(static_cast<size_t>(M) * h->PP, 256), 256, 0, st>>>(h->etok.p, img, B, 3, h->c.img_size, h->c.patch_size);
  LDMAE_LAUNCH_CHECK();
  broadcast_rows_kernel<<<cdiv(static_cast<size_t>(M) * D, 256), 256, 0, st>>>(h->x.p, h->epos.p, M, L, D);
  LDMAE_LAUNCH_CHECK();
  LDMAE_TRY(gemm_residual(h->etok.p, h->PP, h->w_patch.p, h->PP, GemmShape{M, D, h->PP}, h->x.p, D, h->b_patch.p, nullptr, 0, nullptr, 0,
                          nullptr, nullptr, 0, L, st));
  for (int i = 0; i < h->c.depth; ++i) LDMAE_TRY(vmae_vit_block(h, h->eblk[i], B, h->c.num_heads, h->HPe, st));
  layernorm_bf16_kernel<<<cdiv(M, 8), 256, 0, st>>>(h->a.p, h->x.p, h->enw.p, h->enb.p, M, D, h->c.ln_eps);
  LDMAE_LAUNCH_CHECK();
  LDMAE_TRY((gemm_store<float, 0>(h->a.p, D, h->w_tolat.p, D, GemmShape{M, h->NL, D}, h->pred.p, h->NL, h->b_tolat.p, st)));
  tokens_to_nchw_kernel<<<cdiv(static_cast<size_t>(M) * h->NL, 256), 256, 0, st>>>(moments, h->pred.p, B, L, h->NL);
  LDMAE_LAUNCH_CHECK();
  return LDMAE_OK;
}

extern "C" int ldmae_vmae_decode(ldmae_vmae* h, const float* z, const float* mean, const float* stdv, float multiplier,
                                 float* img_f32, uint8_t* img_u8, int32_t B, void* stream) {
  LDMAE_REQUIRE(h && h->finalized, "VMAE handle not finalized");
  LDMAE_REQUIRE(z && (img_f32 || img_u8) && B >= 1, "bad argument");
  LDMAE_REQUIRE((mean == nullptr) == (stdv == nullptr), "mean/std must be given together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B > h->maxB) { LDMAE_CUDA(cudaStreamSynchronize(st)); LDMAE_TRY(vmae_alloc_ws(h, B)); }
  ProfScope ps(8, st);
  const int D = h->D, L = h->L, E = h->c.embed_dim;
  const int M = B * L;
  const int nh = h->c.decoder_num_heads, hd = D / nh;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  latent_to_tokens_kernel<<<cdiv(static_cast<size_t>(M) * 64, 256), 256, 0, st>>>(h->tok.p, z, mean, stdv,
                                                                                 multiplier != 0.f ? 1.f / multiplier : 1.f,
                                                                                 B, h->c.latent_dim, L, 64);
  LDMAE_LAUNCH_CHECK();
  LDMAE_TRY((gemm_store<__nv_bfloat16, 0>(h->tok.p, 64, h->w_from.p, 64, GemmShape{M, E, 64}, h->t1.p, E, h->b_from.p, st)));
  broadcast_rows_kernel<<<cdiv(static_cast<size_t>(M) * D, 256), 256, 0, st>>>(h->x.p, h->pos.p, M, L, D);
  LDMAE_LAUNCH_CHECK();
  auto resid = [&](const void* a, int lda, const void* w, int ldw, int K, const float* bias) {
    return gemm_residual(a, lda, w, ldw, GemmShape{M, D, K}, h->x.p, D, bias, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, L, st);
  };
  LDMAE_TRY(resid(h->t1.p, E, h->w_embed.p, E, E, h->b_embed.p));
  const unsigned ln_grid = cdiv(M, 8);
  for (int i = 0; i < h->c.decoder_depth; ++i) LDMAE_TRY(vmae_vit_block(h, h->blk[i], B, nh, h->HP, st));
  layernorm_bf16_kernel<<<ln_grid, 256, 0, st>>>(h->a.p, h->x.p, h->nfw.p, h->nfb.p, M, D, h->c.ln_eps);
  LDMAE_LAUNCH_CHECK();
  LDMAE_TRY((gemm_store<float, 0>(h->a.p, D, h->w_pred.p, D, GemmShape{M, h->PP, D}, h->pred.p, h->PP, h->b_pred.p, st)));
  const size_t npix = static_cast<size_t>(B) * h->c.img_size * h->c.img_size;
  vmae_pixel_tail_kernel<<<cdiv(npix, 256), 256, 0, st>>>(img_f32, img_u8, h->pred.p, h->conv_w.p, h->conv_b.p, B, h->G,
                                                          h->c.patch_size);
  LDMAE_LAUNCH_CHECK();
  return LDMAE_OK;
}

// ------------------------------------------------------------------------------------------- building blocks
extern "C" int ldmae_gemm_bias(const void* a, const void* w, const float* bias, void* out, int32_t out_is_bf16,
                               int32_t M, int32_t N, int32_t K, int32_t act, int32_t cta_group, int32_t block_n,
                               void* stream) {
  LDMAE_TRY(require_sm100());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LDMAE_REQUIRE(K % 8 == 0, "K must be a multiple of 8");
  LDMAE_REQUIRE((cta_group == 1 && block_n == 128) || (cta_group == 2 && block_n == 256) || (cta_group == 1 && block_n == 256),
                "supported tile configs: (cg1,bn128) (cg1,bn256) (cg2,bn256)");
  ++g_launch_count;
  const GemmShape g{M, N, K};
#define LDMAE_DISPATCH(EPI, PARAMS)                                                             \
  do {                                                                                          \
    if (cta_group == 2) return launch_gemm<256, 2, EPI>(a, K, w, K, g, PARAMS, st);             \
    if (block_n == 256) return launch_gemm<256, 1, EPI>(a, K, w, K, g, PARAMS, st);             \
    return launch_gemm<128, 1, EPI>(a, K, w, K, g, PARAMS, st);                                 \
  } while (0)
  using EpiB0 = EpiStore<__nv_bfloat16, 0>;
  using EpiB1 = EpiStore<__nv_bfloat16, 1>;
  using EpiF0 = EpiStore<float, 0>;
  if (out_is_bf16) {
    if (act == 1) {
      EpiB1::Params p;
      LDMAE_TRY(make_tmap_out_bf16(&p.omap, out, M, N, N));
      p.bias = bias;
      LDMAE_DISPATCH(EpiB1, p);
    }
    EpiB0::Params p;
    LDMAE_TRY(make_tmap_out_bf16(&p.omap, out, M, N, N));
    p.bias = bias;
    LDMAE_DISPATCH(EpiB0, p);
  }
  LDMAE_REQUIRE(act == 0, "fp32 output supports act 0 only");
  EpiF0::Params p;
  LDMAE_TRY(make_tmap_out_f32(&p.omap, out, M, N, N));
  p.bias = bias;
  LDMAE_DISPATCH(EpiF0, p);
#undef LDMAE_DISPATCH
}

extern "C" int ldmae_gemm_residual(const void* a, const void* w, const float* bias, const float* gate, const float* gnext,
                                   float* x, void* anext, float* ssq, int32_t M, int32_t N, int32_t K, int32_t rows_per_sample,
                                   void* stream) {
  LDMAE_TRY(require_sm100());
  LDMAE_REQUIRE(K % 8 == 0 && N % 64 == 0, "K must be a multiple of 8 and N of 64");
  LDMAE_REQUIRE((anext == nullptr) == (gnext == nullptr), "anext and gnext go together");
  return gemm_residual(a, K, w, K, GemmShape{M, N, K}, x, N, bias, gate, N, gnext, N, static_cast<__nv_bfloat16*>(anext), ssq,
                       (N + 127) / 128, rows_per_sample, static_cast<cudaStream_t>(stream));
}

extern "C" int ldmae_attention(const void* qkv, void* out, int32_t B, int32_t T, int32_t H, float scale, void* stream) {
  LDMAE_TRY(require_sm100());
  return run_attention(qkv, 3 * H * 64, out, H * 64, B, T, H, 0, H * 64, 2 * H * 64, scale, static_cast<cudaStream_t>(stream));
}
extern "C" int ldmae_attention_wide_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, float* delta_ws,
                                        void* dqkv, int32_t B, int32_t T, int32_t H, int32_t hd, float scale, void* stream) {
  LDMAE_TRY(require_sm100());
  return run_attention_bwd(qkv, 3 * H * 128, out, dout, H * hd, lse2, delta_ws, dqkv, B, T, H, 0, H * 128, 2 * H * 128, scale,
                           static_cast<cudaStream_t>(stream), hd, 128);
}
extern "C" int ldmae_attention_wide_lse(const void* qkv, void* out, float* lse2, int32_t B, int32_t T, int32_t H, int32_t hd,
                                        float scale, void* stream) {
  LDMAE_TRY(require_sm100());
  return run_attention_hd128(qkv, 3 * H * 128, out, H * hd, B, T, H, hd, 0, H * 128, 2 * H * 128, scale,
                             static_cast<cudaStream_t>(stream), lse2);
}
extern "C" int ldmae_attention_wide(const void* qkv, void* out, int32_t B, int32_t T, int32_t H, int32_t hd, float scale,
                                    void* stream) {
  LDMAE_TRY(require_sm100());
  return run_attention_hd128(qkv, 3 * H * 128, out, H * hd, B, T, H, hd, 0, H * 128, 2 * H * 128, scale,
                             static_cast<cudaStream_t>(stream));
}
extern "C" int ldmae_attention_bounded(const void* qkv, void* out, float* lse2, int32_t B, int32_t T, int32_t H, float scale,
                                       float m0_log2, void* stream) {
  LDMAE_TRY(require_sm100());
  LDMAE_REQUIRE(m0_log2 > 0.f && m0_log2 <= 48.f, "score bound must be in (0, 48] (log2 units)");
  return run_attention(qkv, 3 * H * 64, out, H * 64, B, T, H, 0, H * 64, 2 * H * 64, scale, static_cast<cudaStream_t>(stream), lse2,
                       m0_log2);
}
extern "C" int ldmae_attention_prescaled(const void* qkv, void* out, float* lse2, int32_t B, int32_t T, int32_t H, float m0_log2,
                                         void* stream) {
  LDMAE_TRY(require_sm100());
  LDMAE_REQUIRE(m0_log2 > 0.f && m0_log2 <= 48.f, "score bound must be in (0, 48] (log2 units)");
  return run_attention(qkv, 3 * H * 64, out, H * 64, B, T, H, 0, H * 64, 2 * H * 64, 1.f, static_cast<cudaStream_t>(stream), lse2,
                       m0_log2, true);
}
extern "C" int ldmae_attention_lse(const void* qkv, void* out, float* lse2, int32_t B, int32_t T, int32_t H, float scale,
                                   void* stream) {
  LDMAE_TRY(require_sm100());
  return run_attention(qkv, 3 * H * 64, out, H * 64, B, T, H, 0, H * 64, 2 * H * 64, scale, static_cast<cudaStream_t>(stream), lse2);
}
extern "C" int ldmae_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, float* delta_ws,
                                   void* dqkv, int32_t B, int32_t T, int32_t H, float scale, void* stream) {
  LDMAE_TRY(require_sm100());
  return run_attention_bwd(qkv, 3 * H * 64, out, dout, H * 64, lse2, delta_ws, dqkv, B, T, H, 0, H * 64, 2 * H * 64, scale,
                           static_cast<cudaStream_t>(stream));
}
extern "C" int ldmae_gemm_wgrad(const void* p_bf16, const void* q_bf16, float* c, int32_t N1, int32_t N2, int32_t M, float alpha,
                                void* stream) {
  LDMAE_TRY(require_sm100());
  return gemm_wgrad(p_bf16, N1, q_bf16, N2, c, N2, N1, N2, M, alpha, static_cast<cudaStream_t>(stream));
}

__global__ void f32_to_bf16_kernel(__nv_bfloat16* out, const float* in, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16(in[i]);
}
extern "C" int ldmae_f32_to_bf16(const float* in, void* out, int64_t n, void* stream) {
  f32_to_bf16_kernel<<<cdiv(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<__nv_bfloat16*>(out), in, n);
  LDMAE_LAUNCH_CHECK();
  return LDMAE_OK;
}
