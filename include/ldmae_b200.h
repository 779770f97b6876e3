/* ldmae_b200 -- C ABI of the B200-native LDMAE hot path (libldmae_b200.so).
 *
 * The reference (isno0907/ldmae) is pure Python/PyTorch and has no FFI; its boundary for this path is
 * the Python module API that LDMAE/inference.py and LDMAE/train_accum.py import.  Each entry point
 * below names the reference interface it sits under; ldmae_b200/ (Python) keeps those module names
 * and signatures and calls this library through ctypes.
 *
 * Conventions: every function returns 0 on success and a negative code on failure; the message is
 * available through ldmae_last_error() (thread-local).  Nothing throws across the boundary.  All
 * tensor pointers are DEVICE pointers to contiguous memory in the reference's own layouts (NCHW fp32
 * latents, int64 labels) unless a parameter says "host".  `stream` is a cudaStream_t passed as
 * void* (PyTorch's current stream).  Handles are not thread-safe; one process per GPU.
 */
#ifndef LDMAE_B200_H
#define LDMAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDMAE_OK 0
#define LDMAE_ERR_INVALID (-1)     /* bad argument / unsupported configuration */
#define LDMAE_ERR_CUDA (-2)        /* a CUDA call or kernel failed */
#define LDMAE_ERR_STATE (-3)       /* weights missing, workspace too small, ... */

const char* ldmae_last_error(void);
/* Library / device introspection: writes the SM count and compute capability (e.g. 100). */
int ldmae_device_info(int* sm_count, int* cc);
int ldmae_version(void);
/* Number of kernel launches issued by the library so far in this process. */
long long ldmae_launch_count(void);
/* Per-kernel-class device timing with CUDA events on the launching stream (bench.py roofline).
 * Classes: 0 qkv GEMM, 1 attention, 2 proj GEMM, 3 w12/SwiGLU GEMM, 4 w3 GEMM, 5 adaLN + shift-vector GEMMs,
 * 6 final-layer GEMM, 7 conditioning / patch embed / ODE update, 8 VMAE decode. */
#define LDMAE_PROF_CLASSES 13   /* training adds: 9 data-gradient GEMMs, 10 weight-gradient GEMMs, 11 attention backward, 12 HBM-bound backward kernels */
int ldmae_profile_begin(void);
int ldmae_profile_end(double* ms_per_class, long long* scopes_per_class, int32_t nclasses);

/* ------------------------------------------------------------------------------------------------
 * LightningDiT denoiser -- reference LDMAE/models/lightningdit.py:275-442 (class LightningDiT)
 * ---------------------------------------------------------------------------------------------- */
typedef struct ldmae_dit_config {
  int32_t depth, hidden_size, num_heads, patch_size;   /* registry lightningdit.py:498-531 */
  int32_t input_size, in_channels;                     /* latent side and channels (inference.py:336-347) */
  int32_t num_embeddings;                              /* rows of y_embedder.embedding_table (num_classes [+1]) */
  int32_t mlp_hidden;                                  /* SwiGLU hidden = int(2/3 * 4 * hidden) (lightningdit.py:217) */
  int32_t learn_sigma, use_qknorm, use_swiglu, use_rope, use_rmsnorm, wo_shift;
  int32_t max_batch;                                   /* largest forward batch (2n with CFG) the workspace is sized for */
} ldmae_dit_config;

typedef struct ldmae_dit ldmae_dit;

int ldmae_dit_create(const ldmae_dit_config* cfg, ldmae_dit** out);
void ldmae_dit_destroy(ldmae_dit* h);

/* nn.Module.load_state_dict, one tensor at a time: `name` is the reference state_dict key
 * (SURVEY.md section 8b), `data` a device pointer to contiguous fp32 with `numel` elements.
 * The library copies / packs (bf16, SwiGLU row interleave, adaLN concatenation) into its own buffers. */
int ldmae_dit_load_tensor(ldmae_dit* h, const char* name, const float* data, int64_t numel, void* stream);
/* Call once after all tensors were loaded (checks completeness). */
int ldmae_dit_finalize(ldmae_dit* h, void* stream);

/* LightningDiT.forward(x, t, y) in eval mode (lightningdit.py:391-418).
 *   x [B, C, S, S] fp32, t [B] fp32 (or NULL with t_scalar broadcast), y [B] int64 -> out [B, C, S, S] fp32.
 *   src_mod: sample b reads x[b % src_mod] (pass B for a plain forward; n for forward_with_cfg's cat[half,half]). */
int ldmae_dit_forward(ldmae_dit* h, const float* x, const float* t, float t_scalar, const int64_t* y, float* out,
                      int32_t B, int32_t src_mod, void* stream);

/* Training (reference transport/transport.py:169-215 + train_accum.py:215-246): the forward that keeps what the
 * backward needs, the backward (gradients of every trainable parameter from dL/d(output) [B,C,S,S]), and gradient
 * read-out in the reference's state_dict layout (`name` = state_dict key; pos_embed and the RoPE buffers have none).
 * x [B,C,S,S], t [B], y [B] (after the caller's label dropout, lightningdit.py:157-160).  Gradients are valid until the
 * next ldmae_dit_backward. */
int ldmae_dit_train_forward(ldmae_dit* h, const float* x, const float* t, const int64_t* y, float* out, int32_t B, void* stream);
int ldmae_dit_backward(ldmae_dit* h, const float* dout, int32_t B, void* stream);
/* Only ONE training forward may be pending per handle: the kept activations and the shared conditioning workspace are
 * overwritten by the next forward of any kind.  ldmae_dit_generation() is bumped by every forward / workspace
 * re-allocation; ldmae_dit_backward fails with LDMAE_ERR_STATE when its forward is no longer the latest (the reference's
 * autograd graph would keep both alive, models/lightningdit.py:391-418 under train_accum.py:215-230). */
long long ldmae_dit_generation(ldmae_dit* h);
/* Batched forms of ldmae_dit_load_tensor / ldmae_dit_grad_read / ldmae_dit_grad_accumulate: n (name, pointer, numel) triples per
 * call -- the fused trainer moves ~150 tensors each way per optimizer step. */
int ldmae_dit_load_tensors(ldmae_dit* h, const char* const* names, const float* const* data, const int64_t* numels, int32_t n,
                           void* stream);
int ldmae_dit_grad_read_many(ldmae_dit* h, const char* const* names, float* const* dsts, const int64_t* numels, int32_t n,
                             int32_t accumulate, void* stream);
/* Class labels (reference nn.Embedding, lightningdit.py:146-169) outside the table are clamped and flagged on the device;
 * the flag is reported by the NEXT call on the handle, or right away by this call (which waits for `stream`). */
int ldmae_dit_check_labels(ldmae_dit* h, void* stream);
int ldmae_dit_grad_read(ldmae_dit* h, const char* name, float* dst, int64_t numel, void* stream);
/* Same, but dst += gradient: micro-batch accumulation between optimizer steps (train_accum.py:220-234). */
int ldmae_dit_grad_accumulate(ldmae_dit* h, const char* name, float* dst, int64_t numel, void* stream);
/* Fused torch.optim.AdamW step + EMA update (train_accum.py:121,240-246,337-347) on flat fp32 device buffers of n
 * elements (16-byte aligned); ema may be NULL; grad is multiplied by grad_scale first (1/world_size after an all-reduce);
 * step counts from 1. */
int ldmae_adamw_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t n, float lr,
                         float beta1, float beta2, float eps, float weight_decay, int32_t step, float ema_decay, float grad_scale,
                         void* stream);

/* Trainer input pipeline and loss, one pass each (reference datasets/img_latent_dataset.py:76-94: flip select, posterior
 * sample, per-channel normalise, multiplier; transport/transport.py:136-166 + path.py:114-136: xt = t*x1 + (1-t)*x0,
 * ut = x1 - x0; transport.py:195 + train_accum.py:220-223: loss[b] = mean_flat((out-ut)^2), dout = d(mean_b loss * loss_scale)).
 * moments / moments_flip [B, 2C, HW] (mean || logvar), flip [B] bytes (non-zero = take the flipped row), eps_post [B, C, HW]
 * (NULL = posterior mode), mean / std [C] (NULL = no normalisation); OR x1_in [B, C, HW] ready latents (moments NULL).
 * x0 [B, C, HW] noise, t [B]; outputs xt, ut and optionally x1_out.  HW % 4 == 0. */
int ldmae_flow_prepare(const float* moments, const float* moments_flip, const uint8_t* flip, const float* eps_post,
                       const float* mean, const float* stdv, float multiplier, const float* x1_in, const float* x0, const float* t,
                       float* x1_out, float* xt, float* ut, int32_t B, int32_t C, int32_t HW, void* stream);
int ldmae_flow_loss(const float* out, const float* ut, float* loss, float* dout /* or NULL */, float loss_scale, int32_t B, int32_t n,
                    void* stream);

/* Test hooks: make ldmae_dit_forward return after `stages` launch groups (-1 = run everything; 1 conditioning,
 * 2 adaLN, 3 shift vectors, 4 patch embed, then 5 per block: qkv, attention, proj, w12, w3), and copy a named
 * workspace buffer ("xres", "abuf", "qkv", "obuf", "hbuf", "ssq", "mods", ...) to `dst` (device). */
int ldmae_dit_debug_stop(ldmae_dit* h, int32_t stages);
int ldmae_dit_debug_poison(ldmae_dit* h, int32_t byte, void* stream);   /* memset every workspace buffer */
int ldmae_dit_debug_read(ldmae_dit* h, const char* name, void* dst, int64_t nbytes, void* stream);

/* LightningDiT.forward_with_cfg (lightningdit.py:420-442): x [2n,...], t [2n], y [2n] (last n = null class);
 * guidance on channels [:3]; use_guidance = !(cfg_interval && t[0] < cfg_interval_start), decided by the
 * caller on the host (the reference syncs the device for it). */
int ldmae_dit_forward_with_cfg(ldmae_dit* h, const float* x, const float* t, float t_scalar, const int64_t* y,
                               float* out, int32_t n, float cfg_scale, int32_t use_guidance, void* stream);

/* transport Sampler.sample_ode(...)(x, model_fn, **kw) for path Linear / prediction velocity
 * (transport/transport.py:398-443, integrators.py:77-126) with the fixed-grid solver of torchdiffeq.
 *   x [Btot, C, S, S] fp32 in/out (Btot = 2n when cfg_scale > 1 and model_fn = forward_with_cfg, else n)
 *   tgrid: HOST array of npts fp32 time points (npts - 1 model evaluations)
 *   method: 0 euler, 1 heun2;  use_cfg: model_fn is forward_with_cfg;  cfg_interval_start < 0: no interval
 *   traj (optional, device, [npts, Btot, C, S, S]): every grid state like torchdiffeq returns; NULL keeps only the last. */
/*   flags: LDMAE_ODE_COND_ONLY_WHEN_UNGUIDED -- on steps with t < cfg_interval_start the guided velocity of the conditional
 *   half is its own prediction, so only x[:n] is evaluated and advanced; x[n:] is then NOT the reference's second half
 *   (inference.py:289 discards it: samples.chunk(2)[0]).  Results for x[:n] are identical.  Off by default. */
#define LDMAE_ODE_COND_ONLY_WHEN_UNGUIDED 1
int ldmae_sample_ode(ldmae_dit* h, float* x, const int64_t* y, int32_t n, int32_t use_cfg, float cfg_scale,
                     float cfg_interval_start, const float* tgrid, int32_t npts, int32_t method, float* traj,
                     int32_t flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * VMAE f8d16 ViT decoder -- reference LDMAE/tokenizer/models_mae.py:865-887 (decode), :963-973
 * ---------------------------------------------------------------------------------------------- */
typedef struct ldmae_vmae_config {
  int32_t img_size, patch_size, latent_dim;            /* 256, 8, 16 */
  int32_t embed_dim;                                   /* encoder width = input of decoder_embed (192) */
  int32_t decoder_embed_dim, decoder_depth, decoder_num_heads;  /* 192, 12, 12 */
  int32_t mlp_hidden;                                  /* 4 * decoder_embed_dim */
  float ln_eps;                                        /* 1e-6 (models_mae.py:996) */
  int32_t max_batch;
  /* encoder side (0 = decoder only): ViT depth / heads at width embed_dim, outputs of to_latent (2*latent_dim with KL) */
  int32_t depth, num_heads, to_latent_dim;
} ldmae_vmae_config;

typedef struct ldmae_vmae ldmae_vmae;

int ldmae_vmae_create(const ldmae_vmae_config* cfg, ldmae_vmae** out);
void ldmae_vmae_destroy(ldmae_vmae* h);
int ldmae_vmae_load_tensor(ldmae_vmae* h, const char* name, const float* data, int64_t numel, void* stream);
int ldmae_vmae_finalize(ldmae_vmae* h, void* stream);

/* decode(z) (+ the latent de-normalisation of inference.py:291 when mean/std are given:
 * z*std/multiplier + mean, per channel [C]).  Writes whichever outputs are non-NULL:
 *   img_f32 [B,3,H,W] fp32 (decode(...)[0]) and/or img_u8 [B,H,W,3] uint8 (decode_to_images). */
int ldmae_vmae_decode(ldmae_vmae* h, const float* z, const float* mean, const float* stdv, float multiplier,
                      float* img_f32, uint8_t* img_u8, int32_t B, void* stream);

/* MaskedAutoencoderViT._encode (tokenizer/models_mae.py:819-836; called by extract_features.py:150-152 and by encode):
 * img [B,3,H,W] fp32 -> moments [B, to_latent_dim, H/p, W/p] fp32 (mean || logvar of the KL posterior).  Needs the encoder
 * keys (patch_embed.*, pos_embed, blocks.*, norm.*, to_latent.*) loaded through ldmae_vmae_load_tensor. */
int ldmae_vmae_encode(ldmae_vmae* h, const float* img, float* moments, int32_t B, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Building blocks, exported for the parity tests and micro-benchmarks (device pointers).
 * ---------------------------------------------------------------------------------------------- */
/* out[M,N] = act(a[M,K] . w[N,K]^T + bias); a,w bf16 (raw uint16 storage), out fp32 or bf16.
 * act: 0 none, 1 gelu(erf).  cta_group: 1 or 2.  block_n: 128 or 256. */
int ldmae_gemm_bias(const void* a_bf16, const void* w_bf16, const float* bias, void* out, int32_t out_is_bf16,
                    int32_t M, int32_t N, int32_t K, int32_t act, int32_t cta_group, int32_t block_n, void* stream);
/* Fused residual GEMM (the LightningDiT block's out-projection / w3 epilogue, models/lightningdit.py:248-249):
 *   x[M,N] (fp32, in place) += gate[b,:] * (a . w^T + bias),  b = row / rows_per_sample, gate [M/rows_per_sample, N] or NULL;
 *   optionally anext[M,N] (bf16) = x_new * gnext[b,:] and ssq[M, ceil(N/128)] = partial sums of x_new^2 (slot = 128-column
 *   group of the producing tile; unused slots are 0).  N must be a multiple of 4. */
int ldmae_gemm_residual(const void* a_bf16, const void* w_bf16, const float* bias, const float* gate, const float* gnext,
                        float* x, void* anext_bf16, float* ssq, int32_t M, int32_t N, int32_t K, int32_t rows_per_sample,
                        void* stream);
/* softmax(q k^T * scale) v over qkv [B*T, 3*H*64] bf16 (columns q|k|v, head-major) -> out [B*T, H*64] bf16. */
int ldmae_attention(const void* qkv_bf16, void* out_bf16, int32_t B, int32_t T, int32_t H, float scale, void* stream);
/* Training forward of the same attention: additionally writes lse2 [B, H, T (+64 floats of padding at the end)] =
 * log2(sum_j exp(s_ij * scale)), the row statistic the backward needs. */
int ldmae_attention_lse(const void* qkv_bf16, void* out_bf16, float* lse2, int32_t B, int32_t T, int32_t H, float scale,
                        void* stream);
/* Heads wider than 64 (LightningDiT-XL: head_dim 72): qkv [B*T, 3*H*128] bf16 with every head in a 128-column slot
 * (head_dim `hd` real columns, zeros behind), out [B*T, H*hd] bf16 dense.  hd: multiple of 8, <= 128. */
int ldmae_attention_wide(const void* qkv_bf16, void* out_bf16, int32_t B, int32_t T, int32_t H, int32_t hd, float scale,
                         void* stream);
/* Training forms of the wide-head attention: forward that also writes lse2 [B,H,T (+64)], and its gradient
 * (dqkv in the 128-column slot layout, padding columns zero; delta_ws: 2 * (B*H*T + 64) floats; T % 4 == 0). */
int ldmae_attention_wide_lse(const void* qkv_bf16, void* out_bf16, float* lse2, int32_t B, int32_t T, int32_t H, int32_t hd,
                             float scale, void* stream);
int ldmae_attention_wide_bwd(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2, float* delta_ws,
                             void* dqkv_bf16, int32_t B, int32_t T, int32_t H, int32_t hd, float scale, void* stream);
/* The same attention when the scores are known to be bounded, |q.k * scale| * log2(e) <= m0_log2 (qk-normed heads:
 * 8 * log2(e) * max|q_norm.w| * max|k_norm.w|): the bound replaces the running row maximum (no max pass, no rescale).
 * lse2 may be NULL. */
int ldmae_attention_bounded(const void* qkv_bf16, void* out_bf16, float* lse2, int32_t B, int32_t T, int32_t H, float scale,
                            float m0_log2, void* stream);
/* The bounded-score attention for q that already carries scale * log2(e) (the inference forward folds both into
 * q_norm.weight in the QKV epilogue): probabilities are 2^(q.k) with no per-score multiply-add; |q.k| <= m0_log2 <= 48.
 * lse2 (log2 units) may be NULL. */
int ldmae_attention_prescaled(const void* qkv_bf16, void* out_bf16, float* lse2, int32_t B, int32_t T, int32_t H,
                              float m0_log2, void* stream);
/* Gradient of ldmae_attention (the backward of F.scaled_dot_product_attention at models/lightningdit.py:77):
 * dqkv [B*T, 3*H*64] bf16 (dq | dk | dv, same layout as qkv) from dout [B*T, H*64] bf16, the forward's out and lse2;
 * delta_ws: workspace of 2 * (B*H*T + 64) floats.  T must be a multiple of 4. */
int ldmae_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2, float* delta_ws,
                        void* dqkv_bf16, int32_t B, int32_t T, int32_t H, float scale, void* stream);
/* Weight-gradient GEMM (the dW of every nn.Linear on the path): c[N1,N2] (fp32) += alpha * sum_m p[m,N1] * q[m,N2];
 * p = gradient of the Linear's output [M,N1] bf16, q = the Linear's input [M,N2] bf16; N1, N2 multiples of 8. */
int ldmae_gemm_wgrad(const void* p_bf16, const void* q_bf16, float* c, int32_t N1, int32_t N2, int32_t M, float alpha,
                     void* stream);
/* Debug builds only (-DLDMAE_ATTN_TRACE): device buffer [2][64][8] int64 receiving clock64 phase stamps of CTA 0. */
int ldmae_attention_trace(long long* dev_buf);
/* Debug builds only (-DLDMAE_GEMM_TRACE): device buffer [256][8] int64, clock64 stamps of the residual epilogue (CTA 0). */
int ldmae_gemm_trace(long long* dev_buf);
/* float <-> bf16 conversion helpers for tests (device, n elements) */
int ldmae_f32_to_bf16(const float* in, void* out_bf16, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LDMAE_B200_H */
