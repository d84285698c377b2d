/* dgq_b200 -- C ABI of the B200 (sm_100a) kernels behind DGQ's quantized UNet forward path.
 *
 * The reference (ugonfor/DGQ) has no FFI: its "plugin" boundary is the Python module API
 * (quant.quant_model.QuantModel / quant.quant_layer.QuantLayer / quant.quant_block.*) and every
 * kernel below replaces a sequence of eager ATen ops inside those modules.  Each entry point cites
 * the reference code it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing allocates,
 *     nothing synchronises;
 *   - return value: 0 on success, a positive cudaError_t, or DGQ_ERR_INVALID_VALUE (-1) when an
 *     argument violates a documented constraint.  Nothing throws across the boundary;
 *   - activations are token-major / NHWC: a tensor (B, H, W, C) or (B, T, C) is a row-major
 *     matrix [M = B*H*W, C].  Tensor-core OPERANDS (quantised activations, weights, Q/K/V, the
 *     softmax map) are fp16; the activations BETWEEN kernels (GEMM results, residual stream) are
 *     fp32 by default, because every one of them feeds a quantizer and fp16 storage alone moves
 *     ~1 % of the values across a rounding boundary (DESIGN.md "why fp32 between kernels");
 *     `*_is_f32` flags select fp16 storage instead.  Scales are fp32.
 */
#ifndef DGQ_B200_H_
#define DGQ_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGQ_ERR_INVALID_VALUE (-1)

/* how a (delta, zero_point) pair is indexed -- the three shapes of the reference checkpoint
 * (SURVEY.md 8a'): () scalar, (1,1,X) along the reduction axis, (1,X,1) along rows.          */
#define DGQ_Q_NONE 0    /* no activation quantization (use_aq off / disable_aq)                */
#define DGQ_Q_SCALAR 1  /* delta[0], zp[0]                                                      */
#define DGQ_Q_KWISE 2   /* delta[k], k = position along K (linear: channel; conv: tap*C + c)    */
#define DGQ_Q_ROWWISE 3 /* delta[m % period], m = output row (token / output pixel)             */

typedef struct {
  const float* delta; /* device, 1 or many entries depending on mode */
  const float* zp;
  const float* inv_delta; /* optional: 1/delta, correctly rounded, same indexing as delta; NULL: the
                             kernels compute it (the quantizer multiplies by it on its fast path and
                             falls back to the IEEE division near rounding ties, so codes stay
                             bit-identical to round(x/delta))                                       */
  int mode;     /* DGQ_Q_*            */
  int period;   /* DGQ_Q_ROWWISE only */
  float qmax;   /* 2^bits - 1         */
  int emit_int; /* producers only: 1 = write the integer (code - zp), exact in fp16, instead of
                   delta * (code - zp); the consumer GEMM applies delta per row (row_scale).
                   Valid for SCALAR / ROWWISE, |code - zp| <= 2048.
                   2 = write the u8 CODE itself, one byte per element (operand of dgq_gemm_i8;
                   the output row stride is then in bytes).  SCALAR / ROWWISE only.           */
} dgq_quant_t;

int dgq_version(void);

/* ---- UniformAffineQuantizer.forward, stand-alone (quant/quant_layer.py:295-299) -------------
 * x viewed as [outer, period, inner]; delta/zp index = (i / inner) % period (period = 1: scalar).
 * out_dq (fp32 de-quantised) and out_codes (u8 integer codes) are each optional.               */
int dgq_fake_quant_f32(const float* x, int64_t n, const float* delta, const float* zp, int period,
                       int64_t inner, float qmax, float* out_dq, uint8_t* out_codes, void* stream);

/* ---- T2ILogQuantizer.forward (quant/quant_layer_text.py:96-105) -----------------------------
 * code = clamp(rint(-log2(x / delta)), 0, qmax); out = 2^-code * delta.  `delta` is a device
 * scalar; for real_time=True fill it first with dgq_max_f32.                                   */
int dgq_t2i_log_quant_f32(const float* x, int64_t n, const float* delta, float qmax, float* out_dq,
                          uint8_t* out_codes, void* stream);
/* global max of x into *out (x.max(), quant_layer_text.py:97); scratch: >= 1024 floats         */
int dgq_max_f32(const float* x, int64_t n, float* out, float* scratch, void* stream);

/* ---- offline weight pack (replaces the per-forward wqtizer(self.w), quant/quant_layer.py:642-643;
 *      AdaRound hard rounding quant/adaptive_rounding.py:51-70 when alpha != NULL) --------------
 * w: fp32 [n, k_in] (conv: [Co, Ci, kh, kw] flattened, taps = kh*kw, ci = Ci).  Emits
 *   codes   : u8 [n, k_out] integer codes in GEMM K order (tap-major: tap*ci_pad + c), optional;
 *   packed4 : u8 [n, k_out/2] two 4-bit codes per byte (low nibble first), optional (bits == 4);
 *   operand : fp16 [n_pad, k_out] holding (code - zp) exactly -- the tcgen05 B operand;
 *   k_out = taps * ci_pad; rows n..n_pad-1 and channels ci..ci_pad-1 are zero.
 * use_wq == 0: operand = fp16(w) (conv_in / conv_out stay FP, quant/quant_model.py:118-124).   */
int dgq_pack_weight(const float* w, const float* delta, const float* zp, const float* alpha, int n,
                    int ci, int taps, int ci_pad, int n_pad, float qmax, int use_wq, uint8_t* codes,
                    uint8_t* packed4, void* operand, void* stream);

/* ---- compiled-checkpoint load (replaces the fp32 pickle + per-load re-quantisation of
 *      quant/calibration.py:208-251): codes [n_pad, taps*ci_pad] (bits == 8) or two codes per byte, low
 *      nibble first (bits == 4) as written by dgq_pack_weight -> operand fp16 = code - zp[row]
 *      (0 in the padded rows / channels).                                                            */
int dgq_unpack_weight(const uint8_t* codes, int bits, const float* zp, int n, int ci, int taps, int ci_pad,
                      int n_pad, void* operand, void* stream);

/* ---- activation producer: fused [concat] -> [nearest x2] -> [GroupNorm] -> [SiLU] -> im2col ->
 *      UniformAffineQuantizer, emitting the fp16 A operand of the following GEMM ----------------
 * Replaces F.unfold + aqtizer (quant/quant_layer.py:630-641), GroupNorm+SiLU
 * (quant/quant_block.py:100-110), torch.cat (diffusers_rewrite/sd.py:425,459) and nn.Upsample
 * (sd.py:322-329).                                                                               */
typedef struct {
  const void* src0; /* [batch, hs, ws, c0]; fp16, or fp32 when src_is_f32                  */
  const void* src1; /* optional second source, channels c0..c0+c1-1 (skip concat)           */
  int c0, c1;       /* multiples of 8                                                       */
  int src_is_f32;
  int batch, h, w;  /* conv input size (after the optional upsample)                        */
  int upsample;     /* 1: sources are (h/2, w/2)                                            */
  int ksize, stride, pad; /* 1 or 3; 1 or 2; 0 or 1                                         */
  const float* gn_mean;   /* [batch, 32] or NULL                                            */
  const float* gn_rstd;
  const float* gn_gamma;  /* [c0 + c1]                                                      */
  const float* gn_beta;
  int act;                /* 0 none, 1 SiLU                                                 */
  dgq_quant_t q;
  int pad_quantized;      /* 1: out-of-image taps = quantize(0) (unfold path) */
  void* out;              /* fp16 [M, ldo], M = batch*ho*wo; columns >= K are zero          */
  int ldo;
  uint8_t* codes;         /* optional u8 [M, K] integer codes (verification)                */
} dgq_producer_t;
int dgq_act_producer(const dgq_producer_t* host_args, void* stream);

/* GroupNorm(32, C) statistics over a (two-source) NHWC fp16 tensor -> mean, rstd [batch, 32].
 * scratch: >= batch * 64 * 64 floats.                                                         */
int dgq_gn_stats(const void* src0, const void* src1, int src_is_f32, int c0, int c1, int batch, int hw,
                 float eps, float* mean, float* rstd, float* scratch, void* stream);

/* LayerNorm(C, eps) + up to three UniformAffineQuantizers of the same normalised row
 * (norm1 -> to_q/to_k/to_v, diffusers_rewrite/sd.py:252-269; quant/quant_layer.py:640-641).
 * x: fp16/fp32 [m, c]; out[i]: fp16 [m, c].                                                    */
int dgq_ln_quant(const void* x, int src_is_f32, int m, int c, const float* gamma, const float* beta,
                 float eps, int n_out, const dgq_quant_t* host_q, void* const* host_out, void* stream);

/* row quantizer without a norm (cross-attention to_k/to_v on encoder_hidden_states, fp32 or fp16
 * input [m, c]) -- same outputs as dgq_ln_quant                                                */
int dgq_row_quant(const void* x, int src_is_f32, int m, int c, int n_out, const dgq_quant_t* host_q,
                  void* const* host_out, uint8_t* const* host_codes, void* stream);

/* GEGLU (diffusers_rewrite/sd.py:215-218: x1 * gelu_erf(x2)) + quantizer of ff.net.2.
 * x: fp16/fp32 [m, 2f] -> out fp16 [m, f]                                                       */
int dgq_geglu_quant(const void* x, int src_is_f32, int m, int f, dgq_quant_t q, void* out, void* stream);

/* ---- qGEMM: C[m, n] = (A[m, k] . B[n, k]^T) * row_scale[m] * scale[n] + bias[n] (+ temb[m / rows_per_batch, n])
 *      (+ resid[m, n]) on tcgen05/TMEM, TMA-fed (replaces F.linear / F.conv2d / W.view(Co,-1) @ x_unf,
 *      quant/quant_layer.py:649-659; residual and temb adds quant/quant_block.py:105-117,165-186).
 * A: fp16 [m, lda]; B: fp16 [n_pad, ldb] = (code - zp) from dgq_pack_weight; k multiple of 8;
 * out: fp16 [m, ldc] (and/or out_f32 [m, ldc]); ldc, n multiples of 8.                          */
typedef struct {
  const void* a; int lda;
  const void* b; int ldb;
  int m, n, k;
  const float* scale;  /* [n] per-out-channel weight delta, or NULL (=1) */
  const float* row_scale; /* activation delta applied per row: row_scale[(m % row_period)], or NULL.
                             With integer A (dgq_quant_t.emit_int) and integer B the accumulation is
                             exact, so the result matches the fp32 reference to rounding of the sum. */
  int row_period;      /* 1: scalar delta */
  const float* bias;   /* [n] or NULL */
  const void* temb;    /* [m / rows_per_batch, ld_temb] or NULL; fp32 when ep_is_f32 else fp16 */
  int rows_per_batch, ld_temb;
  const void* resid;   /* [m, ld_resid] or NULL; fp32 when ep_is_f32 else fp16 */
  int ld_resid;
  void* out; int ldc;  /* fp16 result, optional */
  float* out_f32;      /* fp32 result, optional (at least one of out / out_f32) */
  int ep_is_f32;       /* dtype of temb / resid */
  /* ---- fused epilogues: the GEMM result never reaches HBM in fp32, the NEXT op's fp16 operand is
   *      written instead (quantised with that op's activation quantizer q2; emit_int allowed).
   * DGQ_EPI_GEGLU : ff.net.0.proj (diffusers_rewrite/sd.py:210-218) -> x1 * gelu_erf(gate), quantised
   *                 for ff.net.2; B rows must be interleaved [32 x1 | 32 gate] per 64 GEMM columns
   *                 (dgq_b200 does this at pack time), n = 2f, out = fp16 [m, ldc >= f], resid unused.
   * DGQ_EPI_QKV   : to_q / to_k / to_v (sd.py:171-180) -> aqtizer_q/k/v -> head-split operand of
   *                 dgq_attention: out = fp16 [b, heads, tokens, dp] (transpose == 0) or V^T
   *                 [b, heads, dp, tp] (transpose == 1); m = b * tokens, n = heads * d.  q2.mode KWISE
   *                 indexes the channel inside the head, ROWWISE the token (minus skip_first);
   *                 skip_first == 1: token 0 bypasses the quantizer (start-peak).  Padding of the
   *                 destination (dp > d, tp > tokens) is not written: pre-zero it.                  */
  int epi;             /* DGQ_EPI_* */
  dgq_quant_t q2;
  int heads, d, dp, tokens, tp, transpose, skip_first;
  /* ---- DGQ_EPI_QKV producing the K operand of dgq_attention (k_split = 1): the value
   *      kfold[channel] * aqtizer_k(k) is written as an fp16 (hi | lo) pair, out = [b, heads, tokens, 2 dp] with hi
   *      at columns 0..dp-1 and lo = fp16(value - hi) at dp..2dp-1.  kfold = the per-channel delta of aqtizer_q when
   *      that quantizer is K-wise (the Q operand is then the bare integer code - zp: q2.emit_int = 1 is accepted for
   *      every q2.mode under DGQ_EPI_QKV), or NULL.                                                            */
  const float* kfold;
  int k_split;
  /* ---- dgq_gemm_i8 only (ignored by dgq_gemm_f16): integer zero-point corrections, see below */
  const int32_t* colsum; /* [n] sum_k B[n, k] (dgq_weight_to_i8)                                     */
  const int32_t* b_off;  /* [n] e_n = (offset subtracted from the weight codes) - (weight zero point),
                            or NULL when the zero point itself was subtracted (W4: codes - zp fit s8)  */
  const float* row_zp;   /* activation zero point, indexed like row_scale: row_zp[m % row_period]     */
  /* ---- dgq_gemm_i8, DGQ_EPI_PLAIN: implicit 3x3 / stride 1 / pad 1 convolution (conv_h > 0).  `a` is then the
   *      NHWC u8 code tensor [conv_b, conv_h, conv_w, conv_c] itself (what a 1x1 producer writes; lda unused),
   *      m = conv_b * conv_h * conv_w, k = 9 * conv_c with B in tap-major K order (dgq_pack_weight), scalar
   *      activation scale (row_period = 1).  Replaces F.conv2d(x_hat, w_hat, padding=1) of the per-tensor path
   *      (quant/quant_layer.py:659): padding taps are EXACT zeros there, i.e. code = zero point, which the
   *      epilogue restores per border class from conv_csoob[class][n] (dgq_conv_oob_colsum), class =
   *      3 * (0 top | 1 | 2 bottom) + (0 left | 1 | 2 right).                                               */
  int conv_b, conv_h, conv_w, conv_c;
  const int32_t* conv_csoob; /* [9][conv_ldoob] */
  int conv_ldoob;
} dgq_gemm_t;
#define DGQ_EPI_PLAIN 0
#define DGQ_EPI_GEGLU 1
#define DGQ_EPI_QKV 2
int dgq_gemm_f16(const dgq_gemm_t* host_args, void* stream);

/* ---- the same GEMM on tcgen05 kind::i8 (2x the MMA rate; profiles/r2_probes.txt, r2_gemm_i8_vs_f16.txt) for
 *      layers whose activation scale is constant along K (scalar or row-wise: quant/quant_layer.py:295-299 with a
 *      () or per-row delta -- every layer of the group_num = 1 configs).
 * A: u8 activation CODES [m, lda] (producers with dgq_quant_t.emit_int = 2); B: s8 [n_pad, ldb] from
 * dgq_weight_to_i8; k, lda, ldb multiples of 16.  The accumulation is exact (s32) and the zero points are removed
 * in integer arithmetic in the epilogue:
 *   sum_k (a_mk - za_m)(w_nk - wz_n) = acc_mn - za_m * colsum_n + e_n * (rowsum_m - k * za_m)
 * (rowsum_m = sum_k a_mk is computed inside the kernel, only when b_off != NULL), then
 *   C = that * row_scale[m] * scale[n] + bias[n] (+ temb) (+ resid), epilogues as dgq_gemm_f16.
 * row_scale (activation delta) and row_zp are required.                                                     */
int dgq_gemm_i8(const dgq_gemm_t* host_args, void* stream);

/* weight codes (u8 [n_pad, k_out], GEMM K order, from dgq_pack_weight) -> s8 operand of dgq_gemm_i8 + tables:
 *   operand[n, k] = code - off_n, off_n = zp[n] when qmax <= 127 (W4: exact, b_off[n] = 0) else 128
 *   (b_off[n] = 128 - zp[n]); colsum[n] = sum_k operand[n, k]; rows >= n are zero.                           */
int dgq_weight_to_i8(const uint8_t* codes, const float* zp, int n, int n_pad, int k_out, float qmax,
                     int8_t* operand, int32_t* colsum, int32_t* b_off, void* stream);
/* border-class tables of the implicit 3x3 convolution: operand s8 [n_pad, 9 * c] (tap-major) ->
 * csoob[cls][n] = sum over the taps that lie outside the image for border class cls of sum_c operand[n, tap*c + ch]
 * (cls 4, the interior, is all zero).  csoob: int32 [9][n_pad].                                               */
int dgq_conv_oob_colsum(const int8_t* operand, int n_pad, int c, int32_t* csoob, void* stream);

/* ---- attention with quantised operands and quantised softmax map
 *      (Attention.Attention_forward, diffusers_rewrite/sd.py:151-207; T2ILogQuantizer) ---------
 * dgq_qkv_pack: GEMM output [b*t, heads*d] fp16 -> quantised, head-split, zero-padded operand:
 *   transpose == 0: out fp16 [b, heads, t, dp]         (Q and K)
 *   transpose == 1: out fp16 [b, heads, dp, tp]        (V^T), tp = round_up(t, 8)
 *   q.mode: NONE / SCALAR / KWISE (index d) / ROWWISE (index t - skip_first).
 *   skip_first == 1: token 0 bypasses the quantizer (start-peak, sd.py:176-180).
 *   q.emit_int = 1 (Q): the bare integer code - zp in every mode.  kfold / k_split (K): as DGQ_EPI_QKV --
 *   value * kfold[channel] written as an fp16 (hi | lo) pair, out = [b, heads, t, 2 dp].                       */
int dgq_qkv_pack(const void* x, int src_is_f32, int ldx, int b, int t, int heads, int d, int dp, int tp,
                 int transpose, int skip_first, dgq_quant_t q, const float* kfold, int k_split, void* out,
                 void* stream);

#define DGQ_MAP_NONE 0    /* use_aq off: plain softmax                                   */
#define DGQ_MAP_UNIFORM 1 /* UniformAffineQuantizer, always_zero (zp = 0)                */
#define DGQ_MAP_LOG2 2    /* T2ILogQuantizer                                             */
typedef struct {
  const void* q;   /* fp16 [b, heads, t, dp]  */
  const void* k;   /* fp16 [b, heads, s, dp]  */
  const void* vt;  /* fp16 [b, heads, dp, sp] */
  int b, heads, t, s, sp, d, dp;
  float scale;          /* d^-0.5 */
  int map_mode;         /* DGQ_MAP_* */
  int real_time;        /* LOG2: delta = max over the whole (b,h,t,s) map of this call */
  int start_peak;       /* column 0 bypasses the map quantizer (sd.py:191-195) */
  const float* delta;   /* device scalar (static delta); ignored when real_time */
  float qmax;
  float* row_max;       /* scratch [b*heads*t] */
  float* row_sum;       /* scratch [b*heads*t] */
  float* gmax;          /* [1]: receives the real-time delta (zeroed by the call itself)  */
  void* out;            /* [b*t, ldo], head h at columns h*d .. h*d+d-1; fp32 when out_is_f32 */
  int ldo;
  int out_is_f32;
  uint8_t* codes;       /* optional u8 [b, heads, t, s]: integer codes of the map (verification only) */
  dgq_quant_t out_q;    /* quantizer of the consuming QuantLayer (to_out[0], quant/quant_layer.py:640-641)
                           applied to O before the store: KWISE index = column, ROWWISE = row % period;
                           mode NONE: O is stored as is; emit_int = 2: u8 codes (ldo in bytes) */
  /* score operands in exact form (sd.py:171-183: q_hat . k_hat): q holds the bare integers code - zp of aqtizer_q
   * and q_scale[token % q_scale_period] its scalar / per-token delta (NULL: none, or folded into k); k_split = 1:
   * k is [b, heads, s, 2 dp] = fp16 (hi | lo) of the fully scaled key (dgq_gemm_t.k_split).                     */
  const float* q_scale;
  int q_scale_period;
  int k_split;
} dgq_attn_t;
int dgq_attention(const dgq_attn_t* host_args, void* stream);

/* ---- small glue kernels ---------------------------------------------------------------------*/
/* Timesteps.forward (diffusers_rewrite/sd.py:25-39): [n] fp32 -> fp16/fp32 [n, dim] (cos | sin) */
int dgq_timestep_embedding(const float* t, int n, int dim, void* out_f16, float* out_f32, int ldo,
                           void* stream);
/* fp32 NCHW -> fp16/fp32 NHWC with channel padding (conv_in input), and back (conv_out result)  */
int dgq_nchw_to_nhwc(const float* x, int b, int c, int hw, int c_pad, void* out, int out_is_f32, void* stream);
int dgq_nhwc_to_nchw(const void* x, int src_is_f32, int b, int c, int hw, int ldx, float* out, void* stream);
/* y = silu(x) elementwise (SiLU on the time embedding, quant/quant_block.py:107); fp16 or fp32  */
int dgq_silu(const void* x, int is_f32, int64_t n, void* out, void* stream);
/* P[r, :] = softmax(scale * S[r, :]), fp32 [rows, lds] -> fp16 [rows, ldo]: the attention map of the VAE decoder's
 * mid-block attention (F.scaled_dot_product_attention, diffusers/src/diffusers/models/attention_processor.py:1244;
 * one head of 512 channels), between two dgq_gemm_f16 calls.  cols % 4 == 0.                     */
int dgq_softmax_rows(const float* s, int64_t rows, int cols, int64_t lds, float scale, void* out_f16, int64_t ldo,
                     void* stream);
/* out = a + b; fp16 or fp32                                                                     */
int dgq_add(const void* a, const void* b, int is_f32, int64_t n, void* out, void* stream);

/* ---- sampler step (the caller of the path): CFG combine + scheduler update + next model
 *      input in one pass.  Replaces pipeline_stable_diffusion.py:1022-1047 around PNDMScheduler.step_plms
 *      (schedulers/scheduling_pndm.py:321-449) and EulerAncestralDiscreteScheduler.step / scale_model_input
 *      (schedulers/scheduling_euler_ancestral_discrete.py:239-260,323-414), all of which are linear maps:
 *        eps      = use_cfg ? u + guidance * (c - u) : unet_out        (u, c = the two halves of unet_out)
 *        out      = cx * x + c_eps * eps + sum_k c_hist[k] * hist[k] + c_noise * noise
 *        model_in = out * in_scale   (written twice, back to back, when dup: the CFG batch)
 *      fp32, n = elements of ONE latent batch, multiple of 4.                                          */
typedef struct {
  const float* unet_out; /* [use_cfg ? 2n : n] */
  int64_t n;
  float guidance;
  int use_cfg;
  float* eps_store;      /* optional [n]: eps is kept here (the multistep history) */
  const float* x;        /* [n] */
  float cx, c_eps;
  const float* hist[4];  /* earlier eps, NULL = unused */
  float c_hist[4];
  const float* noise;    /* [n], read only when c_noise != 0 */
  float c_noise;
  float* out;            /* [n] */
  float* model_in;       /* optional [dup ? 2n : n] */
  float in_scale;
  int dup;
} dgq_sampler_step_t;
int dgq_sampler_step(const dgq_sampler_step_t* host_args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGQ_B200_H_ */
