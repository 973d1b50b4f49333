/* libw2v2_sm100.so - C ABI of the B200 (sm_100a) Wav2Vec2 forward / CTC kernels.
 *
 * The reference (thevasudevgupta/gsoc-wav2vec2) has NO native plugin/FFI layer: every op on its
 * hot path is a Keras layer or tf.* call inside the un-vendored tensorflow==2.5 wheel.  Each entry
 * point below therefore replaces one group of those op call sites (cited per function as
 * reference file:line) underneath the reference's own Python surface
 * (src/wav2vec2/__init__.py:1-4: Wav2Vec2Config / Wav2Vec2Model / Wav2Vec2ForCTC / CTCLoss).
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (no hidden allocation, no ownership
 *     transfer); activations are channels-last row-major [B, T, C] like the reference
 *     (modeling.py:188); "hi"/"lo" are bf16 planes with value ~= hi + lo (lo may be NULL where
 *     documented: single-pass bf16 mode);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream,
 *     stateless and thread-safe;
 *   - return 0 on success, < 0 on an argument error, > 0 = cudaError_t.  The message for the last
 *     failure on the calling thread is w2v2_last_error_string().  There is no CPU fallback.
 */
#ifndef W2V2_H_
#define W2V2_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int w2v2_version(void);
const char* w2v2_last_error_string(void);

/* ---------------------------------------------------------------------------------------------
 * Dense contraction  D[b, t, :] = epilogue( A[b, t, :K] . W[n, :K] )   (tcgen05 / TMA / TMEM)
 * Replaces: tf.keras.layers.Dense at encoder.py:15-18,24-31 (q/k/v/out), encoder.py:99-104,127-128
 * (FFN), feature_extractor.py:89,94 (projection), modeling.py:231,254 (lm_head); and
 * tf.keras.layers.Conv1D for extractor layers 1..6 at feature_extractor.py:31-37,55 as an implicit
 * GEMM (A rows = overlapping k*Cin windows of the channels-last input).
 * ------------------------------------------------------------------------------------------- */
#define W2V2_GEMM_MN_MAJOR 2u /* D[m][n] = sum_r X[r][m] Y[r][n]: both operands row-major [reduction rows][columns] as the forward
                                stores them (weight gradients dW = X^T dY without transposed copies).  a_hi = X, a_rows = number
                                of reduction rows, a_row_stride = leading dimension of X, rows_per_batch = columns of X (= output
                                rows); w_hi = Y with leading dimension w_row_stride, N = columns of Y; single pass, 1-SM tiles
                                (block_n 64 / 128).  batch = number of split-K slices: slice b reduces rows [b K, (b+1) K), K x batch
                                >= a_rows (rounded up to 64 per slice); with batch > 1 the slices ADD into out_f32 with fp32 atomics
                                (the caller zeroes it; no bias / residual / bf16 output). */
#define W2V2_GEMM_GELU 1u /* GELU after bias (feature_extractor.py:58, encoder.py:127): erf-exact in 3-pass (parity) mode;
                             single-pass mode uses the bf16-grade tanh form (|err| < 5e-4) */

/* Precision modes - the `passes` argument of every tensor-core entry point:
 *    1  bf16     one MMA per product on bf16 operands                                   (throughput mode, ~3e-2 on the logits)
 *    3  bf16x3   split-bf16: hi*hi + lo*hi + hi*lo, planes bf16 hi + bf16 lo              (parity mode, ~1.4e-4)
 *   17  fp16     one MMA on fp16 operands; planes hold activation * 2^4, weight * 2^11    (~3e-3, the reference's own 4e-3 logits bar)
 *   19  fp16x3   split-fp16, planes fp16 hi + fp16 lo (same scaling)
 *   25  fp16f8   fp16 main product + BOTH cross terms as e4m3 MMAs (kind::f8f6f4, twice the bf16 rate): planes = fp16 hi and a
 *                byte plane [rows][2 K] holding, per 64-element k-block, 64 x e4m3((v - hi) * 2^6) then 64 x e4m3(hi * 2^-6)
 *                (weights: the two halves swapped), so the two cross terms are ONE K = 128 e4m3 product       (~2e-4, 2 MMA units)
 * All fp16 modes accumulate at scale 2^15 in fp32 and un-scale in the epilogue. */
#define W2V2_MODE_BF16 1
#define W2V2_MODE_BF16X3 3
#define W2V2_MODE_FP16 17
#define W2V2_MODE_FP16X3 19
#define W2V2_MODE_FP16F8 25
#define W2V2_OUT_BF16 0   /* out_hi (/ out_lo): bf16 planes of the value */
#define W2V2_OUT_FP16 1   /* fp16 planes of value * 2^4 (hi, optional lo = residual) */
#define W2V2_OUT_FP16F8 2 /* out_hi = fp16(value * 2^4), out_lo = e4m3 pair plane [rows][2 N] bytes (N % 64 == 0) */

#define W2V2_GEMM_GELU_TANH 4u /* GELU after bias in tf.nn.gelu(approximate=True) form - the reference's `is_gelu_approx` switch
                                  (config.py:14); fp32-grade in every precision mode */

typedef struct w2v2_gemm_args {
  /* A operand: bf16 planes, logical shape [batch][a_rows][a_row_len], element strides given. */
  const void* a_hi;
  const void* a_lo;        /* NULL unless passes == 3 */
  int64_t a_row_len;       /* elements addressable in one row (>= K; conv: k*Cin) */
  int64_t a_rows;          /* rows addressable per batch entry (TMA bound; >= rows_per_batch) */
  int64_t a_row_stride;    /* elements between consecutive rows (conv: stride*Cin - rows overlap) */
  int64_t a_batch_stride;  /* elements between batch entries */
  /* W operand: bf16 [w_rows][K] row-major ("K-major"), w_rows >= N rounded up to block_n. */
  const void* w_hi;
  const void* w_lo;        /* NULL unless passes == 3 */
  int32_t w_rows;
  int32_t K;               /* multiple of 64 */
  int32_t N;               /* output columns; leading dimension of residual and outputs */
  int32_t rows_per_batch;  /* output rows per batch entry */
  int32_t batch;
  int32_t passes;          /* precision mode W2V2_MODE_*: 1 = bf16 operands; 3 = split-bf16 (hi*hi + lo*hi + hi*lo); 17 / 19 / 25 */
  int32_t kb_split;        /* 0, or: 64-wide k-blocks >= kb_split come from (k - 64*kb_split, row+1) */
  int32_t block_n;         /* 0 = auto (256/128/64/32) */
  int32_t max_ctas;        /* 0 = one persistent CTA per SM */
  int32_t cluster;         /* 0 = auto (256-wide tiles: cta_group::2 CTA pairs), 1 = single CTAs,
                              3 = 1-SM MMAs with pair-multicast weight tiles */
  uint32_t flags;          /* W2V2_GEMM_* */
  const float* bias;       /* [N] (or [batch][N], see bias_batch_stride) or NULL */
  const float* scale;      /* NULL, or per-column scale applied to the accumulator before the bias:
                              [N] or [batch][N]  (GroupNorm of extractor layer 0 folded as scale/shift) */
  int64_t bias_batch_stride; /* elements between batch entries of bias / scale; 0 = shared by all entries */
  const float* residual;   /* fp32 [batch*rows_per_batch][N] or NULL, added after bias/GELU */
  const int32_t* row_valid;/* [batch] or NULL: rows t >= row_valid[b] are stored as zeros
                              (padded frames, encoder.py:253) */
  float* out_f32;          /* any subset of the three outputs; out_lo requires out_hi */
  void* out_hi;
  void* out_lo;
  /* optional: the residual term is LayerNorm(residual) recomputed on the fly, fmaf((r - mean) * rstd, gamma, beta) with the
   * per-row (mean, rstd) that w2v2_ln_rows wrote to `stats` - so a post-norm layer never materialises the fp32 LayerNorm
   * output: x1 = LN(y) is consumed as bf16 by the next GEMM and as "LN(y)" by the next residual add (encoder.py:119-132). */
  const float* res_ln_stats;  /* [batch*rows_per_batch][2] or NULL */
  const float* res_ln_gamma;  /* [N] */
  const float* res_ln_beta;   /* [N] */
  int64_t w_row_stride;       /* W2V2_GEMM_MN_MAJOR only: leading dimension (elements) of Y */
  /* optional epilogue steps of the TRAINING forward, applied in this order after bias / GELU:
   *   dropout (feature_extractor.py:95, encoder.py:118,128): the stateless stream of w2v2_dropout_rows for (drop_seed, drop_site)
   *     over the flat output element index, survivors x 1/(1-p); then the residual is added;
   *   SpecAugment (modeling.py:193-199, spec_augment.py:119-128): rows whose row_replace_mask byte is non-zero are written as
   *     row_replace_value[0:N] (masked_spec_embed) instead; then row_valid zeroing. */
  const uint8_t* row_replace_mask; /* [batch*rows_per_batch] or NULL */
  const float* row_replace_value;  /* [N] */
  float drop_p;                    /* 0 = off */
  uint32_t drop_site;
  uint64_t drop_seed;
  int32_t out_format;              /* layout of out_hi / out_lo: W2V2_OUT_BF16 (0), W2V2_OUT_FP16 (1), W2V2_OUT_FP16F8 (2) */
  /* LayerNorm folded into the Dense that follows it (encoder.py:116-132: LN -> q/k/v Dense, LN -> intermediate Dense):
   *   LN(x) W + b = rstd (x (gamma o W)) - rstd mean colsum(gamma o W) + (beta W + b)
   * The A operand is then the UN-normalised x, W is packed as gamma o W, `bias` = beta W + b, `scale` = colsum(gamma o W) [N]
   * (NOT a multiplier in this mode), and the per-row (mean, rstd) come from ln_fold_stats: [ln_fold_parts][rows][2] partial
   * (sum, sum of squares) pairs over the K input columns, written by the producer of x through row_stats_out. */
  int32_t ln_fold_parts;           /* 0: ln_fold_stats holds (mean, rstd) per row [rows][2]; > 0: partial sums [parts][rows][2] */
  const float* ln_fold_stats;      /* NULL = no fold */
  float ln_eps;                    /* epsilon of the folded LayerNorm and of res_ln_* when given as partial sums */
  int32_t res_ln_parts;            /* 0: res_ln_stats holds (mean, rstd) per row; > 0: that many partial (sum, sum of squares) pairs per row, laid out [parts][rows][2] */
  float* row_stats_out;            /* optional [N / 64][rows][2]: (sum, sum of squares) of the fp32 result per 64-column group
                                      (needs out_f32 semantics: the statistics are those of the value written to out_f32; N % 64 == 0) */
  /* optional, with row_stats_out: the per-row (mean, rstd) [rows][2] over all N columns (biased variance, ln_eps) - what
   * w2v2_row_stats_finalize computes from row_stats_out - produced by this call itself.  In the cta_group::2 kernel the LAST of the
   * N / 64 column groups to finish a 32-row block (an arrival counter per block in row_stats_counter, [ceil(rows / 32)] uint32,
   * zero before the first use; the kernel leaves it zero) sums the partials in index order: bit-identical to the separate
   * launch, without it.  Other tile shapes run w2v2_row_stats_finalize on the same stream.  Needs batch == 1 or rows_per_batch % 32 == 0. */
  float* row_stats_final;
  uint32_t* row_stats_counter;
} w2v2_gemm_args;

int w2v2_gemm_bf16(const w2v2_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Extractor layer 0: Conv1D(k=10, s=5, Cin=1, no bias) + GroupNorm(groups == channels) + GELU.
 * Replaces feature_extractor.py:31-37,40-47,54-59 and GroupNormalization._apply_normalization,
 * tensorflow_addons.py:207-231 (per-(b,c) mean / biased variance over TIME, eps 1e-5).
 *   1. w2v2_wave_stats : stats[b][0:10] = sum_t x[5t+j];  stats[b][10:65] = upper triangle of
 *      G[i][j] = sum_t x[5t+i] x[5t+j]  (fp64, T0 = 1 + (L-10)/5 windows).
 *   2. w2v2_conv0_fold : folds mean / rstd / gamma / beta into per-(b,c) weights [B][10][C] and
 *      bias [B][C]  (sum_t y = w.s, sum_t y^2 = w^T G w because Cin == 1).
 *   3. w2v2_conv0      : y = gelu?(bias + sum_j w[j][c] x[5t+j]) -> bf16 hi(/lo) or fp32, written once.
 *      With weights_batch_stride == 0 and the raw kernel it is the plain conv (+bias) used by the
 *      "layer"-norm extractor variant (feature_extractor.py:48-50), followed by w2v2_ln_rows.
 * ------------------------------------------------------------------------------------------- */
int w2v2_wave_stats(const float* wave, int batch, int num_samples, double* stats /*[batch][65]*/, void* stream);
int w2v2_conv0_fold(const float* kernel /*[10][C] (TF layout [k,1,C])*/, const float* gamma, const float* beta,
                    const double* stats, int batch, int num_samples, int channels, float eps,
                    float* folded_w /*[batch][10][C] or NULL*/, float* folded_b /*[batch][C]: shift*/,
                    float* scale /*[batch][C] or NULL: gamma * rstd*/, void* stream);
/* Tensor-core route for layer 0: windows of the waveform as zero-padded bf16 rows a[b][t][0:64] (taps 0..9 used), to be
 * multiplied by the raw kernel W[C][64] with w2v2_gemm_bf16 (scale / bias per batch entry, GELU). */
int w2v2_conv0_im2col(const float* wave, int batch, int num_samples, void* a_hi, void* a_lo, void* stream);
/* Fused layer 0 (the HBM-bound kernel of the path): out = gelu(scale[b][c] * conv(wave)[b][t][c] + shift[b][c]) as bf16
 * hi (passes == 1; tanh-form bf16-grade GELU) or hi + lo planes (passes == 3; erf-exact GELU).  kernel = raw TF kernel
 * [10][C], scale / shift = outputs of w2v2_conv0_fold.  The window products run on the tensor cores straight from a
 * shared-memory copy of the waveform slice: no im2col tensor, the activation is written once. */
int w2v2_conv0_gn_gelu(const float* wave, int batch, int num_samples, int channels, const float* kernel /*[10][C]*/,
                       const float* scale /*[batch][C]*/, const float* shift /*[batch][C]*/, void* out_hi,
                       void* out_lo /*NULL unless passes == 3*/, int passes,
                       int gelu_approx /*0: erf GELU (config.py:14 default); 1: tf.nn.gelu(approximate=True)*/, void* stream);
/* Fused layer 0 of the "layer"-norm extractor (robust / large checkpoints: feature_extractor.py:48-50,54-59 with
 * config.feature_extractor_norm_type == "layer"): out = gelu(LayerNorm_c(conv(wave)[b][t][:] + conv_bias) * gamma + beta), the
 * LayerNorm over the 512 channels of a frame (biased variance, two-pass, eps) computed inside the CTA that holds them - replaces
 * w2v2_conv0 (fp32 output) + w2v2_ln_rows.  Same planes / passes / gelu_approx as w2v2_conv0_gn_gelu. */
int w2v2_conv0_ln_gelu(const float* wave, int batch, int num_samples, int channels, const float* kernel /*[10][C]*/,
                       const float* conv_bias /*[C] or NULL*/, const float* gamma /*[C]*/, const float* beta /*[C]*/, float eps,
                       void* out_hi, void* out_lo /*NULL unless passes == 3 or 25*/, int passes, int gelu_approx, void* stream);
int w2v2_conv0(const float* wave, int batch, int num_samples, int channels, const float* weights,
               int weights_batch_stride, const float* bias /*or NULL*/, int bias_batch_stride, int gelu,
               float* out_f32, void* out_hi, void* out_lo, void* stream);

/* LayerNormalization over the last axis (biased variance) with optional GELU (gelu: 0 none, 1 erf-exact, 2 tf-approximate
 * tanh form = is_gelu_approx, config.py:14); fp32 in, outputs
 * fp32 and/or bf16 hi(/lo).  Replaces tf.keras.layers.LayerNormalization at encoder.py:96-108,
 * 116,121,126,132,232-234,268,275 and feature_extractor.py:50,86-88,93. */
int w2v2_ln_rows(const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int d, int gelu,
                 float* out_f32, void* out_hi, void* out_lo, void* stream);
/* same, additionally writing the per-row (mean, rstd) to stats[rows][2] (see w2v2_gemm_args.res_ln_stats) */
int w2v2_ln_rows_stats(const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int d, int gelu,
                       float* out_f32, void* out_hi, void* out_lo, float* stats, void* stream);

/* same, with the layout of out_hi / out_lo chosen by out_format (W2V2_OUT_BF16 / W2V2_OUT_FP16 / W2V2_OUT_FP16F8) */
int w2v2_ln_rows_ex(const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int d, int gelu,
                    float* out_f32, void* out_hi, void* out_lo, float* stats, int out_format, void* stream);

/* (mean, rstd) per row [rows][2] from the partial (sum, sum of squares) pairs [nparts][rows][2] that a residual GEMM wrote through
 * w2v2_gemm_args.row_stats_out: the statistics of the LayerNorm (over `dim` columns, biased variance) that the NEXT GEMM folds
 * (ln_fold_stats with ln_fold_parts == 0) and that the next residual add recomputes (res_ln_stats with res_ln_parts == 0). */
int w2v2_row_stats_finalize(const float* parts, int nparts, int64_t rows, int dim, float eps, float* stats, void* stream);

/* Wav2Vec2Processor._normalize (processor.py:101-106) on the device: per utterance (x - mean) / sqrt(var + eps) with the
 * biased variance over its lengths[b] real samples (NULL: all num_samples); the padded tail is written as 0
 * (normalise BEFORE padding, data_utils.py:233,63).  In place (out == wave) is allowed. */
int w2v2_normalize_utterances(const float* wave, const int32_t* lengths, int batch, int num_samples, float eps, float* out,
                              void* stream);

/* fp32 -> bf16 hi(/lo) planes (weight packing). */
int w2v2_split_bf16(const float* x, int64_t n, void* hi, void* lo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-head attention core on the packed projection qkv[b][t][0:3d] = [q | k | v] (q pre-scaled):
 * ctx[b][t][h*64:(h+1)*64] = softmax_k(q.k) v.  Replaces TransformerAttention.get_context and
 * _prepare_either_qkv, encoder.py:34-54, and the additive key mask of encoder.py:256-263
 * (kv_len[b] = number of real frames, or NULL).
 * ------------------------------------------------------------------------------------------- */
int w2v2_attn_fwd(const void* qkv_hi, const void* qkv_lo, int batch, int frames, int num_heads, int head_size,
                  const int32_t* kv_len, void* out_hi, void* out_lo, int passes, void* stream);
/* same with the precision modes 17 / 19 (fp16 planes of q / k / v * 2^4) and the layout of out_hi / out_lo chosen by out_format
 * (W2V2_OUT_*; W2V2_OUT_FP16F8 feeds an output projection running in mode 25) */
int w2v2_attn_fwd_ex(const void* qkv_hi, const void* qkv_lo, int batch, int frames, int num_heads, int head_size,
                     const int32_t* kv_len, void* out_hi, void* out_lo, int passes, int out_format, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Positional convolution + GELU + residual: out = resid + gelu(bias + grouped_conv_k(x)).
 * Replaces PositionalConvEmbedding.call encoder.py:177-181, Conv1DWithWeightNorm.call
 * tensorflow_addons.py:50-53 (weight norm folded into w by the host) and encoder.py:265.
 * w_hi/w_lo: bf16 packed [groups][ktaps][cpg/8][cpg(out)][8(in)].
 * ------------------------------------------------------------------------------------------- */
typedef struct w2v2_posconv_args {
  const void* x_hi;      /* bf16 [batch][frames][hidden] */
  const void* x_lo;      /* NULL unless passes == 3 */
  const void* w_hi;
  const void* w_lo;      /* NULL unless passes == 3 */
  const float* bias;     /* [hidden] */
  const float* resid;    /* fp32 [batch][frames][hidden] */
  float* out_f32;        /* fp32 [batch][frames][hidden] */
  int32_t batch, frames, hidden, groups, ktaps, passes;
  /* training (zero for inference): */
  float* pre_out;        /* optional fp32 [batch][frames][hidden]: bias + conv, the pre-activation kept for the backward */
  int32_t shift;         /* the tap window starts at frame t - ktaps/2 + shift */
  int32_t linear;        /* 1: out = resid + conv(x) (no bias, no GELU) - with flipped / transposed taps and shift = 1
                            this is the input gradient of the convolution */
  int32_t gelu_approx;   /* 1: tf.nn.gelu(approximate=True) instead of the erf form (config.py:14, encoder.py:181) */
} w2v2_posconv_args;

int w2v2_posconv(const w2v2_posconv_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CTC loss and its gradient w.r.t. the unnormalised logits.  Replaces tf.nn.ctc_loss as called by
 * CTCLoss.call, losses.py:29-45: blank = pad_id, label_length = #labels != pad, logit_length =
 * frames for every utterance, per-utterance negative log-likelihood scaled by `scale`
 * (= 1 / division_factor); the caller sums over the batch (Keras SUM reduction, losses.py:6).
 * workspace: w2v2_ctc_workspace_bytes(batch, frames, max_label_len) bytes (the alpha and beta tables).  grad_logits may be NULL.
 * ------------------------------------------------------------------------------------------- */
int64_t w2v2_ctc_workspace_bytes(int batch, int frames, int max_label_len);
int w2v2_ctc_loss(const float* logits /*[batch][frames][vocab]*/, const int32_t* labels /*[batch][max_label_len]*/,
                  int batch, int frames, int vocab, int max_label_len, int blank, float scale, void* workspace,
                  int64_t* reserved, float* loss_per_sample /*[batch]*/, float* grad_logits, void* stream);

/* Greedy CTC decode, device part: ids[r] = argmax_k logits[r][k] (first maximum), the input of
 * Wav2Vec2Processor.decode (processor.py:71-89; tests/test_wav2vec2.py:159-165). */
int w2v2_frame_argmax(const float* logits, int64_t rows, int vocab, int32_t* ids, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage-1 fine-tune step (src/main.py:210-223: the Wav2Vec2 body is frozen, only lm_head trains).
 * w2v2_lm_head_wgrad: d loss / d kernel [hidden][vocab] = hidden^T . grad_logits and d loss / d bias
 *   (backward of tf.keras.layers.Dense at modeling.py:231,254; what Keras fit computes at main.py:217-223).
 * w2v2_adam: Keras Adam update (main.py:213, training_utils.py:24-31 supplies lr) on a flat fp32 buffer;
 *   lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) is passed by the host.
 * ------------------------------------------------------------------------------------------- */
int w2v2_lm_head_wgrad(const float* hidden /*[rows][hidden_size]*/, const float* grad_logits /*[rows][vocab]*/,
                       int64_t rows, int hidden_size, int vocab, float* grad_kernel, float* grad_bias, void* stream);
int w2v2_adam(float* weights, const float* grads, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2,
              float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage-2 fine-tune step (src/main.py:234-250: Keras `fit` differentiates the whole encoder; the conv extractor
 * stays frozen, main.py:236-237).  The matrix products of the backward pass reuse w2v2_gemm_bf16:
 *   dgrad  dX = dY . W^T    A = dY [rows][out], weight operand = the TF Dense kernel itself ([in][out] is W^T, K-major)
 *   wgrad  dW = X^T . dY    W2V2_GEMM_MN_MAJOR: a_hi = X [rows][in], w_hi = dY [rows][out] as stored (MN-major operands),
 *                           out_f32 = the gradient in the TF layout [in][out]
 * and these kernels supply everything else.
 * ------------------------------------------------------------------------------------------- */
/* Backward of LayerNormalization (encoder.py:96-108,232-234, feature_extractor.py:86-88): x = the LN input kept by the
 * forward, dy = gradient of its output.  dx -> fp32 and/or bf16; dgamma / dbeta / colsum (= sum over rows of dx, the
 * bias gradient of the Dense that produced x) are ACCUMULATED (atomicAdd): the caller zeroes them once per step. */
int w2v2_ln_bwd(const float* x, const float* gamma, const float* dy, float eps, int64_t rows, int d, float* dx_f32,
                void* dx_hi, float* dgamma, float* dbeta, float* colsum /*or NULL*/, void* stream);
/* GELU of a saved fp32 pre-activation -> bf16 (training forward keeps the pre-activation for the backward).
 * fast != 0: the tanh-form of the single-pass mode (same values as the GEMM epilogue of inference). */
int w2v2_gelu_rows(const float* pre, int64_t n, int fast, void* out_hi, void* out_lo /*or NULL*/, float drop_p, uint64_t seed,
                   uint32_t site, void* stream);   /* drop_p > 0: dropout on the activated values (encoder.py:128) */
/* out = dy * gelu'(pre) (exact erf form, config.py:14) as bf16, colsum += column sums of the ROUNDED result (bias
 * gradient).  pre == NULL: no activation (plain column sums of dy, out may be NULL). */
int w2v2_dact_colsum(const void* dy_hi, const float* pre, int64_t rows, int cols, void* out_hi, float* colsum, float drop_p,
                     uint64_t seed, uint32_t site, void* stream);   /* drop_p > 0: dy is first masked / rescaled (gradient of a dropout
                                                                       that followed the activation, or a Dense when pre == NULL) */
/* out[n][m] = in[m][n] (bf16), rows m in [rows, out_ld) are written as zeros (K padding of the wgrad GEMMs). */
int w2v2_transpose_bf16(const void* in, int64_t rows, int cols, void* out, int64_t out_ld, void* stream);
/* dHidden[rows][hidden] = dLogits[rows][vocab] . kernel[hidden][vocab]^T (backward of the Dense at modeling.py:231,254), fp32. */
int w2v2_lm_head_dgrad(const float* grad_logits, const float* kernel, int64_t rows, int hidden_size, int vocab, float* out,
                       void* stream);
/* Backward of the attention core (forward: w2v2_attn_fwd; reference encoder.py:34-54): from the packed projection qkv,
 * the context ctx = softmax(q k^T) v and its gradient dctx (all bf16) to dqkv in the packed layout; the q part is
 * multiplied by q_scale (= head_size^-1/2: the forward folds that factor into the q projection, encoder.py:28).
 * workspace: w2v2_attn_bwd_workspace_bytes(batch, frames, num_heads). */
int64_t w2v2_attn_bwd_workspace_bytes(int batch, int frames, int num_heads);
int w2v2_attn_bwd(const void* qkv_hi, const void* ctx_hi, const void* dctx_hi, int batch, int frames, int num_heads,
                  int head_size, const int32_t* kv_len, float q_scale, void* workspace, void* dqkv_hi, float drop_p,
                  uint64_t seed, uint32_t site, void* stream);   /* drop_p, seed, site: those of w2v2_attn_fwd_train */
/* Dropout (tf.keras.layers.Dropout at feature_extractor.py:95, encoder.py:42,118,128,270, modeling.py:253; rate config.py:9).
 * Masks come from a stateless counter-based generator: 64 bits per group of 4 consecutive elements =
 * hash(seed, site, index / 4) (two chained murmur3 finalisers), element kept when its 16-bit lane >= round(p * 65536), kept values scaled by 1 / (1 - p).
 * The backward pass regenerates the mask from (seed, site); w2v2_dropout_mask / w2v2_attn_dropout_mask export it (tests).
 *   w2v2_dropout_rows : out = (resid ? resid : 0) + dropout(x)  (fp32, optional bf16 copy; in place allowed)
 *   w2v2_attn_fwd_train : w2v2_attn_fwd with dropout on the attention probabilities (after the softmax, encoder.py:41-43) */
int w2v2_dropout_rows(const float* x, const float* resid /*or NULL*/, int64_t n, float drop_p, uint64_t seed, uint32_t site,
                      float* out_f32, void* out_hi, void* stream);
int w2v2_dropout_mask(int64_t n, float drop_p, uint64_t seed, uint32_t site, uint8_t* out, void* stream);
int w2v2_attn_dropout_mask(int batch_heads, int frames, float drop_p, uint64_t seed, uint32_t site, uint8_t* out /*[bh][q][k]*/,
                           void* stream);
int w2v2_attn_fwd_train(const void* qkv_hi, const void* qkv_lo, int batch, int frames, int num_heads, int head_size,
                        const int32_t* kv_len, void* out_hi, void* out_lo, int passes, float drop_p, uint64_t seed,
                        uint32_t site, void* stream);

/* Weight gradient of the positional convolution in the TF kernel layout [ktaps][hidden/groups][hidden]
 * (w.r.t. the weight-NORMALISED kernel; the weight-norm chain rule is parameter-sized host algebra). */
int w2v2_posconv_wgrad(const void* x_hi, const void* dpre_hi, int batch, int frames, int hidden, int groups, int ktaps,
                       float* grad_kernel, void* stream);

/* Re-packing of the updated fp32 master weights into the kernels' operand layouts, ONE launch per optimizer step.
 * jobs_dev: device array of jobs; tiles_dev: device array of int32 triples (job index, first row, first column) - one
 * 64 x 64 tile per CTA.  Both tables are built once by the host (the pointers never change: the variables are views of
 * one flat buffer and the packed operands are persistent). */
typedef struct w2v2_pack_job {
  const void* src;       /* fp32 [rows][src_ld] */
  void* dst;             /* bf16 (or fp32 when dst_f32) with leading dimension dst_ld; transposed jobs write dst[c][r] */
  int32_t rows, cols, src_ld, dst_ld;
  int32_t transpose, dst_f32;
  float scale;
  int32_t reserved;
} w2v2_pack_job;
int w2v2_pack_weights(const w2v2_pack_job* jobs_dev, const int32_t* tiles_dev, int num_tiles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* W2V2_H_ */
