/* w2v2_b200.h -- C ABI of the B200-native wav2vec2 speaker hot path (libw2v2_b200.so).
 *
 * The reference (nikvaessen/w2v2-speaker) is pure Python: its hot path dispatches to
 * torch / cuDNN / cuBLAS through `torch.nn` modules, there is no FFI of its own.  Each entry
 * point below therefore replaces the torch op sequence of one reference call site (cited as
 * R:<file>:<line> = /root/reference/..., HF:<line> = transformers/models/wav2vec2/
 * modeling_wav2vec2.py 5.5.0, the third-party file that holds the encoder arithmetic the
 * reference calls at R:src/models/wav2vec2.py:71).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated;
 *   - the CALLER owns every buffer (kernels never allocate); `stream` is a cudaStream_t;
 *     all work is asynchronous on that stream;
 *   - return value 0 = ok, negative = error (message via w2v2_last_error()); no exceptions;
 *   - activations are channels-last: conv stack [B, L, C], transformer [B*T, H] row-major;
 *   - "f16" buffers hold IEEE binary16 (GEMM operands, RNE-rounded by the producer), "f32"
 *     buffers hold the residual stream / statistics / outputs.
 */
#ifndef W2V2_B200_H_
#define W2V2_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------ */
const char* w2v2_last_error(void);          /* thread-local, valid until the next failing call */
int w2v2_abi_version(void);                 /* bump on any signature change */
int w2v2_sm_count(void);                    /* SM count of the current device (grid sizing) */
int64_t w2v2_launch_count(void);            /* kernels launched by this library since it was loaded */

/* ---- dense contractions (tcgen05) ---------------------------------------------------------- */
/* out[b,r,n] = act(sum_{tap,c} A[b][r*a_row_stride + tap*a_tap_stride + c] * W[n][tap*cin + c] + bias[n])
 * A, W fp16; accumulation fp32; out fp16 (out_dtype 0) or fp32 (1).  All strides in ELEMENTS.
 * Replaces: Conv1d k=3/2,s=2 + GELU (HF:254-272; ntaps = kernel size, a_row_stride = 2*C,
 * a_tap_stride = C), nn.Linear (+GELU) of HF:429-434 / HF:524-549 / HF:566-573, the ASP 1x1
 * convs and the classifier GEMMs (R:src/lightning_modules/speaker/wav2vec2_fc.py:199-210,
 * R:src/optim/loss/aam_softmax.py:55).   act: 0 = none, 1 = exact-erf GELU.  bias may be NULL.
 * Requirements: cin % 64 == 0; row pitches (bytes) multiples of 16; bases 16-byte aligned. */
int w2v2_gemm_f16(const void* A, int64_t a_rows, int64_t a_row_stride, int64_t a_batch_stride, int batch,
                  int ntaps, int64_t a_tap_stride, int cin, const void* W, int64_t ldw, int N,
                  const float* bias, int act, void* out, int out_dtype, int64_t ldo, int64_t out_batch_stride,
                  void* stream);

/* ---- feature extractor layer 0 -------------------------------------------------------------- */
/* Conv1d(1->C,k=10,s=5,no bias) + GroupNorm(C groups == per-(b,c) instance norm over time, eps,
 * affine) + exact GELU  (HF:302-323).  wav f32 [B,N]; w f32 [C,10]; gamma/beta f32 [C];
 * out f16 channels-last [B,L0,C], L0 = (N-10)/5+1.  `workspace`: w2v2_conv0_workspace_bytes(B,N,C)
 * bytes, 256-byte aligned (GroupNorm scale/shift, fp64 window moments, im2col operand).
 * The GroupNorm statistics come from the window moments of the waveform (the conv is linear), so the
 * [B,C,L0] pre-norm tensor is never materialised; the conv itself runs on the tensor cores as a
 * K=64 GEMM over error-compensated fp16 windows with the GroupNorm affine + GELU in the epilogue. */
int64_t w2v2_conv0_workspace_bytes(int B, int N, int C);
int w2v2_conv0_gn_gelu(const float* wav, int B, int N, const float* w, const float* gamma, const float* beta,
                       float eps, void* workspace, void* out_f16, int C, void* stream);

/* ---- normalisation --------------------------------------------------------------------------- */
/* y = LayerNorm(x (+ bias) (+ residual)) * gamma + beta over the last dim (HF:429-431, 690-693,
 * 596-607).  x: f32 or f16 [rows, H] (x_dtype 1/0); bias f32 [H] or NULL; residual f32 [rows,H]
 * or NULL; outputs: y32 (f32, may be NULL) and y16 (f16, may be NULL). */
int w2v2_layernorm(const void* x, int x_dtype, const float* bias, const float* residual, const float* gamma,
                   const float* beta, float eps, float* y32, void* y16, int64_t rows, int H, void* stream);
/* Train mode: y = LayerNorm(dropout(x + bias) + residual), the hidden dropout of HF:546-549 / 572-574
 * fused in (mask = w2v2_dropout's counter-based mask over the flat [rows, H] index; drop_p = 0: none). */
int w2v2_layernorm_ex(const void* x, int x_dtype, const float* bias, const float* residual, const float* gamma,
                      const float* beta, float eps, float* y32, void* y16, int64_t rows, int H, float drop_p,
                      uint64_t drop_seed, void* stream);

/* Same, additionally writing the per-row 1 / sqrt(var + eps) (f32 [rows], NULL = skip): with it the backward can work
 * from the LayerNorm OUTPUT (w2v2_layernorm_bwd_from_output) instead of re-reading the two inputs. */
int w2v2_layernorm_ex2(const void* x, int x_dtype, const float* bias, const float* residual, const float* gamma,
                       const float* beta, float eps, float* y32, void* y16, float* rstd_out, int64_t rows, int H,
                       float drop_p, uint64_t drop_seed, void* stream);

/* ---- positional conv embedding --------------------------------------------------------------- */
/* Number of taps U that share one tensor-core A tile for sequences of T frames (0 = T too long
 * for the single-slab kernel).  The folded weight layout depends on it. */
int w2v2_posconv_taps_per_mma(int T, int H, int groups);
/* Fold weight_norm (HF:340-358): w[o,i,k] = g[k] * v[o,i,k] / ||v[:,:,k]||, and re-lay it out as
 * fp16 [G][K/U][I/8][U][O][8] (per (group, tap group): a [U*O x I] K-major block in UMMA core-matrix
 * order).  w16 must have room for H*(H/groups)*K halfs followed by K floats of scratch (per-tap norms). */
int w2v2_posconv_fold_weight(const float* v, const float* g, void* w16, int H, int groups, int K, int U, int mode,
                             void* stream);   /* mode 0: forward weight; 1: data-gradient weight (in/out swapped, taps reversed) */
/* y[b,t,:] = GELU(conv1d(x, w, bias, pad=K/2, groups)[.., :T]) (HF:360-379); x f16 [B,T,H];
 * out f32 [B,T,H] (the encoder then does LN(h + y), fused in w2v2_layernorm via `residual`). */
int w2v2_posconv(const void* x16, const void* w16, const float* bias, float* out, int B, int T, int H, int groups,
                 int K, void* stream);
/* General form: act = 0 leaves out the GELU (training keeps the pre-activation), bias may be NULL,
 * in_shift = s reads input frame t + s (zero beyond T).  With the mode-1 folded weight, act = 0,
 * bias = NULL, in_shift = 1 this is the data gradient of the positional conv. */
int w2v2_posconv_ex(const void* x16, const void* w16, const float* bias, float* out, int B, int T, int H, int groups,
                    int K, int act, int in_shift, void* stream);

/* ---- self-attention core ---------------------------------------------------------------------- */
/* o = softmax(q k^T) v per (batch, head), no mask (HF:438-463; the 1/sqrt(d) scale is folded
 * into the q projection).  qkv f16 [B*T, 3H] (q | k | v blocks); out f16 [B*T, H];
 * lse f32 [B, heads, T] = log-sum-exp of every score row (NULL in inference; saved for backward). */
int w2v2_attention(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, void* stream);

/* ---- pooling (R:src/layers/pooling.py) -------------------------------------------------------- */
/* mode 0: mean -> [B,H] (:24-30); mode 1: [std_unbiased || mean] -> [B,2H] (:38-44);
 * mode 2: max -> [B,H] (:74-80); mode 3: the ASP front statistics [mean || sqrt(clamp(var_population, 1e-12))] -> [B,2H].
 * x f32 [B,T,H]. */
int w2v2_stat_pool(const float* x, float* out, int B, int T, int H, int mode, void* stream);
/* attentive statistics pooling tail (speechbrain ASP as used at :87-106): given x f32 [B,T,H] and
 * attention logits a f32 [B,T,H]: softmax over T per (b,c), weighted mean and
 * std = sqrt(clamp(sum a (x-mean)^2, 1e-12)) -> out [B, 2H] = [mean || std]. */
int w2v2_asp_pool(const float* x, const float* logits, float* out, int B, int T, int H, void* stream);
/* ASP front: global mean/std over T (uniform weights, eps 1e-12) and the concatenated fp16
 * operand [x, mean, std] -> cat16 [B*T, 3H]. */
int w2v2_asp_concat(const float* x, void* cat16, int B, int T, int H, void* stream);
/* ASP middle: y = tanh(BatchNorm(ReLU(z))) with given per-channel scale/shift (eval: from running
 * stats); z f32 [rows, A] -> y16 f16 [rows, A]. */
int w2v2_asp_relu_bn_tanh(const float* z, const float* scale, const float* shift, void* y16, int64_t rows, int A,
                          void* stream);
/* Same with a per-utterance bias added to z first: ubias f32 [rows / rows_per_utt, A].  The evaluation path feeds the
 * TDNN's 1x1 conv with the frames only (K = H) and adds the product of its mean / std columns with the utterance's
 * statistics here, instead of materialising [x | mean | std] for every frame (K = 3H). */
int w2v2_asp_relu_bn_tanh_ubias(const float* z, const float* scale, const float* shift, const float* ubias, int rows_per_utt,
                                void* y16, int64_t rows, int A, void* stream);

/* ---- heads / losses --------------------------------------------------------------------------- */
/* Row-wise softmax + cross entropy + argmax over logits f32 [B, ldl] (first S valid):
 * prob f32 [B,S] (may be NULL), loss_rows f32 [B] (-log p[label]), argmax i32 [B].
 * (R:src/optim/loss/cross_entropy.py:19-33; mean over rows is done by w2v2_mean_rows.) */
int w2v2_softmax_ce(const float* logits, int64_t ldl, const int64_t* labels, float* prob, float* loss_rows,
                    int32_t* argmax, int B, int S, void* stream);
/* AAM-softmax margin (R:src/optim/loss/aam_softmax.py:56-69) applied in place on cosine f32
 * [B, ldl]: phi = cos*cos_m - sqrt(clamp(1-cos^2,0,1))*sin_m, phi = (cos-th>0) ? phi : cos-mm at
 * the label column, then * scale everywhere; followed by softmax/CE/argmax as above. */
int w2v2_aam_softmax_ce(float* cosine, int64_t ldl, const int64_t* labels, float margin, float scale,
                        int easy_margin, float* prob, float* loss_rows, int32_t* argmax, int B, int S, void* stream);
/* L2-normalise rows (F.normalize, eps 1e-12; :55) f32 [rows, E] -> f16 [rows, E]. */
int w2v2_l2norm_rows_f16(const float* x, void* y16, int64_t rows, int E, void* stream);
/* Error-compensated fp16 operands for the precision-critical classifier GEMMs: v = hi + lo with
 * hi = half(v), lo = half(v - hi); rows are emitted 3E wide as [hi|lo|hi] (which = 0, activation
 * side) or [hi|hi|lo] (which = 1, weight side) so that ONE w2v2_gemm_f16 call with cin = 3E computes
 * x_hi w_hi + x_lo w_hi + x_hi w_lo in the fp32 TMEM accumulator (~2^-20 relative instead of 2^-11). */
int w2v2_split3_rows(const float* x, void* y16, int64_t rows, int E, int which, void* stream);
int w2v2_l2norm_rows_split3(const float* x, void* y16, int64_t rows, int E, int which, void* stream);
/* mean of n floats -> out[0] (loss reduction 'mean'). */
int w2v2_mean_rows(const float* x, float* out, int n, void* stream);

/* ---- backward pass (autograd of the path above: R:src/lightning_modules/speaker/speaker_recognition_module.py:148-220
 *      calls loss.backward(); torch autograd then runs the backward of every op listed above) ---------------- */
/* dW[n,k] += sum_m dY[m,n] * X[m,k]   (weight gradient of out = X W^T).  dY f16 [M, ldy], X f16 [M, ldx],
 * dW f32 [N, ldw] (accumulated into: zero it first for a fresh gradient).  Both operands are consumed
 * MN-major straight from their row-major layout; the M range is split over CTAs and reduced with TMA
 * reduce-add stores. */
int w2v2_gemm_wgrad_f16(const void* dY, int64_t ldy, const void* X, int64_t ldx, int64_t M, int N, int K, float* dW,
                        int64_t ldw, void* stream);
/* w f32 [R, C] -> wT f16 [C, ldt] (transpose + cast, columns >= R zero-filled, optional per-row scale):
 * the B operand of the data-gradient GEMM dX = dY W, which is then a plain w2v2_gemm_f16 call. */
int w2v2_cast_f16_transpose(const float* w, void* wt16, int R, int C, int ldt, const float* row_scale, void* stream);
/* LayerNorm backward for y = LN(xa (+bias) (+residual)); dy = dy_a (+ dy_b).  x, mean, rstd are
 * recomputed.  dx32 / dx16 may be NULL; dgamma / dbeta (f32 [H], may be NULL) are accumulated. */
int w2v2_layernorm_bwd(const float* dy_a, const float* dy_b, const void* xa, int xa_dtype, const float* bias,
                       const float* residual, const float* gamma, float eps, float* dx32, void* dx16, float* dgamma,
                       float* dbeta, int64_t rows, int H, void* stream);
/* Backward of w2v2_layernorm_ex: the mask is regenerated from (drop_p, drop_seed); dx32 is the gradient
 * of the residual input, dx16 the gradient of the dropped branch input (dx * mask / (1 - p)); dbias (f32 [H]
 * or NULL) accumulates the column sums of that branch gradient = the gradient of `bias`. */
int w2v2_layernorm_bwd_ex(const float* dy_a, const float* dy_b, const void* xa, int xa_dtype, const float* bias,
                          const float* residual, const float* gamma, float eps, float* dx32, void* dx16, float* dgamma,
                          float* dbeta, float* dbias, int64_t rows, int H, float drop_p, uint64_t drop_seed,
                          void* stream);
/* LayerNorm backward from the OUTPUT y = LN(drop(xa + bias) + residual): xhat = (y - beta) / gamma and the row's
 * rstd (saved by w2v2_layernorm_ex2) replace the recomputation from xa / residual -- one [rows, H] fp32 stream
 * less and no statistics.  Outputs and the dropout of the branch gradient as in w2v2_layernorm_bwd_ex.
 * H in {512, 768, 1024}. */
int w2v2_layernorm_bwd_from_output(const float* dy_a, const float* dy_b, const float* y32, const float* rstd,
                                   const float* gamma, const float* beta, float* dx32, void* dx16, float* dgamma,
                                   float* dbeta, float* dbias, int64_t rows, int H, float drop_p, uint64_t drop_seed,
                                   void* stream);
/* dz = dg * gelu'(z), all f16, n % 8 == 0. */
int w2v2_gelu_bwd(const void* dg16, const void* z16, void* dz16, int64_t n, void* stream);
/* Same over [rows, cols], plus dbias[c] += sum_r dz[r, c] (gradient of the bias that was added to form z). */
int w2v2_gelu_bwd_colsum(const void* dg16, const void* z16, void* dz16, int64_t rows, int cols, float* dbias,
                         void* stream);
/* out[c] += scale * sum_r x[r, c]  (bias gradients); x f16 (x_dtype 0) or f32 (1), row pitch ld. */
int w2v2_colsum(const void* x, int x_dtype, int64_t rows, int cols, int64_t ld, float scale, float* out, void* stream);
/* dlogits = (prob - onehot(label)) * coef -> f16 [B, ldd] (columns >= S zero).  coef = loss_scale / B. */
int w2v2_softmax_ce_bwd(const float* prob, const int64_t* labels, float coef, void* dlogits16, int B, int S, int ldd,
                        void* stream);
/* mean pooling backward: dh[b,t,:] = demb[b,:] / T. */
int w2v2_mean_pool_bwd(const float* demb, float* dh, int B, int T, int H, void* stream);
/* y = GELU(x) as a standalone pass (training: the GEMM / posconv ran without the fused activation);
 * x, y f16 (dtype 0) or f32 (1); x16_copy (may be NULL) receives the rounded pre-activation. */
int w2v2_gelu_fwd(const void* x, int x_dtype, void* y, int y_dtype, void* x16_copy, int64_t n, void* stream);
/* Weight gradient of the positional conv without an im2col:
 *   dW[g*I + o][k][i] += sum_{b,t} dz[b,t,g*I+o] * x[b, t + k - K/2, g*I + i]      (I = H / groups, 48 or 64)
 * dz16, x16: f16 [B,T,H]; dw_hki f32 [H][K][I] is accumulated into (zero it for a fresh gradient), the
 * layout w2v2_weight_norm_bwd consumes.  Any T.  (autograd of HF:360-368 / cuDNN grouped backward-filter) */
int w2v2_posconv_wgrad(const void* dz16, const void* x16, float* dw_hki, int B, int T, int H, int groups, int K,
                       void* stream);
/* Positional-conv weight gradient: X_g[(b,t), j*I + i] = x[b, t+j-K/2, g*I+i] (f16 [B*T, K*I]) for one
 * group; the gradient of the folded weight is then w2v2_gemm_wgrad_f16(dz[:, g*O:(g+1)*O], X_g). */
int w2v2_posconv_im2col(const void* x16, void* xg16, int B, int T, int H, int groups, int K, int g, void* stream);
/* weight_norm backward (HF:340-358): dw f32 [H][K][I] (gradient of the folded weight, rows as produced by
 * the wgrad above) -> dv [H,I,K] += ..., dg [K] += ...; scratch: 2K floats; scale multiplies both. */
int w2v2_weight_norm_bwd(const float* dw_hki, const float* v, const float* g, float* scratch_2k, float scale, float* dv,
                         float* dg, int H, int I, int K, void* stream);
/* Attention backward (recomputes P from q, k and the saved log-sum-exp): qkv f16 [B*T, 3H], o f16 [B*T, H]
 * (forward output), d_o f16 [B*T, H], lse f32 [B, heads, T]  ->  dqkv f16 [B*T, 3H].  T <= 192. */
int w2v2_attention_bwd(const void* qkv16, const void* o16, const void* do16, const float* lse, void* dqkv16, int B, int T,
                       int H, int heads, void* stream);
/* mean+std pooling backward: dx from dout = [dstd || dmean] (recomputes mean / std). */
int w2v2_meanstd_pool_bwd(const float* x, const float* dout, float* dx, int B, int T, int H, void* stream);
/* AAM-softmax forward that also saves the pre-margin cosine of the label column (cos_label f32 [B]). */
int w2v2_aam_softmax_ce_ex(float* cosine, int64_t ldl, const int64_t* labels, float margin, float scale,
                           int easy_margin, float* prob, float* loss_rows, int32_t* argmax, float* cos_label, int B,
                           int S, void* stream);
/* dcos = scale * (prob - onehot) * coef * dloss[0] * (label column ? phi'(cos) : 1) -> f16 [B, ldd]. */
int w2v2_aam_bwd_dcos(const float* prob, const float* cos_label, const int64_t* labels, const float* dloss, float coef,
                      float margin, float scale, int easy_margin, void* dcos16, int B, int S, int ldd, void* stream);
/* inv[r] = 1 / max(||x_r||, 1e-12);  and the backward of row L2 normalisation:
 * dx (+)= scale * (dxh - xh (xh . dxh)) / ||x||   (dxh row pitch ldd). */
int w2v2_row_inv_norm(const float* x, float* inv, int64_t rows, int E, void* stream);
int w2v2_l2norm_rows_bwd(const float* x, const float* dxh, int64_t ldd, float* dx, int64_t rows, int E, float scale,
                         int accumulate, void* stream);
/* ---- train-mode regularisation (HF:433, 456, 546-600, 693, 1280-1324) -------------------------------
 * Dropout masks are a pure function of (seed, element index) -- regenerated in the backward, never stored:
 * keep <=> 16 random bits >= round(p * 65536); bits = a 32-bit multiply / xor-shift mix of (index/2) and the
 * two seed halves (csrc/common.cuh::dropout_hash; numpy replica in tests/test_gpu_regularise.py), low / high
 * half for the even / odd element.
 * w2v2_dropout: y = keep ? (x + bias[col]) / (1-p) : 0 over n elements (x, y f32 (dtype 1) or f16 (0),
 * y may alias x; optional extra f16 copy y16).  The attention kernels take the probability-dropout the
 * same way (index = ((b*heads + h)*T + q) * TK + k).  Time mask: rows with mask != 0 are overwritten by
 * the learned embedding; backward zeroes those gradient rows and accumulates them into d_embed. */
int w2v2_dropout(const void* x, int dtype, const float* bias, int H, void* y, void* y16, int64_t n, float p,
                 uint64_t seed, void* stream);
int w2v2_attention_ex(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, float drop_p,
                      uint64_t drop_seed, void* stream);
int w2v2_attention_bwd_ex(const void* qkv16, const void* o16, const void* do16, const float* lse, void* dqkv16, int B,
                          int T, int H, int heads, float drop_p, uint64_t drop_seed, void* stream);
/* Same, with the chain rule of the folded softmax scale and the bias gradients done on the way out of TMEM:
 * dq is multiplied by qscale before it is stored, and dbias (f32 [3H], NULL = skip) accumulates the column
 * sums of (dq * qscale | dk | dv) -- the q/k/v projection bias gradients. */
int w2v2_attention_bwd_ex2(const void* qkv16, const void* o16, const void* do16, const float* lse, void* dqkv16, int B,
                           int T, int H, int heads, float drop_p, uint64_t drop_seed, float qscale, float* dbias,
                           void* stream);
int w2v2_time_mask_apply(float* h, const uint8_t* mask, const float* embed, int64_t rows, int H, void* stream);
/* SpecAugment along the feature axis (HF:1312-1322, mask_feature_prob): h[b, t, c] = 0 where mask[b * H + c] != 0, in place
 * (h f32 [B, T, H], H % 4 == 0, mask 4-byte aligned); the same call masks the gradient in the backward. */
int w2v2_feature_mask(float* h, const uint8_t* mask, int B, int T, int H, void* stream);
int w2v2_time_mask_bwd(float* dh, const uint8_t* mask, float* dembed, int64_t rows, int H, float scale, void* stream);
/* Gradient plumbing: out = a + b (b may be NULL) as f32 and/or f16 (n % 4 == 0); row-wise f32 -> f16 cast
 * with zero padding to ldy and a scale; in-place scale; f32 CE gradient (prob - onehot) * coef * dloss[0]
 * (dloss may be NULL = 1). */
int w2v2_add2_cast(const float* a, const float* b, float* out32, void* out16, int64_t n, void* stream);
int w2v2_cast_f16_rows(const float* x, int64_t ldx, void* y16, int64_t ldy, int64_t rows, int cols, float scale,
                       void* stream);
int w2v2_scale_f32(float* x, int64_t n, float s, void* stream);
/* y = x * s, out of place (the loss scale entering / leaving an autograd Function: gradients that cross a Function
 * boundary are plain unscaled fp32 and belong to autograd). */
int w2v2_scale_copy_f32(const float* x, float* y, int64_t n, float s, void* stream);
/* Gradient entry of an autograd Function with a device-chosen normalisation (an outer torch.amp.GradScaler --
 * `precision: 16`, R:config/experiment/speaker_wav2vec2_aam.yaml:17 -- multiplies the incoming gradient by up to 2^16, which
 * would overflow the fp16 operand copies of the backward GEMMs): y = x * s * k, where k = 1 while amax|x| * s lies in
 * [lo, hi] (the normal case: bit-identical to w2v2_scale_copy_f32) and otherwise the power of two that brings it to `mid`;
 * state f32[2] (device): [0] scratch, [1] = 1 / k.  No host synchronisation.  w2v2_scale_f32_dev: x *= s * dev_scale[0]. */
/* out32[M, N] += A[M, K] (f16) W[N, K]^T (f16): the data-gradient GEMMs of the backward add straight into the gradient of
 * the residual stream (every tile is written with TMA reduce-add, partial tiles of the stream-K schedule included), so the
 * LayerNorm backward that follows reads ONE fp32 gradient stream instead of two (HF:596-607 residual adds, backward). */
/* ---- ragged evaluation batches (the reference embeds test utterances one at a time,
 * R:src/lightning_modules/speaker/speaker_recognition_module.py:462-500, R:src/predict.py:132-170) --------------------
 * Utterances of different length are zero-padded to a common [B, N] and carried with their lengths so that every
 * utterance gets exactly what a batch of one would give it: GroupNorm statistics of conv layer 0 over its own frames
 * (lens = samples), zeros behind its end for the positional conv (w2v2_cast_f16_rowmask, lens = frames), attention
 * over its own keys (w2v2_attention_lens; every query row is still computed so that padding rows stay finite), pooling
 * over its own frames.  lens: int32 device vectors of B entries. */
int w2v2_conv0_gn_lens(const float* wav, int B, int N, const int* lens, const float* w, const float* gamma, const float* beta,
                       float eps, void* workspace, void* out_f16, int C, int act, void* stream);
/* conv layer 0 + GroupNorm (+ GELU) straight from the RAW waveform: wav is 16-bit PCM (in_dtype 0,
 * x = pcm / 32768) or float32 (1); normalize = 1 folds the reference's input normaliser, x' = (x - mean) / (std + 1e-5)
 * per utterance (R:src/data/preprocess/input_normalisation.py:53-67), into the GroupNorm affine -- no normalised copy
 * of the waveform is ever written, the host uploads half the bytes and runs no arithmetic.  lens as above (may be NULL). */
int w2v2_conv0_raw(const void* wav, int in_dtype, int normalize, int B, int N, const int* lens, const float* w,
                   const float* gamma, const float* beta, float eps, void* workspace, void* out_f16, int C, int act,
                   void* stream);
int w2v2_cast_f16_rowmask(const float* x, void* y16, int B, int T, int H, const int* lens, void* stream);
int w2v2_attention_lens(const void* qkv16, void* out16, int B, int T, int H, int heads, const int* lens, void* stream);
int w2v2_stat_pool_lens(const float* x, float* out, int B, int T, int H, int mode, const int* lens, void* stream);
int w2v2_asp_concat_split3_lens(const float* x, void* cat16x3, int B, int T, int H, const int* lens, void* stream);
int w2v2_asp_concat_lens(const float* x, void* cat16, int B, int T, int H, const int* lens, void* stream);
int w2v2_asp_pool_lens(const float* x, const float* logits, float* out, int B, int T, int H, const int* lens, void* stream);
/* Sum all-reduce of multicast_base[lo, hi) (fp32) over `world` ranks through the NVSwitch: two-shot, rank r pulls the
 * switch-reduced sum of its 1/world slice (multimem.ld_reduce) and broadcasts it (multimem.st).  multicast_base: this
 * rank's address of the multicast mapping of a symmetric buffer (the same offsets on every rank); lo / hi multiples of
 * 4 elements.  The caller orders the ranks with barriers on `stream` before (every replica written) and after (every
 * slice broadcast).  Replaces the ncclAllReduce DDP issues for the gradient buckets (R:config/trainer/trainer.yaml:6-9). */
int w2v2_nvls_allreduce_f32(void* multicast_base, int64_t lo, int64_t hi, int rank, int world, int ctas, void* stream);
int w2v2_dgrad_accumulates(void);   /* 1: w2v2_encoder_layer_bwd leaves the whole input gradient in dx1_32 (dh_in32 unused) */
int w2v2_gemm_f16_accum(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N, float* out32,
                        int64_t ldo, void* stream);
int w2v2_grad_entry_scale(const float* x, float* y, int64_t n, float s, float lo, float hi, float mid, float* state,
                          void* stream);
int w2v2_scale_f32_dev(float* x, int64_t n, float s, const float* dev_scale, void* stream);
int w2v2_softmax_ce_bwd_f32(const float* prob, const int64_t* labels, const float* dloss, float coef, float* dlogits,
                            int B, int S, void* stream);
/* torch.optim.Adam step (weight_decay 0) over flat fp32 buffers; `g` is multiplied by grad_scale first
 * (undoes the loss scale and applies the 1/world_size of the data-parallel mean). */
int w2v2_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   int step, float grad_scale, void* stream);
/* Same, optionally clearing the gradient in the same pass (zero_grad != 0): the next backward re-accumulates it. */
int w2v2_adam_step_ex(float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                      int step, float grad_scale, int zero_grad, void* stream);

/* ---- attentive-statistics pooling in training (R:src/layers/pooling.py:87-106; speechbrain
 * AttentiveStatisticsPooling, restated in oracle/w2v2_oracle.py::attentive_stat_pool) -------------------
 * w2v2_asp_bn_batch_stats: BatchNorm1d(A) in training mode over the rows of relu(z) [rows, A]: writes the
 *   affine the forward kernel w2v2_asp_relu_bn_tanh applies (scale = gamma * rstd, shift = beta - mean *
 *   scale), the batch mean / rstd (kept for the backward) and updates running_mean / running_var (NULL =
 *   skip) with torch semantics (momentum, unbiased variance).  sums_ws: 2*A doubles of workspace.
 * w2v2_asp_pool_bwd: from d out [B, 2C] = (d mean || d std): d logits (f16 [B*T, C]) of the softmax over
 *   time and the direct part of d x (f32 [B,T,C], overwritten).  `out` is the saved forward output.
 * w2v2_asp_act_bwd: d h [rows, A] -> d z (f16) through tanh, BatchNorm (batch_stats = 1: training-mode
 *   statistics; 0: running statistics) and ReLU; d gamma / d beta (x grad_scale; NULL = skip).
 * w2v2_asp_front_bwd: d x += d cat[:, :C] + the gradient through the uniform mean / std broadcast in
 *   cat[:, C:3C] (d cat f32, row pitch ldc). */
/* cat = [x | mean | std] (uniform statistics) as error-compensated fp16 operand [hi | lo | hi], f16 [B*T, 9H]:
 * z = cat W1^T is then one GEMM against w2v2_split3_rows(W1, 1) = [hi | hi | lo]; cat16x3[:, :3H] is the plain
 * fp16 cat. */
int w2v2_asp_concat_split3(const float* x, void* cat16x3, int B, int T, int H, void* stream);
int w2v2_asp_bn_batch_stats(const float* z, int64_t rows, int A, const float* gamma, const float* beta, float eps,
                            float momentum, float* running_mean, float* running_var, double* sums_ws, float* scale,
                            float* shift, float* mean, float* rstd, void* stream);
int w2v2_asp_pool_bwd(const float* x, const float* logits, const float* out, const float* dout, void* dlogits16, float* dx,
                      int B, int T, int H, void* stream);
int w2v2_asp_act_bwd(const float* dh, const float* z, const float* scale, const float* shift, const float* mean,
                     const float* rstd, int batch_stats, double* sums_ws, void* dz16, float* dgamma, float* dbeta,
                     float grad_scale, int64_t rows, int A, void* stream);
int w2v2_asp_front_bwd(const float* x, const float* dcat, int64_t ldc, float* dx, int B, int T, int H, void* stream);

/* ---- native launch schedules of one encoder layer (HF:592-609) ------------------------------------------
 * One call issues every kernel of a transformer layer (forward: 8-10 launches, backward: ~17) on `stream`
 * into caller-owned buffers; they exist because 300 interpreter -> C round trips per training step cost as
 * much host time as the step takes on the device.  All pointers are device pointers; f16 buffers are
 * row-major with the natural pitch ([B*T, H], [B*T, 3H], [B*T, FF]).  Dropout seeds are derived from `seed`
 * and `layer` exactly as the python schedule did (attention seed+100+l, LN1 branch +200+l, LN2 branch
 * +300+l, activation +400+l), so forward and backward regenerate the same masks.
 * Forward: z16 == NULL selects inference (GELU fused into the FFN1 epilogue, lse may be NULL). */
typedef struct {
  int B, T, H, heads, FF, layer;
  float eps, p_hidden, p_attn, p_act;
  uint64_t seed;
  const void* wqkv;   /* f16 [3H, H], q rows pre-scaled by d^-0.5 */
  const float* bqkv;  /* f32 [3H] */
  const void* wo;     /* f16 [H, H] */
  const float* bo;
  const float* ln1_g;
  const float* ln1_b;
  const void* w1;     /* f16 [FF, H] */
  const float* b1;
  const void* w2;     /* f16 [H, FF] */
  const float* b2;
  const float* ln2_g;
  const float* ln2_b;
  const float* h_in32; /* f32 [B*T, H] residual stream */
  const void* h_in16;  /* its f16 copy */
  void* qkv16;
  void* att16;
  float* lse;          /* f32 [B, heads, T] */
  float* o32;
  float* h1_32;
  void* h1_16;
  void* z16;           /* f16 [B*T, FF] pre-activation (training) or NULL */
  void* g16;           /* f16 [B*T, FF] gelu(z) */
  float* f2_32;
  float* h2_32;        /* layer output */
  void* h2_16;
  float* rstd1;        /* f32 [B*T] 1/sigma of LayerNorm 1 / 2 (training; NULL in inference) */
  float* rstd2;
  const int* key_lens; /* inference of a ragged batch: int32 [B] frames per utterance (NULL = all T frames are real) */
} w2v2_layer_fwd_args;
int w2v2_encoder_layer_fwd(const w2v2_layer_fwd_args* args, void* stream);

/* Backward of the same layer.  dy_a + dy_b (dy_b may be NULL) is the gradient of the layer output; on return
 * dh_in32 + dx1_32 is the gradient of the layer input (kept as two terms: the next LayerNorm backward sums
 * them for free).  Weight / bias / LayerNorm gradients are ACCUMULATED into the d_* buffers (fp32, natural
 * shapes; d_wqkv / d_bqkv span q|k|v), activation gradients carry the caller's loss scale.  The remaining
 * pointers are scratch of the sizes their names imply. */
typedef struct {
  int B, T, H, heads, FF, layer;
  float eps, p_hidden, p_attn, p_act, qscale;
  uint64_t seed;
  const void* wqkvT;  /* f16 [H, 3H]; the q block is the UNSCALED Wq^T (dq arrives multiplied by qscale) */
  const void* woT;    /* f16 [H, H] */
  const void* w1T;    /* f16 [H, FF] */
  const void* w2T;    /* f16 [FF, H] */
  const float* bo;
  const float* b2;
  const float* ln1_g;
  const float* ln2_g;
  const float* h_in32;
  const void* h_in16;
  const void* qkv16;
  const void* att16;
  const float* lse;
  const float* o32;
  const float* h1_32;
  const void* h1_16;
  const void* z16;
  const void* g16;
  const float* f2_32;
  const float* dy_a;
  const float* dy_b;
  float* d_wqkv;
  float* d_bqkv;
  float* d_wo;
  float* d_bo;
  float* d_ln1_g;
  float* d_ln1_b;
  float* d_w1;
  float* d_b1;
  float* d_w2;
  float* d_b2;
  float* d_ln2_g;
  float* d_ln2_b;
  float* dx2_32;      /* scratch f32 [B*T, H] */
  void* dx2_16;       /* scratch f16 [B*T, H] */
  void* dg16;         /* scratch f16 [B*T, FF] */
  void* dz16;         /* scratch f16 [B*T, FF] */
  float* dh1_32;      /* scratch f32 [B*T, H] */
  void* dx1_16;       /* scratch f16 [B*T, H] */
  void* datt16;       /* scratch f16 [B*T, H] */
  void* dqkv16;       /* scratch f16 [B*T, 3H] */
  float* dx1_32;      /* out: residual-path term of the input gradient */
  float* dh_in32;     /* out: term through the q/k/v projections */
  const float* h2_32; /* the layer's output (LayerNorm 2 output) and the saved rstd of both LayerNorms: */
  const float* rstd1; /* the LayerNorm backward works from the outputs (h1_32, h2_32) */
  const float* rstd2;
  const float* ln1_b;
  const float* ln2_b;
} w2v2_layer_bwd_args;
int w2v2_encoder_layer_bwd(const w2v2_layer_bwd_args* args, void* stream);

/* Optional timing of every tensor-core GEMM launch (tap-GEMM and wgrad) between start and stop: CUDA events
 * on the launching stream; stop synchronises the device and returns summed milliseconds, FLOPs (from the
 * launch arguments) and the number of launches.  Measurement aid of bench.py's roofline figure. */
int w2v2_gemm_profile_start(void);
int w2v2_gemm_profile_stop(double* total_ms, double* total_flops, int* launches);

/* ---- backward of the CNN feature extractor (HF:254-323; the reference trains it when
 * completely_freeze_feature_extractor is false) ------------------------------------------------------------
 * w2v2_conv0_gn_ex: w2v2_conv0_gn_gelu with the GELU optional (act 0: the GroupNorm output, which the
 *   training forward keeps as the pre-activation); w2v2_conv0_workspace_offsets: where the per-(b, c) affine
 *   (scale = gamma * rstd, shift) and the error-compensated im2col operand ([B*L0, 64] f16, columns
 *   [x_hi(10) | x_lo(10) | x_hi(10) | 0]) live inside that workspace -- the backward reads both.
 * w2v2_gemm_f16_taps: tap-GEMM whose taps are ROW offsets into one activation (may be negative; rows outside
 *   [0, a_extent) read as zero): out[b, r, n] = sum_t sum_c A[b, r + tap_row[t], c] W[n, t*cin + c].  The data
 *   gradient of a stride-2 Conv1d is two of these (even / odd input rows), written through ldo = 2 * Cin.
 * w2v2_gemm_wgrad_f16_batched: dW[n, k] += sum_b sum_{r<rows} dY[b][r, n] X[b][r, k] with independent row
 *   pitches / batch strides (X = every stride-th input frame of one tap).
 * w2v2_groupnorm_bwd: per-(b, c) GroupNorm backward over time of conv layer 0 (see csrc/backward_cnn.cu). */
int w2v2_conv0_gn_ex(const float* wav, int B, int N, const float* w, const float* gamma, const float* beta, float eps,
                     void* workspace, void* out_f16, int C, int act, void* stream);
int w2v2_conv0_workspace_offsets(int B, int N, int C, int64_t* scale_off, int64_t* shift_off, int64_t* im2col_off);
/* out_act = gelu(A W^T + bias), out_pre = A W^T + bias (both f16 [M, N], row pitch ldo) from ONE GEMM: the FFN1 of
 * the training forward keeps the pre-activation for the backward. */
int w2v2_gemm_f16_dual_gelu(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N,
                            const float* bias, void* out_act16, void* out_pre16, int64_t ldo, void* stream);
/* dz = (A W^T) * gelu'(z) (f16 [M, N], row pitch ldo) and dbias[n] += sum_r dz[r, n] from ONE GEMM: the FFN2 data
 * gradient of the training backward (HF:560-573 reversed) with the GELU backward and the FFN1 bias gradient in its
 * epilogue.  z: the pre-activation w2v2_gemm_f16_dual_gelu kept, f16 [M, ldz]. */
int w2v2_gemm_f16_gelu_bwd(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N,
                           const void* z16, int64_t ldz, void* dz16, int64_t ldo, float* dbias, void* stream);
/* EXPERIMENTAL (compiled, not yet measured on a GPU; off unless W2V2_SAVE_GELU_GRAD=1, see csrc/schedule.cu): the same
 * pair of GEMMs with gelu'(z) kept instead of z, so that the backward epilogue is a plain multiply.
 * w2v2_gemm_f16_dual_gelu_grad: out_act = gelu(A W^T + bias), out_grad = gelu'(A W^T + bias).
 * w2v2_gemm_f16_mul_colsum:     out = (A W^T) * mul, colsum[n] += sum_r out[r, n]. */
int w2v2_gemm_f16_dual_gelu_grad(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N,
                                 const float* bias, void* out_act16, void* out_grad16, int64_t ldo, void* stream);
int w2v2_gemm_f16_mul_colsum(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N,
                             const void* mul16, int64_t ldm, void* out16, int64_t ldo, float* colsum, void* stream);
int w2v2_gemm_f16_taps(const void* A, int64_t out_rows, int64_t a_extent, const int* tap_row, int64_t a_row_stride,
                       int64_t a_batch_stride, int batch, int ntaps, int cin, const void* W, int64_t ldw, int N, void* out,
                       int out_dtype, int64_t ldo, int64_t out_batch_stride, void* stream);
int w2v2_gemm_wgrad_f16_batched(const void* dY, int64_t ldy, int64_t dy_batch_stride, const void* X, int64_t ldx,
                                int64_t x_batch_stride, int64_t rows, int batch, int N, int K, float* dW, int64_t ldw,
                                void* stream);
int w2v2_groupnorm_bwd(const void* dy16, const void* y16, const float* gamma, const float* beta, const float* scale,
                       void* dc16, float* dgamma, float* dbeta, float grad_scale, int B, int L, int C, void* stream);

/* ---- evaluation: trial scoring (R:src/evaluation/speaker/cosine_distance.py:107-132, 249-262) -----------
 * scores[p] = cosine(f(emb[idx_a[p]]), f(emb[idx_b[p]])), f(x) = (x - mean) / (std + 1e-12) when mean / std are
 * given (the evaluator's centring), identity otherwise; torch.nn.CosineSimilarity semantics (eps 1e-8). */
int w2v2_cosine_pairs(const float* emb, const float* mean, const float* stdv, const int32_t* idx_a, const int32_t* idx_b,
                      float* scores, int64_t P, int E, void* stream);

/* ---- in front of the path: waveform standardisation (R:src/data/preprocess/input_normalisation.py:53-67) ----
 * out[b, :] = (x[b, :] - mean_b) / (std_b + 1e-5), std unbiased; x = in (float32, in_dtype 1) or in / 32768
 * (int16 PCM, in_dtype 0).  mean / std (f32 [B]) may be NULL. */
int w2v2_normalize_wav(const void* in, int in_dtype, float* out, float* mean, float* stdv, int B, int N, void* stream);

/* ---- utility ---------------------------------------------------------------------------------- */
/* f32 -> f16 (RNE) with optional scale: y = half(x * scale). */
int w2v2_cast_f16(const float* x, void* y16, int64_t n, float scale, void* stream);
/* Batched weight preparation after an optimizer step (one launch for all trainable matrices):
 * per job, v = src[r, c] * scale is written to any of  dst16[r*ld + c] (f16),  dst32[r*ld + c] (f32),
 * dstT16[c*ldt + r] (f16, transposed; point it at a column offset to assemble fused W^T operands).
 * The table lives in device memory; tile_begin is the exclusive prefix sum of ceil(R/E)*ceil(C/E), E =
 * w2v2_prepare_tile_edge() (host-only query: 64, or 32 with W2V2_PREP_V3=0).
 * (replaces the per-step .half() / .t() / torch.cat the reference's AMP autocast does implicitly) */
typedef struct {
  const void* src;      /* f32 [R, C] row-major */
  void* dst16;          /* or NULL */
  void* dstT16;         /* or NULL */
  void* dst32;          /* or NULL */
  int32_t R, C;
  int32_t ld, ldt;      /* row pitch (elements) of dst16 / dst32, and of dstT16 */
  float scale;
  float scale_t;       /* scale of the transposed copy relative to src (the plain copies use `scale`) */
  int64_t tile_begin;
} w2v2_prep_job;
int w2v2_prepare_weights(const w2v2_prep_job* jobs_dev, int njobs, int64_t total_tiles, void* stream);
int w2v2_prepare_tile_edge(void);
/* Conv1d weight [Cout, Cin, K] f32 -> tap-major fp16 [Cout, K, Cin] (the W operand of w2v2_gemm_f16). */
int w2v2_conv_weight_tapmajor(const float* w, void* w16, int cout, int cin, int k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* W2V2_B200_H_ */
