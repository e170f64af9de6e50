// Native launch schedules of one transformer encoder layer (HF:592-609): the ~10 (forward) / ~20 (backward)
// kernel launches of a layer issued from ONE C call, all on the caller's stream, into caller-owned buffers.
// The host side of the training step was the bottleneck (300 python -> ctypes round trips of ~35 us each
// against ~11 ms of device time); these entry points are the same launches without the interpreter in
// between.  Nothing here computes: every line is one of the C-ABI kernels of this library.
//
//   forward :  qkv = h Wqkv^T + b ; att = softmax(q k^T) v ; o = att Wo^T ; h1 = LN1(drop(o + bo) + h)
//              z = h1 W1^T + b1 ; g = gelu(z) ; f2 = g W2^T ; h2 = LN2(drop(f2 + b2) + h1)
//   backward:  the reverse, with the weight gradients accumulated into the caller's flat fp32 buffer.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "w2v2_b200.h"

#define W2V2_TRY(expr)     \
  do {                     \
    const int _rc = (expr); \
    if (_rc != 0) return _rc; \
  } while (0)

// The FFN2 data-gradient GEMM applies the GELU backward in its epilogue; W2V2_FUSE_GELU_BWD=0 restores the two-pass
// form (A/B measurements: 11.35 -> 11.22 ms per train step)
static bool fuse_gelu_bwd() {
  static const bool on = []() { const char* e = getenv("W2V2_FUSE_GELU_BWD"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}

// W2V2_SAVE_GELU_GRAD (default on since round 2; =0 restores the z-keeping form; GEMM time per step 6.41 -> 6.17 ms): the training forward keeps
// gelu'(z) in the z buffer instead of z, and the FFN2 data-gradient epilogue only multiplies by it.  Needs the fused
// backward epilogue and no activation dropout (the two-pass fallback reads z itself); both directions take the same
// decision from the same inputs.
static bool save_gelu_grad(float p_act) {
  static const bool on = []() { const char* e = getenv("W2V2_SAVE_GELU_GRAD"); return !(e != nullptr && e[0] == '0'); }();
  return on && fuse_gelu_bwd() && !(p_act > 0.f);
}

// W2V2_DGRAD_ACCUM (default on; =0 restores the two-term form): the FFN1 and QKV data-gradient GEMMs add straight into the
// residual-path gradient (w2v2_gemm_f16_accum: TMA reduce-add stores), so each LayerNorm backward reads one fp32 gradient
// stream instead of two -- the read moves from an HBM-bound kernel into a tensor-bound one.  The layer's input gradient is
// then dx1_32 alone (dh_in32 stays untouched); w2v2_dgrad_accumulates() tells the caller which form is active.
static bool dgrad_accum() {
  static const bool on = []() { const char* e = getenv("W2V2_DGRAD_ACCUM"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}
extern "C" int w2v2_dgrad_accumulates(void) { return dgrad_accum() ? 1 : 0; }

extern "C" int w2v2_encoder_layer_fwd(const w2v2_layer_fwd_args* a, void* stream) {
  W2V2_REQUIRE(a != nullptr, "w2v2_encoder_layer_fwd: null argument block");
  const int64_t M = int64_t(a->B) * a->T;
  const int H = a->H, FF = a->FF;
  const uint64_t seed = a->seed;
  const int l = a->layer;
  // attention block
  W2V2_TRY(w2v2_gemm_f16(a->h_in16, M, H, 0, 1, 1, 0, H, a->wqkv, H, 3 * H, a->bqkv, 0, a->qkv16, 0, 3 * H, 0, stream));
  if (a->key_lens != nullptr) {   // ragged evaluation batch: keys of each utterance's own frames only
    W2V2_REQUIRE(a->z16 == nullptr && a->lse == nullptr, "w2v2_encoder_layer_fwd: key_lens is an inference-only option");
    W2V2_TRY(w2v2_attention_lens(a->qkv16, a->att16, a->B, a->T, H, a->heads, a->key_lens, stream));
  } else {
    W2V2_TRY(w2v2_attention_ex(a->qkv16, a->att16, a->lse, a->B, a->T, H, a->heads, a->p_attn, seed + 100 + l, stream));
  }
  W2V2_TRY(w2v2_gemm_f16(a->att16, M, H, 0, 1, 1, 0, H, a->wo, H, H, nullptr, 0, a->o32, 1, H, 0, stream));
  W2V2_TRY(w2v2_layernorm_ex2(a->o32, 1, a->bo, a->h_in32, a->ln1_g, a->ln1_b, a->eps, a->h1_32, a->h1_16, a->rstd1, M, H,
                              a->p_hidden, seed + 200 + l, stream));
  // feed-forward block
  if (a->z16 != nullptr) {      // training: keep the pre-activation, GELU as its own pass
    // one GEMM, two outputs: z (kept for the backward) and g = gelu(z)
    if (save_gelu_grad(a->p_act))
      W2V2_TRY(w2v2_gemm_f16_dual_gelu_grad(a->h1_16, M, H, H, a->w1, H, FF, a->b1, a->g16, a->z16, FF, stream));
    else
      W2V2_TRY(w2v2_gemm_f16_dual_gelu(a->h1_16, M, H, H, a->w1, H, FF, a->b1, a->g16, a->z16, FF, stream));
    if (a->p_act > 0.f)
      W2V2_TRY(w2v2_dropout(a->g16, 0, nullptr, FF, a->g16, nullptr, M * FF, a->p_act, seed + 400 + l, stream));
  } else {                      // inference: GELU in the GEMM epilogue
    W2V2_TRY(w2v2_gemm_f16(a->h1_16, M, H, 0, 1, 1, 0, H, a->w1, H, FF, a->b1, 1, a->g16, 0, FF, 0, stream));
  }
  W2V2_TRY(w2v2_gemm_f16(a->g16, M, FF, 0, 1, 1, 0, FF, a->w2, FF, H, nullptr, 0, a->f2_32, 1, H, 0, stream));
  W2V2_TRY(w2v2_layernorm_ex2(a->f2_32, 1, a->b2, a->h1_32, a->ln2_g, a->ln2_b, a->eps, a->h2_32, a->h2_16, a->rstd2, M, H,
                              a->p_hidden, seed + 300 + l, stream));
  return 0;
}

extern "C" int w2v2_encoder_layer_bwd(const w2v2_layer_bwd_args* a, void* stream) {
  W2V2_REQUIRE(a != nullptr, "w2v2_encoder_layer_bwd: null argument block");
  const int64_t M = int64_t(a->B) * a->T;
  const int H = a->H, FF = a->FF;
  const uint64_t seed = a->seed;
  const int l = a->layer;
  // LN2:  h2 = LN(drop(f2 + b2) + h1); dx2_16 is the gradient of the dropped branch, the residual keeps dx2_32
  // (the LayerNorm backward works from the LayerNorm OUTPUT + the saved rstd: one fp32 stream less than from its inputs)
  W2V2_TRY(w2v2_layernorm_bwd_from_output(a->dy_a, a->dy_b, a->h2_32, a->rstd2, a->ln2_g, a->ln2_b, a->dx2_32, a->dx2_16,
                                          a->d_ln2_g, a->d_ln2_b, a->d_b2, M, H, a->p_hidden, seed + 300 + l, stream));
  W2V2_TRY(w2v2_gemm_wgrad_f16(a->dx2_16, H, a->g16, FF, M, H, FF, a->d_w2, FF, stream));
  if (a->p_act > 0.f || !fuse_gelu_bwd()) {
    W2V2_TRY(w2v2_gemm_f16(a->dx2_16, M, H, 0, 1, 1, 0, H, a->w2T, H, FF, nullptr, 0, a->dg16, 0, FF, 0, stream));
    if (a->p_act > 0.f)
      W2V2_TRY(w2v2_dropout(a->dg16, 0, nullptr, FF, a->dg16, nullptr, M * FF, a->p_act, seed + 400 + l, stream));
    W2V2_TRY(w2v2_gelu_bwd_colsum(a->dg16, a->z16, a->dz16, M, FF, a->d_b1, stream));
  } else {
    // no activation dropout (the reference's default): dz = (dx2 W2) * gelu'(z) and d_b1 straight from the GEMM epilogue
    if (save_gelu_grad(a->p_act))
      W2V2_TRY(w2v2_gemm_f16_mul_colsum(a->dx2_16, M, H, H, a->w2T, H, FF, a->z16, FF, a->dz16, FF, a->d_b1, stream));
    else
      W2V2_TRY(w2v2_gemm_f16_gelu_bwd(a->dx2_16, M, H, H, a->w2T, H, FF, a->z16, FF, a->dz16, FF, a->d_b1, stream));
  }
  W2V2_TRY(w2v2_gemm_wgrad_f16(a->dz16, FF, a->h1_16, H, M, FF, H, a->d_w1, H, stream));
  const bool accum = dgrad_accum();
  if (accum) W2V2_TRY(w2v2_gemm_f16_accum(a->dz16, M, FF, FF, a->w1T, FF, H, a->dx2_32, H, stream));
  else W2V2_TRY(w2v2_gemm_f16(a->dz16, M, FF, 0, 1, 1, 0, FF, a->w1T, FF, H, nullptr, 0, a->dh1_32, 1, H, 0, stream));
  // LN1:  h1 = LN(drop(o + bo) + h_in)
  W2V2_TRY(w2v2_layernorm_bwd_from_output(accum ? a->dx2_32 : a->dh1_32, accum ? nullptr : a->dx2_32, a->h1_32, a->rstd1, a->ln1_g, a->ln1_b, a->dx1_32, a->dx1_16,
                                          a->d_ln1_g, a->d_ln1_b, a->d_bo, M, H, a->p_hidden, seed + 200 + l, stream));
  W2V2_TRY(w2v2_gemm_wgrad_f16(a->dx1_16, H, a->att16, H, M, H, H, a->d_wo, H, stream));
  W2V2_TRY(w2v2_gemm_f16(a->dx1_16, M, H, 0, 1, 1, 0, H, a->woT, H, H, nullptr, 0, a->datt16, 0, H, 0, stream));
  // The q projection was used pre-scaled by d^-0.5: the attention backward multiplies dq by the same factor (chain
  // rule for the unscaled parameters) and emits the q/k/v bias gradients while it drains TMEM; wqkvT holds the
  // UNSCALED Wq^T, so neither the weight gradient nor the data gradient needs a correction pass.
  W2V2_TRY(w2v2_attention_bwd_ex2(a->qkv16, a->att16, a->datt16, a->lse, a->dqkv16, a->B, a->T, H, a->heads, a->p_attn,
                                  seed + 100 + l, a->qscale, a->d_bqkv, stream));
  W2V2_TRY(w2v2_gemm_wgrad_f16(a->dqkv16, 3 * H, a->h_in16, H, M, 3 * H, H, a->d_wqkv, H, stream));
  if (accum) W2V2_TRY(w2v2_gemm_f16_accum(a->dqkv16, M, 3 * H, 3 * H, a->wqkvT, 3 * H, H, a->dx1_32, H, stream));
  else W2V2_TRY(w2v2_gemm_f16(a->dqkv16, M, 3 * H, 0, 1, 1, 0, 3 * H, a->wqkvT, 3 * H, H, nullptr, 0, a->dh_in32, 1, H, 0, stream));
  return 0;
}
