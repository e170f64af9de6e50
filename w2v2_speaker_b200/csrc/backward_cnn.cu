// Backward of conv layer 0 of the feature extractor: Conv1d(1 -> C, k=10, s=5) -> GroupNorm(C groups) -> GELU
// (HF:302-323), needed when the reference trains with the CNN unfrozen
// (completely_freeze_feature_extractor: false, R:config/network/wav2vec2_fc.yaml:16).
//
// The training forward keeps y = GroupNorm(conv) (pre-GELU, f16 channels-last [B, L, C]) and the per-(b, c)
// affine scale = gamma * rstd it applied.  With chat = (y - beta) / gamma the normalised activation:
//     dy    = d out * gelu'(y)                                   (w2v2_gelu_bwd)
//     dgamma[c] += sum_{b,t} dy chat,   dbeta[c] += sum_{b,t} dy
//     dc    = rstd (g - mean_t(g) - chat mean_t(g chat)),  g = dy gamma          (per (b, c) over time)
//     dW0[c, j] = sum_{b,t} dc[b,t,c] wav[b, 5 t + j]            (tcgen05 wgrad against the im2col operand)
// Layers 1..6 need no kernels of their own: their data gradients are tap-GEMMs with per-tap row offsets
// (w2v2_gemm_f16_taps), their weight gradients batched wgrad GEMMs (w2v2_gemm_wgrad_f16_batched).
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

constexpr int GB_TS = 8;      // time slices per block
constexpr int GB_CH = 32;     // channels per block

__device__ __forceinline__ float gb_sum(float v, float (*sm)[GB_CH], int ts, int ch) {
  __syncthreads();
  sm[ts][ch] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < GB_TS; ++i) s += sm[i][ch];
  return s;
}

// one block per (32 channels, batch element); threads (channel, time slice): coalesced along channels
__global__ void __launch_bounds__(GB_TS* GB_CH) groupnorm_bwd_kernel(const __half* __restrict__ dy, const __half* __restrict__ y,
                                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                      const float* __restrict__ scale, __half* __restrict__ dc,
                                                                      float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                      float grad_scale, int L, int C) {
  __shared__ float sm[GB_TS][GB_CH];
  const int ch = threadIdx.x % GB_CH, ts = threadIdx.x / GB_CH;
  const int c = blockIdx.x * GB_CH + ch;
  const int b = blockIdx.y;
  const int64_t base = int64_t(b) * L * C + c;
  const float g = gamma[c], bt = beta[c];
  const float inv_g = fabsf(g) > 1e-20f ? 1.0f / g : 0.f;
  const float rstd = scale[int64_t(b) * C + c] * inv_g;             // scale = gamma * rstd
  float s1 = 0.f, s2 = 0.f;
  for (int t = ts; t < L; t += GB_TS) {
    const float d = __half2float(dy[base + int64_t(t) * C]);
    const float chat = (__half2float(y[base + int64_t(t) * C]) - bt) * inv_g;
    s1 += d;
    s2 = fmaf(d, chat, s2);
  }
  s1 = gb_sum(s1, sm, ts, ch);
  s2 = gb_sum(s2, sm, ts, ch);
  if (ts == 0) {
    atomicAdd(dgamma + c, s2 * grad_scale);
    atomicAdd(dbeta + c, s1 * grad_scale);
  }
  const float m1 = s1 / float(L), m2 = s2 / float(L);
  const float k = rstd * g;
  for (int t = ts; t < L; t += GB_TS) {
    const float d = __half2float(dy[base + int64_t(t) * C]);
    const float chat = (__half2float(y[base + int64_t(t) * C]) - bt) * inv_g;
    dc[base + int64_t(t) * C] = __float2half_rn(k * (d - m1 - chat * m2));
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_groupnorm_bwd(const void* dy16, const void* y16, const float* gamma, const float* beta, const float* scale,
                                  void* dc16, float* dgamma, float* dbeta, float grad_scale, int B, int L, int C, void* stream) {
  W2V2_REQUIRE(C % GB_CH == 0, "w2v2_groupnorm_bwd: C=%d must be a multiple of %d", C, GB_CH);
  W2V2_REQUIRE(B >= 1 && L >= 1, "w2v2_groupnorm_bwd: empty problem");
  dim3 g(C / GB_CH, B);
  groupnorm_bwd_kernel<<<g, GB_TS * GB_CH, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(dy16), static_cast<const __half*>(y16), gamma, beta, scale, static_cast<__half*>(dc16), dgamma,
      dbeta, grad_scale, L, C);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}
