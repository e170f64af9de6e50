// Backward of the pooling / AAM-softmax head pieces that are not GEMMs
// (R:src/layers/pooling.py:38-44, R:src/optim/loss/aam_softmax.py:50-74).
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int device_sm_count();

constexpr int HP_TS = 8, HP_CH = 32;

__device__ __forceinline__ float slices_sum(float v, float (*sm)[HP_CH], int ts, int ch) {
  __syncthreads();
  sm[ts][ch] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < HP_TS; ++i) s += sm[i][ch];
  return s;
}

// mean+std pooling backward: out = [std_unbiased || mean];  dx = dmean / T + dstd * (x - mean) / ((T-1) std)
__global__ void __launch_bounds__(HP_TS* HP_CH) meanstd_pool_bwd_kernel(const float* __restrict__ x,
                                                                        const float* __restrict__ dout,
                                                                        float* __restrict__ dx, int T, int H) {
  __shared__ float sm[HP_TS][HP_CH];
  const int ch = threadIdx.x % HP_CH, ts = threadIdx.x / HP_CH;
  const int c = blockIdx.x * HP_CH + ch;
  const int b = blockIdx.y;
  const float* xb = x + int64_t(b) * T * H + c;
  float s = 0.f;
  for (int t = ts; t < T; t += HP_TS) s += xb[int64_t(t) * H];
  const float mean = slices_sum(s, sm, ts, ch) / float(T);
  float q = 0.f;
  for (int t = ts; t < T; t += HP_TS) {
    const float d = xb[int64_t(t) * H] - mean;
    q = fmaf(d, d, q);
  }
  q = slices_sum(q, sm, ts, ch);
  const float stdv = sqrtf(q / float(T - 1));
  const float dstd = dout[int64_t(b) * 2 * H + c], dmean = dout[int64_t(b) * 2 * H + H + c];
  const float a = dmean / float(T);
  const float k = stdv > 0.f ? dstd / (float(T - 1) * stdv) : 0.f;
  float* db = dx + int64_t(b) * T * H + c;
  for (int t = ts; t < T; t += HP_TS) db[int64_t(t) * H] = fmaf(k, xb[int64_t(t) * H] - mean, a);
}

// dcos (f16 operand, [B, ldd]) of the AAM-softmax loss:
//   logit = scale * (col == label ? phi(c) : c),  dlogit = (prob - onehot) * coef
//   phi'(c) = cos_m + sin_m * c / sqrt(1 - c^2)  where phi is active, 1 where the fallback (c - mm) is
__global__ void aam_bwd_dcos_kernel(const float* __restrict__ prob, const float* __restrict__ cos_label,
                                    const int64_t* __restrict__ labels, const float* __restrict__ dloss, float coef,
                                    float cos_m, float sin_m, float th, float scale, int easy,
                                    __half* __restrict__ dc, int S, int ldd) {
  const int b = blockIdx.x;
  const int label = int(labels[b]);
  const float k = coef * scale * (dloss != nullptr ? dloss[0] : 1.0f);
  const float c = cos_label[b];
  float dphi = 1.0f;
  const bool active = easy ? (c > 0.f) : ((c - th) > 0.f);
  if (active) {
    const float s2 = fminf(fmaxf(1.0f - c * c, 0.f), 1.f);
    dphi = s2 > 0.f ? cos_m + sin_m * c * rsqrtf(s2) : cos_m;
  } else if (easy) {
    dphi = 1.0f;
  }
  for (int i = threadIdx.x; i < ldd; i += blockDim.x) {
    float v = 0.f;
    if (i < S) v = (prob[int64_t(b) * S + i] - (i == label ? 1.0f : 0.0f)) * k * (i == label ? dphi : 1.0f);
    dc[int64_t(b) * ldd + i] = __float2half_rn(v);
  }
}

__device__ __forceinline__ float block_sum256(float v, float* sm) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < int(blockDim.x >> 5); ++i) s += sm[i];
  return s;
}

// inv_norm[r] = 1 / max(||x_r||, 1e-12)
__global__ void __launch_bounds__(256) row_inv_norm_kernel(const float* __restrict__ x, float* __restrict__ inv, int E) {
  __shared__ float sm[8];
  const int64_t r = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) s = fmaf(x[r * E + i], x[r * E + i], s);
  s = block_sum256(s, sm);
  if (threadIdx.x == 0) inv[r] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
}

// backward of xh = x / ||x||:  dx (+)= scale * (dxh - xh (xh . dxh)) / ||x||
__global__ void __launch_bounds__(256) l2norm_rows_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dxh,
                                                              int64_t ldd, float* __restrict__ dx, int E, float scale,
                                                              int accumulate) {
  __shared__ float sm[8];
  const int64_t r = blockIdx.x;
  const float* xr = x + r * E;
  const float* gr = dxh + r * ldd;
  float s = 0.f, d = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    s = fmaf(xr[i], xr[i], s);
    d = fmaf(xr[i], gr[i], d);
  }
  s = block_sum256(s, sm);
  d = block_sum256(d, sm);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  const float proj = d * inv * inv;                      // (xh . dxh) / ||x||
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    const float v = scale * inv * (gr[i] - xr[i] * proj);
    dx[r * E + i] = accumulate ? dx[r * E + i] + v : v;
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" {

int w2v2_meanstd_pool_bwd(const float* x, const float* dout, float* dx, int B, int T, int H, void* stream) {
  W2V2_REQUIRE(H % HP_CH == 0 && T >= 2, "w2v2_meanstd_pool_bwd: need H %% 32 == 0 and T >= 2");
  dim3 g(H / HP_CH, B);
  meanstd_pool_bwd_kernel<<<g, HP_TS * HP_CH, 0, (cudaStream_t)stream>>>(x, dout, dx, T, H);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_aam_bwd_dcos(const float* prob, const float* cos_label, const int64_t* labels, const float* dloss, float coef,
                      float margin, float scale, int easy_margin, void* dcos16, int B, int S, int ldd, void* stream) {
  W2V2_REQUIRE(ldd >= S, "w2v2_aam_bwd_dcos: ldd < S");
  const double m = margin, pi = 3.14159265358979323846;
  aam_bwd_dcos_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(prob, cos_label, labels, dloss, coef, float(cos(m)), float(sin(m)),
                                                          float(cos(pi - m)), scale, easy_margin, (__half*)dcos16, S, ldd);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_row_inv_norm(const float* x, float* inv, int64_t rows, int E, void* stream) {
  if (rows == 0) return 0;
  row_inv_norm_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, inv, E);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_l2norm_rows_bwd(const float* x, const float* dxh, int64_t ldd, float* dx, int64_t rows, int E, float scale,
                         int accumulate, void* stream) {
  if (rows == 0) return 0;
  l2norm_rows_bwd_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, dxh, ldd, dx, E, scale, accumulate);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
