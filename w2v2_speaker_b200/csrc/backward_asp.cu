// Attentive-statistics pooling in training: batch-statistics BatchNorm of the attention TDNN and the
// backward of every piece (R:src/layers/pooling.py:87-106 -> speechbrain AttentiveStatisticsPooling,
// restated in oracle/w2v2_oracle.py::attentive_stat_pool).  With x [B,T,C]:
//     mu, sigma = uniform stats of x;  cat = [x | mu | sigma]            (asp_concat, elementwise.cu)
//     z = cat W1^T + b1;  r = relu(z);  y = BN(r);  h = tanh(y)           (GEMM; asp_bn_*; asp_relu_bn_tanh)
//     l = h W2^T + b2;  alpha = softmax_t(l)                              (GEMM; asp_pool)
//     m = sum_t alpha x;  s = sqrt(clamp(sum_t alpha (x - m)^2, 1e-12));  out = [m | s]
// Backward (all activation gradients carry the caller's loss scale, they are linear in d out):
//     d alpha_t = dm x_t + dv (x_t - m)^2,  dv = ds / (2 s);   sum_t alpha_t d alpha_t = dm m + dv s^2
//     d l_t = alpha_t (d alpha_t - dm m - dv s^2);   d x_t (direct) = alpha_t (dm + 2 dv (x_t - m))
//     d h = d l W2;  d y = d h (1 - h^2);  BN backward over the B*T rows;  d z = d r [z > 0]
//     d cat = d z W1;  d x_t += d cat_t[:C] + d mu / T + d sigma (x_t - mu) / (T sigma)
// The GEMMs are the tcgen05 kernels (gemm_tc.cu / gemm_wgrad.cu); this file holds the rest.
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int device_sm_count();

constexpr int AP_TS = 8;     // time slices per block
constexpr int AP_CH = 32;    // channels per block

__device__ __forceinline__ float ap_sum(float v, float (*sm)[AP_CH], int ts, int ch) {
  __syncthreads();
  sm[ts][ch] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < AP_TS; ++i) s += sm[i][ch];
  return s;
}
__device__ __forceinline__ float ap_max(float v, float (*sm)[AP_CH], int ts, int ch) {
  __syncthreads();
  sm[ts][ch] = v;
  __syncthreads();
  float s = sm[0][ch];
#pragma unroll
  for (int i = 1; i < AP_TS; ++i) s = fmaxf(s, sm[i][ch]);
  return s;
}

// ---- ASP front with error-compensated operands: cat = [x | mu | sigma] written as [hi | lo | hi] (fp16, row
// pitch 9H) so that z = cat W1^T can run as ONE GEMM over [hi|lo|hi] . [hi|hi|lo] with ~2^-20 relative error.
// A plain fp16 cat changes z by ~1e-3, which flips ReLU derivatives of the units near zero and costs a few
// percent of gradient accuracy in the TDNN (measured against the fp32 oracle); the hi block alone is the fp16
// cat the weight-gradient GEMM consumes.
__global__ void __launch_bounds__(AP_TS* AP_CH) asp_concat_split3_kernel(const float* __restrict__ x, __half* __restrict__ cat,
                                                                          int T, int H, const int* __restrict__ lens) {
  __shared__ float sm[AP_TS][AP_CH];
  const int ch = threadIdx.x % AP_CH, ts = threadIdx.x / AP_CH;
  const int c = blockIdx.x * AP_CH + ch;
  const int b = blockIdx.y;
  const float* xb = x + int64_t(b) * T * H + c;
  const int Tv = lens != nullptr ? lens[b] : T;      // ragged batch: statistics over the utterance's own frames
  const float w = 1.0f / float(Tv);
  float s = 0.f;
  for (int t = ts; t < Tv; t += AP_TS) s = fmaf(w, xb[int64_t(t) * H], s);
  const float mean = ap_sum(s, sm, ts, ch);
  float q = 0.f;
  for (int t = ts; t < Tv; t += AP_TS) {
    const float d = xb[int64_t(t) * H] - mean;
    q = fmaf(w * d, d, q);
  }
  const float stdv = sqrtf(fmaxf(ap_sum(q, sm, ts, ch), 1e-12f));
  const __half mh = __float2half_rn(mean), sh = __float2half_rn(stdv);
  const __half ml = __float2half_rn(mean - __half2float(mh)), sl = __float2half_rn(stdv - __half2float(sh));
  const int64_t ld = 9 * int64_t(H);
  __half* cb = cat + int64_t(b) * T * ld;
  for (int t = ts; t < T; t += AP_TS) {
    __half* row = cb + int64_t(t) * ld;
    const float v = xb[int64_t(t) * H];
    const __half vh = __float2half_rn(v);
    const __half vl = __float2half_rn(v - __half2float(vh));
    row[c] = vh;             row[H + c] = mh;             row[2 * H + c] = sh;            // hi
    row[3 * H + c] = vl;     row[4 * H + c] = ml;         row[5 * H + c] = sl;            // lo
    row[6 * H + c] = vh;     row[7 * H + c] = mh;         row[8 * H + c] = sh;            // hi
  }
}

// ---- BatchNorm1d(A) over the rows of relu(z): column sums of r and r^2 (double atomics: E[r^2] - E[r]^2)
__global__ void __launch_bounds__(256) asp_bn_stats_kernel(const float* __restrict__ z, double* __restrict__ sums, int64_t rows,
                                                           int A) {
  const int col = threadIdx.x % A, slice = threadIdx.x / A, nslice = blockDim.x / A;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t r = int64_t(blockIdx.x) * nslice + slice; r < rows; r += int64_t(gridDim.x) * nslice) {
    const float v = fmaxf(z[r * A + col], 0.f);
    s1 += v;
    s2 += double(v) * v;
  }
  atomicAdd(sums + col, s1);
  atomicAdd(sums + A + col, s2);
}

// mean / rstd of the batch, the affine (scale, shift) the forward kernel applies, running-stat update
// (torch semantics: momentum on the mean and the UNBIASED variance)
__global__ void asp_bn_finalize_kernel(const double* __restrict__ sums, int64_t rows, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, float momentum,
                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                       float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                       float* __restrict__ rstd_out, int A) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A) return;
  const double mean = sums[a] / double(rows);
  double var = sums[A + a] / double(rows) - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = float(1.0 / sqrt(var + double(eps)));
  const float sc = gamma[a] * rstd;
  scale[a] = sc;
  shift[a] = beta[a] - float(mean) * sc;
  mean_out[a] = float(mean);
  rstd_out[a] = rstd;
  if (running_mean != nullptr) {
    const double unbiased = rows > 1 ? var * double(rows) / double(rows - 1) : var;
    running_mean[a] = (1.f - momentum) * running_mean[a] + momentum * float(mean);
    running_var[a] = (1.f - momentum) * running_var[a] + momentum * float(unbiased);
  }
}

// ---- pooling tail backward: per (b, c) column over time
__global__ void __launch_bounds__(AP_TS* AP_CH) asp_pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ lg,
                                                                     const float* __restrict__ out, const float* __restrict__ dout,
                                                                     __half* __restrict__ dlg16, float* __restrict__ dx, int T,
                                                                     int H) {
  __shared__ float sm[AP_TS][AP_CH];
  const int ch = threadIdx.x % AP_CH, ts = threadIdx.x / AP_CH;
  const int c = blockIdx.x * AP_CH + ch;
  const int b = blockIdx.y;
  const int64_t base = int64_t(b) * T * H + c;
  const float* xb = x + base;
  const float* lb = lg + base;
  float mx = -INFINITY;
  for (int t = ts; t < T; t += AP_TS) mx = fmaxf(mx, lb[int64_t(t) * H]);
  mx = ap_max(mx, sm, ts, ch);
  float se = 0.f;
  for (int t = ts; t < T; t += AP_TS) se += expf(lb[int64_t(t) * H] - mx);
  se = ap_sum(se, sm, ts, ch);
  const float inv = 1.0f / se;
  const float m = out[int64_t(b) * 2 * H + c], s = out[int64_t(b) * 2 * H + H + c];
  const float dm = dout[int64_t(b) * 2 * H + c], ds = dout[int64_t(b) * 2 * H + H + c];
  const float dv = s > 1.0001e-6f ? ds / (2.f * s) : 0.f;           // clamp(., 1e-12) active: no gradient
  const float dot = dm * m + dv * s * s;                            // sum_t alpha_t d alpha_t
  for (int t = ts; t < T; t += AP_TS) {
    const float a = expf(lb[int64_t(t) * H] - mx) * inv;
    const float d = xb[int64_t(t) * H] - m;
    const float da = fmaf(dm, xb[int64_t(t) * H], dv * d * d);
    dlg16[base + int64_t(t) * H] = __float2half_rn(a * (da - dot));
    dx[base + int64_t(t) * H] = a * fmaf(2.f * dv, d, dm);
  }
}

// ---- tanh / BatchNorm / ReLU backward.  Pass 1: column sums of dy and dy * rhat.
__global__ void __launch_bounds__(256) asp_act_bwd_stats_kernel(const float* __restrict__ dh, const float* __restrict__ z,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                double* __restrict__ sums, int64_t rows, int A) {
  const int col = threadIdx.x % A, slice = threadIdx.x / A, nslice = blockDim.x / A;
  const float sc = scale[col], sh = shift[col], mu = mean[col], rs = rstd[col];
  double s1 = 0.0, s2 = 0.0;
  for (int64_t r = int64_t(blockIdx.x) * nslice + slice; r < rows; r += int64_t(gridDim.x) * nslice) {
    const float rl = fmaxf(z[r * A + col], 0.f);
    const float h = tanhf(fmaf(rl, sc, sh));
    const float dy = dh[r * A + col] * (1.f - h * h);
    s1 += dy;
    s2 += double(dy) * ((rl - mu) * rs);
  }
  atomicAdd(sums + col, s1);
  atomicAdd(sums + A + col, s2);
}

// Pass 2: dz (fp16 operand of the next GEMMs).  batch_stats = 1: training-mode BatchNorm (the statistics
// depend on the batch); 0: running statistics (a plain affine).  Also emits d gamma / d beta (x grad_scale).
__global__ void __launch_bounds__(256) asp_act_bwd_apply_kernel(const float* __restrict__ dh, const float* __restrict__ z,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                const double* __restrict__ sums, int batch_stats,
                                                                __half* __restrict__ dz16, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta, float grad_scale, int64_t rows, int A) {
  const int64_t n = rows * A;
  const double inv_rows = 1.0 / double(rows);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int a = int(i % A);
    const float zz = z[i];
    const float rl = fmaxf(zz, 0.f);
    const float sc = __ldg(scale + a);
    const float h = tanhf(fmaf(rl, sc, __ldg(shift + a)));
    const float dy = dh[i] * (1.f - h * h);
    float dr = dy;
    if (batch_stats) {
      const float rhat = (rl - __ldg(mean + a)) * __ldg(rstd + a);
      dr = dy - float(sums[a] * inv_rows) - rhat * float(sums[A + a] * inv_rows);
    }
    dz16[i] = __float2half_rn(zz > 0.f ? dr * sc : 0.f);
  }
  if (blockIdx.x == 0 && int(threadIdx.x) < A) {
    if (dgamma != nullptr) dgamma[threadIdx.x] = float(sums[A + threadIdx.x]) * grad_scale;
    if (dbeta != nullptr) dbeta[threadIdx.x] = float(sums[threadIdx.x]) * grad_scale;
  }
}

// ---- front backward: dx = dx_direct + dcat[:, :C] + d mu / T + d sigma (x - mu) / (T sigma)
__global__ void __launch_bounds__(AP_TS* AP_CH) asp_front_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dcat,
                                                                      int64_t ldc, float* __restrict__ dx, int T, int H) {
  __shared__ float sm[AP_TS][AP_CH];
  const int ch = threadIdx.x % AP_CH, ts = threadIdx.x / AP_CH;
  const int c = blockIdx.x * AP_CH + ch;
  const int b = blockIdx.y;
  const float* xb = x + int64_t(b) * T * H + c;
  const float* db = dcat + int64_t(b) * T * ldc + c;
  float* dxb = dx + int64_t(b) * T * H + c;
  const float w = 1.0f / float(T);
  float s = 0.f, dmu = 0.f, dsg = 0.f;
  for (int t = ts; t < T; t += AP_TS) {
    s = fmaf(w, xb[int64_t(t) * H], s);
    dmu += db[int64_t(t) * ldc + H];
    dsg += db[int64_t(t) * ldc + 2 * H];
  }
  const float mean = ap_sum(s, sm, ts, ch);
  dmu = ap_sum(dmu, sm, ts, ch);
  dsg = ap_sum(dsg, sm, ts, ch);
  float q = 0.f;
  for (int t = ts; t < T; t += AP_TS) {
    const float d = xb[int64_t(t) * H] - mean;
    q = fmaf(w * d, d, q);
  }
  q = ap_sum(q, sm, ts, ch);
  const float k_sigma = q > 1e-12f ? dsg * w / sqrtf(q) : 0.f;
  for (int t = ts; t < T; t += AP_TS) {
    const float d = xb[int64_t(t) * H] - mean;
    dxb[int64_t(t) * H] += db[int64_t(t) * ldc] + dmu * w + k_sigma * d;
  }
}

static int ew_grid(int64_t n, int per_block) {
  int64_t g = (n + per_block - 1) / per_block;
  const int64_t cap = int64_t(device_sm_count()) * 8;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_asp_bn_batch_stats(const float* z, int64_t rows, int A, const float* gamma, const float* beta, float eps,
                                       float momentum, float* running_mean, float* running_var, double* sums_ws,
                                       float* scale, float* shift, float* mean, float* rstd, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  W2V2_REQUIRE(A >= 1 && A <= 256 && 256 % A == 0, "w2v2_asp_bn_batch_stats: attention channels %d must divide 256", A);
  W2V2_REQUIRE(rows >= 1, "w2v2_asp_bn_batch_stats: empty batch");
  W2V2_CHECK_CUDA(cudaMemsetAsync(sums_ws, 0, sizeof(double) * 2 * A, st));
  asp_bn_stats_kernel<<<ew_grid(rows, 64 * (256 / A)), 256, 0, st>>>(z, sums_ws, rows, A);
  asp_bn_finalize_kernel<<<(A + 127) / 128, 128, 0, st>>>(sums_ws, rows, gamma, beta, eps, momentum, running_mean,
                                                          running_var, scale, shift, mean, rstd, A);
  count_launches(2);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_asp_concat_split3(const float* x, void* cat16x3, int B, int T, int H, void* stream) {
  return w2v2_asp_concat_split3_lens(x, cat16x3, B, T, H, nullptr, stream);
}
extern "C" int w2v2_asp_concat_split3_lens(const float* x, void* cat16x3, int B, int T, int H, const int* lens, void* stream) {
  W2V2_REQUIRE(H % AP_CH == 0, "w2v2_asp_concat_split3: H=%d must be a multiple of %d", H, AP_CH);
  dim3 g(H / AP_CH, B);
  asp_concat_split3_kernel<<<g, AP_TS * AP_CH, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__half*>(cat16x3), T, H, lens);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_asp_pool_bwd(const float* x, const float* logits, const float* out, const float* dout, void* dlogits16,
                                 float* dx, int B, int T, int H, void* stream) {
  W2V2_REQUIRE(H % AP_CH == 0, "w2v2_asp_pool_bwd: H=%d must be a multiple of %d", H, AP_CH);
  dim3 g(H / AP_CH, B);
  asp_pool_bwd_kernel<<<g, AP_TS * AP_CH, 0, static_cast<cudaStream_t>(stream)>>>(x, logits, out, dout,
                                                                                  static_cast<__half*>(dlogits16), dx, T, H);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_asp_act_bwd(const float* dh, const float* z, const float* scale, const float* shift, const float* mean,
                                const float* rstd, int batch_stats, double* sums_ws, void* dz16, float* dgamma, float* dbeta,
                                float grad_scale, int64_t rows, int A, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  W2V2_REQUIRE(A >= 1 && A <= 256 && 256 % A == 0, "w2v2_asp_act_bwd: attention channels %d must divide 256", A);
  W2V2_CHECK_CUDA(cudaMemsetAsync(sums_ws, 0, sizeof(double) * 2 * A, st));
  asp_act_bwd_stats_kernel<<<ew_grid(rows, 64 * (256 / A)), 256, 0, st>>>(dh, z, scale, shift, mean, rstd, sums_ws, rows, A);
  asp_act_bwd_apply_kernel<<<ew_grid(rows * A, 1024), 256, 0, st>>>(dh, z, scale, shift, mean, rstd, sums_ws, batch_stats,
                                                                   static_cast<__half*>(dz16), dgamma, dbeta, grad_scale,
                                                                   rows, A);
  count_launches(2);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_asp_front_bwd(const float* x, const float* dcat, int64_t ldc, float* dx, int B, int T, int H,
                                  void* stream) {
  W2V2_REQUIRE(H % AP_CH == 0, "w2v2_asp_front_bwd: H=%d must be a multiple of %d", H, AP_CH);
  dim3 g(H / AP_CH, B);
  asp_front_bwd_kernel<<<g, AP_TS * AP_CH, 0, static_cast<cudaStream_t>(stream)>>>(x, dcat, ldc, dx, T, H);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}
