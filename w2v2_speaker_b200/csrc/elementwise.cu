// HBM-bound kernels of the hot path: conv layer 0 + GroupNorm + GELU, LayerNorm (+bias+residual),
// pooling, ASP glue, softmax/CE/AAM heads, casts.  Warp-shuffle reductions, 16-byte vector
// accesses, channels-last layouts.  See include/w2v2_b200.h for the contracts.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int device_sm_count();
int gemm_f16_impl(const void* A, int64_t a_rows, int64_t a_row_stride, int64_t a_batch_stride, int batch, int ntaps,
                  int64_t a_tap_stride, int cin, const void* W, int64_t ldw, int N, const float* bias,
                  const float* shift, int act, void* out, int out_dtype, int64_t ldo, int64_t out_batch_stride,
                  cudaStream_t stream);
static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

// =================================================================================================
// casts / weight re-layout

__global__ void cast_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, int64_t n, float scale) {
  int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + i + 4));
    uint4 q;
    q.x = pack_half2(a.x * scale, a.y * scale);
    q.y = pack_half2(a.z * scale, a.w * scale);
    q.z = pack_half2(b.x * scale, b.y * scale);
    q.w = pack_half2(b.z * scale, b.w * scale);
    *reinterpret_cast<uint4*>(y + i) = q;
  }
  if (i < n) {  // tail (at most one thread lands here per tail element group)
    for (int64_t j = i; j < n && j < i + 8; ++j) y[j] = __float2half_rn(x[j] * scale);
  }
}

// y16[b, t, :] = t < lens[b] ? x[b, t, :] : 0 -- the positional conv of a padded ragged batch must see zeros behind the
// end of each utterance (its own zero padding in a batch of one, HF:326-379)
__global__ void cast_f16_rowmask_kernel(const float* __restrict__ x, __half* __restrict__ y, int64_t n4, int T, int H4,
                                        const int* __restrict__ lens) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / H4;
    const int b = int(r / T), t = int(r - int64_t(b) * T);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < lens[b]) v = reinterpret_cast<const float4*>(x)[i];
    uint2 q;
    q.x = pack_half2(v.x, v.y);
    q.y = pack_half2(v.z, v.w);
    reinterpret_cast<uint2*>(y)[i] = q;
  }
}

__global__ void conv_weight_tapmajor_kernel(const float* __restrict__ w, __half* __restrict__ o, int cout, int cin,
                                            int k) {
  // o[co][j][ci] = w[co][ci][j]
  const int64_t n = int64_t(cout) * cin * k;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int ci = i % cin;
    const int j = (i / cin) % k;
    const int co = i / (int64_t(cin) * k);
    o[i] = __float2half_rn(w[(int64_t(co) * cin + ci) * k + j]);
  }
}

// =================================================================================================
// conv layer 0 + GroupNorm + GELU   (HF:302-323)
//
// GroupNorm with C groups over C channels = per-(b,c) standardisation over time.  Because the conv is
// linear, its per-channel mean/variance over time follow EXACTLY from the first and second moments
// of the ten stride-5 sample windows of the waveform:
//     mean_c = sum_k w[c,k] m_k,     var_c = sum_{k,k'} w[c,k] w[c,k'] (S_kk'/L - m_k m_k')
// so pass 1 only reads the waveform (65 fp64 moments per utterance) and the [B,C,L] pre-norm
// activation is never materialised; pass 2 recomputes the conv, normalises, applies GELU and
// writes fp16 channels-last with 16-byte stores.

constexpr int C0_K = 10, C0_S = 5, C0_NMOM = 10 + 55;
constexpr int C0_MOMS = C0_NMOM + 2;      // per-utterance record: 65 window moments, sum(x), sum(x^2)

// lens (optional): samples of utterance b in a zero-padded ragged batch -- its GroupNorm statistics cover its own
// L_b = (lens[b] - 10) / 5 + 1 frames only, exactly what a batch of one would see.
// PCM16: the waveform is raw 16-bit PCM (x = pcm / 32768, what torchaudio.load returns for a 16-bit file).
// SUMS: additionally accumulate sum(x) and sum(x^2) over the utterance's samples (mom[65], mom[66]): the input
// normaliser (R:src/data/preprocess/input_normalisation.py:53-67) is then folded into the GroupNorm affine below.
template <bool PCM16>
__device__ __forceinline__ float c0_load(const void* wav, int64_t i) {
  if constexpr (PCM16) return float(__ldg(static_cast<const int16_t*>(wav) + i)) * (1.0f / 32768.0f);
  else return __ldg(static_cast<const float*>(wav) + i);
}
template <bool PCM16, bool SUMS>
__global__ void conv0_moments_kernel(const void* __restrict__ wav, int N, int L, double* __restrict__ mom,
                                     const int* __restrict__ lens) {
  const int b = blockIdx.y;
  const int n_b = lens != nullptr ? lens[b] : N;
  if (lens != nullptr) L = (n_b - C0_K) / C0_S + 1;
  const int64_t x0 = int64_t(b) * N;
  double acc[C0_NMOM];
#pragma unroll
  for (int i = 0; i < C0_NMOM; ++i) acc[i] = 0.0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L; t += gridDim.x * blockDim.x) {
    float v[C0_K];
#pragma unroll
    for (int k = 0; k < C0_K; ++k) v[k] = c0_load<PCM16>(wav, x0 + t * C0_S + k);
    int idx = C0_K;
#pragma unroll
    for (int k = 0; k < C0_K; ++k) {
      acc[k] += double(v[k]);
#pragma unroll
      for (int k2 = k; k2 < C0_K; ++k2) acc[idx++] += double(v[k]) * double(v[k2]);
    }
  }
  double s1 = 0.0, s2 = 0.0;
  if constexpr (SUMS) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_b; i += gridDim.x * blockDim.x) {
      const double v = double(c0_load<PCM16>(wav, x0 + i));
      s1 += v;
      s2 += v * v;
    }
  }
  __shared__ double red[8][C0_NMOM + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < C0_NMOM + (SUMS ? 2 : 0); ++i) {
    double a = i < C0_NMOM ? acc[i < C0_NMOM ? i : 0] : (i == C0_NMOM ? s1 : s2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[warp][i] = a;
  }
  __syncthreads();
  if (threadIdx.x < C0_NMOM + (SUMS ? 2 : 0)) {
    double a = 0.0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) a += red[w][threadIdx.x];
    atomicAdd(mom + int64_t(b) * C0_MOMS + threadIdx.x, a);
  }
}

__global__ void conv0_stats_kernel(const double* __restrict__ mom, const float* __restrict__ w,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, int C, int L,
                                   float eps, float* __restrict__ scale, float* __restrict__ shift,
                                   const int* __restrict__ lens, int N, int normalize) {
  // GroupNorm affine folded to y = conv * scale + shift  (scale = gamma * rstd, shift = beta - mean * scale)
  // normalize: the network input is x' = (x - mu) * s with the utterance's mu and s = 1 / (std_unbiased + 1e-5), but the
  // GEMM convolves the RAW x.  conv(x') = s (conv(x) - mu sum_k w_k), so mean' = s (mean - mu W), var' = s^2 var and
  //     y = gamma (conv(x') - mean') / sqrt(var' + eps) + beta = [gamma s / sqrt(s^2 var + eps)] (conv(x) - mean) + beta:
  // the normaliser costs nothing but a different scale.
  const int b = blockIdx.y;
  const int n_b = lens != nullptr ? lens[b] : N;
  if (lens != nullptr) L = (n_b - C0_K) / C0_S + 1;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* m = mom + int64_t(b) * C0_MOMS;
  double s_in = 1.0;
  if (normalize) {
    const double ts = m[C0_NMOM], tq = m[C0_NMOM + 1];
    const double mu = ts / double(n_b);
    double var_x = n_b > 1 ? (tq - ts * mu) / double(n_b - 1) : 0.0;       // unbiased, like torch.std_mean
    if (var_x < 0.0) var_x = 0.0;
    s_in = 1.0 / (double(float(sqrt(var_x))) + 1e-5);                        // (std rounded to fp32 like the reference's tensor)
  }
  const double invL = 1.0 / double(L);
  double wk[C0_K];
#pragma unroll
  for (int k = 0; k < C0_K; ++k) wk[k] = double(w[c * C0_K + k]);
  double mean = 0.0;
#pragma unroll
  for (int k = 0; k < C0_K; ++k) mean += wk[k] * m[k] * invL;
  double var = 0.0;
  int idx = C0_K;
#pragma unroll
  for (int k = 0; k < C0_K; ++k) {
#pragma unroll
    for (int k2 = k; k2 < C0_K; ++k2) {
      const double cov = m[idx++] * invL - (m[k] * invL) * (m[k2] * invL);
      var += (k == k2 ? 1.0 : 2.0) * wk[k] * wk[k2] * cov;
    }
  }
  if (var < 0.0) var = 0.0;
  const double sc = double(gamma[c]) * s_in / sqrt(s_in * s_in * var + double(eps));
  scale[int64_t(b) * C + c] = float(sc);
  shift[int64_t(b) * C + c] = float(double(beta[c]) - mean * sc);
}

// Operands of the conv-as-GEMM: the 10-tap windows of the waveform and the filter bank, both as
// error-compensated fp16 (v = hi + lo) over K = 64:  A' = [x_hi(10) | x_lo(10) | x_hi(10) | 0...],
// W' = [w_hi | w_hi | w_lo | 0...]  ->  the fp32 TMEM accumulator holds x.w to ~2^-20 relative.
constexpr int C0_KP = 64;
template <bool PCM16>
__global__ void conv0_im2col_kernel(const void* __restrict__ wav, int N, int L, __half* __restrict__ a) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  const int64_t x0 = int64_t(b) * N + int64_t(t) * C0_S;
  __half hi[C0_K], lo[C0_K];
#pragma unroll
  for (int k = 0; k < C0_K; ++k) {
    const float v = c0_load<PCM16>(wav, x0 + k);
    hi[k] = __float2half_rn(v);
    lo[k] = __float2half_rn(v - __half2float(hi[k]));
  }
  __half row[C0_KP];
#pragma unroll
  for (int k = 0; k < C0_KP; ++k) row[k] = __float2half_rn(0.f);
#pragma unroll
  for (int k = 0; k < C0_K; ++k) { row[k] = hi[k]; row[C0_K + k] = lo[k]; row[2 * C0_K + k] = hi[k]; }
  uint4* dst = reinterpret_cast<uint4*>(a + (int64_t(b) * L + t) * C0_KP);
  const uint4* src = reinterpret_cast<const uint4*>(row);
#pragma unroll
  for (int q = 0; q < C0_KP / 8; ++q) dst[q] = src[q];
}
__global__ void conv0_weight_split_kernel(const float* __restrict__ w, __half* __restrict__ o, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  __half* row = o + int64_t(c) * C0_KP;
  for (int k = 0; k < C0_KP; ++k) row[k] = __float2half_rn(0.f);
  for (int k = 0; k < C0_K; ++k) {
    const float v = w[c * C0_K + k];
    const __half hi = __float2half_rn(v);
    row[k] = hi;
    row[C0_K + k] = hi;
    row[2 * C0_K + k] = __float2half_rn(v - __half2float(hi));
  }
}

// =================================================================================================
// LayerNorm(drop(x + bias) + residual) -> f32 and/or f16   (one warp per row, values kept in registers;
// thr > 0: counter-based inverted dropout of the branch, regenerated in the backward from the same seed)

// EXACT: H == 128 * MAXV (no per-slot bounds checks, so all loads of a row issue back to back)
template <bool X_F32, int MAXV, bool EXACT>
__global__ void __launch_bounds__(256) layernorm_kernel(const void* __restrict__ x_, const float* __restrict__ bias,
                                                        const float* __restrict__ residual,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, float* __restrict__ y32, __half* __restrict__ y16,
                                                        int64_t rows, int H, uint32_t thr, float inv_keep, uint64_t seed,
                                                        float* __restrict__ rstd_out) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const DropKeys dkeys = drop_keys(seed);
  float4 v[MAXV];                               // MAXV = ceil(H / 128) float4 slots per lane
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (EXACT || c < H) {
      float4 a;
      if constexpr (X_F32) {
        a = *reinterpret_cast<const float4*>(static_cast<const float*>(x_) + row * H + c);
      } else {
        const uint2 q = *reinterpret_cast<const uint2*>(static_cast<const __half*>(x_) + row * H + c);
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
        a = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
      if (bias != nullptr) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
        a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
      }
      if (thr != 0) dropout4(a, dkeys, row * H + c, thr, inv_keep);
      if (residual != nullptr) {
        const float4 r = *reinterpret_cast<const float4*>(residual + row * H + c);
        a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
      }
      v[i] = a;
      sum += (a.x + a.y) + (a.z + a.w);
    }
  }
  const float mean = warp_sum(sum) / float(H);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (EXACT || c < H) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / float(H) + eps);
  if (rstd_out != nullptr && lane == 0) rstd_out[row] = rstd;       // kept for w2v2_layernorm_bwd_from_output
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (EXACT || c < H) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + bt.x;
      o.y = (v[i].y - mean) * rstd * g.y + bt.y;
      o.z = (v[i].z - mean) * rstd * g.z + bt.z;
      o.w = (v[i].w - mean) * rstd * g.w + bt.w;
      if (y32 != nullptr) *reinterpret_cast<float4*>(y32 + row * H + c) = o;
      if (y16 != nullptr) {
        uint2 q;
        q.x = pack_half2(o.x, o.y);
        q.y = pack_half2(o.z, o.w);
        *reinterpret_cast<uint2*>(y16 + row * H + c) = q;
      }
    }
  }
}

// =================================================================================================
// pooling over time.  x [B,T,H]; one thread per (b, channel), T split over `TS` thread rows of the
// block and combined through shared memory; reads are coalesced along channels.

constexpr int POOL_TS = 8;     // time slices per block
constexpr int POOL_CH = 32;    // channels per block

__device__ __forceinline__ float block_slices_sum(float v, float (*sm)[POOL_CH], int ts, int ch) {
  __syncthreads();
  sm[ts][ch] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < POOL_TS; ++i) s += sm[i][ch];
  return s;
}
__device__ __forceinline__ float block_slices_max(float v, float (*sm)[POOL_CH], int ts, int ch) {
  __syncthreads();
  sm[ts][ch] = v;
  __syncthreads();
  float s = sm[0][ch];
#pragma unroll
  for (int i = 1; i < POOL_TS; ++i) s = fmaxf(s, sm[i][ch]);
  return s;
}

// lens (optional, all three pooling kernels): valid frames of utterance b in a padded ragged batch; the row pitch stays T
__global__ void __launch_bounds__(POOL_TS* POOL_CH) stat_pool_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                      int T, int H, int mode, const int* __restrict__ lens) {
  __shared__ float sm[POOL_TS][POOL_CH];
  const int ch = threadIdx.x % POOL_CH, ts = threadIdx.x / POOL_CH;
  const int c = blockIdx.x * POOL_CH + ch;
  const int b = blockIdx.y;
  const float* xb = x + int64_t(b) * T * H + c;
  if (lens != nullptr) T = lens[b];                  // (xb already points at the utterance)
  if (mode == 2) {
    float m = -INFINITY;
    for (int t = ts; t < T; t += POOL_TS) m = fmaxf(m, xb[int64_t(t) * H]);
    m = block_slices_max(m, sm, ts, ch);
    if (ts == 0) out[int64_t(b) * H + c] = m;
    return;
  }
  float s = 0.f;
  for (int t = ts; t < T; t += POOL_TS) s += xb[int64_t(t) * H];
  const float mean = block_slices_sum(s, sm, ts, ch) / float(T);
  if (mode == 0) {
    if (ts == 0) out[int64_t(b) * H + c] = mean;
    return;
  }
  float q = 0.f;
  for (int t = ts; t < T; t += POOL_TS) {
    const float d = xb[int64_t(t) * H] - mean;
    q = fmaf(d, d, q);
  }
  q = block_slices_sum(q, sm, ts, ch);
  if (ts == 0) {
    if (mode == 3) {                                                // ASP front statistics: [mean || std (population, clamped)]
      out[int64_t(b) * 2 * H + c] = mean;
      out[int64_t(b) * 2 * H + H + c] = sqrtf(fmaxf(q / float(T), 1e-12f));
    } else {
      out[int64_t(b) * 2 * H + c] = sqrtf(q / float(T - 1));      // [std (unbiased) || mean]
      out[int64_t(b) * 2 * H + H + c] = mean;
    }
  }
}

// ASP front: uniform-weight mean/std (eps-clamped) + concatenated fp16 operand [x | mean | std]
__global__ void __launch_bounds__(POOL_TS* POOL_CH) asp_concat_kernel(const float* __restrict__ x,
                                                                       __half* __restrict__ cat, int T, int H,
                                                                       const int* __restrict__ lens) {
  __shared__ float sm[POOL_TS][POOL_CH];
  const int ch = threadIdx.x % POOL_CH, ts = threadIdx.x / POOL_CH;
  const int c = blockIdx.x * POOL_CH + ch;
  const int b = blockIdx.y;
  const float* xb = x + int64_t(b) * T * H + c;
  const int Tv = lens != nullptr ? lens[b] : T;      // statistics over the valid frames; every row of `cat` is written
  const float m = 1.0f / float(Tv);
  float s = 0.f;
  for (int t = ts; t < Tv; t += POOL_TS) s = fmaf(m, xb[int64_t(t) * H], s);
  const float mean = block_slices_sum(s, sm, ts, ch);
  float q = 0.f;
  for (int t = ts; t < Tv; t += POOL_TS) {
    const float d = xb[int64_t(t) * H] - mean;
    q = fmaf(m * d, d, q);
  }
  const float stdv = sqrtf(fmaxf(block_slices_sum(q, sm, ts, ch), 1e-12f));
  const __half hm = __float2half_rn(mean), hs = __float2half_rn(stdv);
  __half* cb = cat + int64_t(b) * T * 3 * H;
  for (int t = ts; t < T; t += POOL_TS) {
    __half* row = cb + int64_t(t) * 3 * H;
    row[c] = __float2half_rn(xb[int64_t(t) * H]);
    row[H + c] = hm;
    row[2 * H + c] = hs;
  }
}

// ubias (optional, f32 [rows / rows_per_utt, A]): a per-utterance bias added to z first -- the share of the TDNN's 1x1 conv
// that multiplies the utterance's mean / std columns, which are constant over its frames
__global__ void asp_relu_bn_tanh_kernel(const float* __restrict__ z, const float* __restrict__ scale,
                                        const float* __restrict__ shift, __half* __restrict__ y, int64_t n, int A,
                                        const float* __restrict__ ubias, int64_t per_utt) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int a = i % A;
    float zi = z[i];
    if (ubias != nullptr) zi += __ldg(ubias + (i / per_utt) * A + a);
    const float r = fmaxf(zi, 0.f);
    y[i] = __float2half_rn(tanhf(fmaf(r, __ldg(scale + a), __ldg(shift + a))));
  }
}

// ASP tail: softmax over T per (b,c) of the attention logits, weighted mean / std
__global__ void __launch_bounds__(POOL_TS* POOL_CH) asp_pool_kernel(const float* __restrict__ x,
                                                                     const float* __restrict__ lg, float* __restrict__ out,
                                                                     int T, int H, const int* __restrict__ lens) {
  __shared__ float sm[POOL_TS][POOL_CH];
  const int ch = threadIdx.x % POOL_CH, ts = threadIdx.x / POOL_CH;
  const int c = blockIdx.x * POOL_CH + ch;
  const int b = blockIdx.y;
  const float* xb = x + int64_t(b) * T * H + c;
  const float* lb = lg + int64_t(b) * T * H + c;
  if (lens != nullptr) T = lens[b];                  // softmax over the valid frames only (speechbrain's length mask)
  float mx = -INFINITY;
  for (int t = ts; t < T; t += POOL_TS) mx = fmaxf(mx, lb[int64_t(t) * H]);
  mx = block_slices_max(mx, sm, ts, ch);
  float se = 0.f, sx = 0.f;
  for (int t = ts; t < T; t += POOL_TS) {
    const float e = expf(lb[int64_t(t) * H] - mx);
    se += e;
    sx = fmaf(e, xb[int64_t(t) * H], sx);
  }
  se = block_slices_sum(se, sm, ts, ch);
  sx = block_slices_sum(sx, sm, ts, ch);
  const float inv = 1.0f / se;
  const float mean = sx * inv;
  float q = 0.f;
  for (int t = ts; t < T; t += POOL_TS) {
    const float e = expf(lb[int64_t(t) * H] - mx) * inv;
    const float d = xb[int64_t(t) * H] - mean;
    q = fmaf(e * d, d, q);
  }
  q = block_slices_sum(q, sm, ts, ch);
  if (ts == 0) {
    out[int64_t(b) * 2 * H + c] = mean;                               // [mean || std]
    out[int64_t(b) * 2 * H + H + c] = sqrtf(fmaxf(q, 1e-12f));
  }
}

// =================================================================================================
// heads: softmax + CE + argmax per row (one block per row); optional AAM margin applied in place

__device__ __forceinline__ float block_reduce_sum(float v, float* sm) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < int(blockDim.x >> 5); ++i) s += sm[i];
  return s;
}

__global__ void __launch_bounds__(256) softmax_ce_kernel(float* __restrict__ logits, int64_t ldl,
                                                         const int64_t* __restrict__ labels, int aam, float cos_m,
                                                         float sin_m, float th, float mm, float scale, int easy,
                                                         float* __restrict__ prob, float* __restrict__ loss_rows,
                                                         int32_t* __restrict__ argmax, float* __restrict__ cos_label,
                                                         int S) {
  __shared__ float smf[8];
  __shared__ float smv[8];
  __shared__ int smi[8];
  const int b = blockIdx.x;
  float* row = logits + int64_t(b) * ldl;
  const int64_t label64 = labels[b];
  // torch raises on a class index outside [0, S) (device-side assert); here the kernel traps instead of reading
  // row[label] out of bounds / leaving cos_label unset for the backward
  if (label64 < 0 || label64 >= S) __trap();
  const int label = int(label64);
  if (aam) {
    // R:src/optim/loss/aam_softmax.py:56-69
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
      const float c = row[i];
      float o = c;
      if (i == label) {
        if (cos_label != nullptr) cos_label[b] = c;          // pre-margin cosine, needed by the backward
        const float sine = sqrtf(fminf(fmaxf(1.0f - c * c, 0.f), 1.f));
        const float phi = c * cos_m - sine * sin_m;
        o = easy ? (c > 0.f ? phi : c) : ((c - th) > 0.f ? phi : c - mm);
      }
      row[i] = o * scale;
    }
    __syncthreads();
  }
  // max + argmax (first occurrence)
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float v = row[i];
    if (v > mx) { mx = v; mi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { smv[threadIdx.x >> 5] = mx; smi[threadIdx.x >> 5] = mi; }
  __syncthreads();
  mx = smv[0]; mi = smi[0];
  for (int i = 1; i < int(blockDim.x >> 5); ++i)
    if (smv[i] > mx || (smv[i] == mx && smi[i] < mi)) { mx = smv[i]; mi = smi[i]; }
  float se = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) se += expf(row[i] - mx);
  se = block_reduce_sum(se, smf);
  const float inv = 1.0f / se;
  if (prob != nullptr)
    for (int i = threadIdx.x; i < S; i += blockDim.x) prob[int64_t(b) * S + i] = expf(row[i] - mx) * inv;
  if (threadIdx.x == 0) {
    loss_rows[b] = -(row[label] - mx - logf(se));
    argmax[b] = mi;
  }
}

__global__ void __launch_bounds__(256) l2norm_rows_kernel(const float* __restrict__ x, __half* __restrict__ y, int E,
                                                          int split3, int which) {
  // y = x / max(||x||, 1e-12).  split3: emit the error-compensated fp16 operand of width 3E:
  //   which == 0 (activation side): [hi | lo | hi];  which == 1 (weight side): [hi | hi | lo]
  __shared__ float smf[8];
  const int64_t r = blockIdx.x;
  const float* xr = x + r * E;
  float s = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) s = fmaf(xr[i], xr[i], s);
  s = block_reduce_sum(s, smf);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  if (!split3) {
    for (int i = threadIdx.x; i < E; i += blockDim.x) y[r * E + i] = __float2half_rn(xr[i] * inv);
  } else {
    __half* yr = y + r * 3 * E;
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
      const float v = xr[i] * inv;
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      yr[i] = hi;
      yr[E + i] = which == 0 ? lo : hi;
      yr[2 * E + i] = which == 0 ? hi : lo;
    }
  }
}

__global__ void split3_rows_kernel(const float* __restrict__ x, __half* __restrict__ y, int64_t rows, int E, int which) {
  const int64_t n = rows * E;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / E;
    const int c = i % E;
    const float v = x[i];
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    __half* yr = y + r * 3 * E;
    yr[c] = hi;
    yr[E + c] = which == 0 ? lo : hi;
    yr[2 * E + c] = which == 0 ? hi : lo;
  }
}

__global__ void mean_rows_kernel(const float* __restrict__ x, float* __restrict__ out, int n) {
  __shared__ float smf[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = block_reduce_sum(s, smf);
  if (threadIdx.x == 0) out[0] = s / float(n);
}

// weight-norm fold of the positional conv (HF:340-358): norm over (out, in) per tap
__global__ void posconv_norm_kernel(const float* __restrict__ v, float* __restrict__ norm, int H, int I, int K) {
  __shared__ float smf[8];
  const int k = blockIdx.x;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < int64_t(H) * I; i += blockDim.x) {
    const float a = v[i * K + k];
    s = fmaf(a, a, s);
  }
  s = block_reduce_sum(s, smf);
  if (threadIdx.x == 0) norm[k] = sqrtf(s);
}
__global__ void posconv_fold_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                    const float* __restrict__ norm, __half* __restrict__ w16, int H, int G, int K, int U,
                                    int mode) {
  // w16[grp][j'][c][u][o][e] = g[k] * v[grp*O + o][c*8 + e][k] / norm[k],  k = U*j' + u   (O = I = H/G)
  // = per (group, tap group j') a [U*O x I] K-major block already in UMMA no-swizzle core-matrix order
  // (planes of U*O rows x 16 bytes), so posconv.cu can stream it with a flat bulk copy.
  const int O = H / G, I = H / G;
  const int64_t n = int64_t(H) * I * K;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < n; idx += int64_t(gridDim.x) * blockDim.x) {
    const int e = idx % 8;
    const int o = (idx / 8) % O;
    const int u = (idx / (8 * int64_t(O))) % U;
    const int c = (idx / (8 * int64_t(O) * U)) % (I / 8);
    const int jp = (idx / (int64_t(I) * O * U)) % (K / U);
    const int grp = idx / (int64_t(I) * O * K);
    // mode 0: forward weight.  mode 1: data-gradient weight = in/out channels swapped, taps reversed
    //         (dx[s,i] = sum_{j',o} dz[s + j' + 1 - K/2, o] * w[o, i, K-1-j'];  the +1 is posconv's in_shift)
    const int k = mode == 0 ? U * jp + u : K - 1 - (U * jp + u);
    const int oo = mode == 0 ? o : c * 8 + e;
    const int ii = mode == 0 ? c * 8 + e : o;
    const float val = v[(int64_t(grp * O + oo) * I + ii) * K + k] * (g[k] / norm[k]);
    w16[idx] = __float2half_rn(val);
  }
}

// Coalesced, deterministic two-stage form of the same reduction (taps are the fastest axis of v [H, I, K]): stage 1,
// every block sums the squares of its slice of (out, in) rows per tap (lane = tap); stage 2 adds the per-block
// partials in a fixed order.  ~6 us instead of 32 for the 4.7 M weights (the strided kernel above reads one
// 4-byte value per 32-byte sector).
constexpr int PN_BLOCKS = 592;     // 4 per SM: enough loads in flight for a 19 MB streaming read
__device__ float g_pn_partial[PN_BLOCKS * 256];

__global__ void __launch_bounds__(256) posconv_norm_partial_kernel(const float* __restrict__ v, float* __restrict__ partial,
                                                                   int64_t rows, int K) {
  __shared__ float sm[256];
  const int k = threadIdx.x % K, rl = threadIdx.x / K, nrl = 256 / K;
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  for (int64_t r = r0 + rl; r < r1; r += nrl) {
    const float a = v[r * K + k];
    s = fmaf(a, a, s);
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  if (rl == 0) {
    for (int j = 1; j < nrl; ++j) s += sm[j * K + k];
    partial[blockIdx.x * K + k] = s;
  }
}
// one warp per tap: lanes stride over the per-block partials, then a shuffle tree (a fixed order: deterministic)
__global__ void __launch_bounds__(256) posconv_partial_finish_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                                     int nblocks, int K, int take_sqrt) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= K) return;
  float s = 0.f;
  for (int b = threadIdx.x & 31; b < nblocks; b += 32) s += partial[b * K + k];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) out[k] = take_sqrt ? sqrtf(s) : s;
}

float* posconv_partial_buffer() {
  static float* p = nullptr;
  if (p == nullptr) cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_pn_partial);
  return p;
}

void posconv_tap_norms(const float* v, float* norm, int H, int I, int K, cudaStream_t stream) {
  if (K > 256 || 256 % K != 0) {
    posconv_norm_kernel<<<K, 256, 0, stream>>>(v, norm, H, I, K);
    return;
  }
  float* partial = posconv_partial_buffer();
  posconv_norm_partial_kernel<<<PN_BLOCKS, 256, 0, stream>>>(v, partial, int64_t(H) * I, K);
  posconv_partial_finish_kernel<<<(K + 7) / 8, 256, 0, stream>>>(partial, norm, PN_BLOCKS, K, 1);
}

// Fold with coalesced traffic: one block per (group, 8-channel plane c, 16-tap chunk).  It gathers
// v[., ., 16 taps] of the 8 x O (or 8 x I) channel pairs it needs -- 64 contiguous bytes each -- into shared
// memory, then writes 16-byte groups of 8 halves.  Same output layout as posconv_fold_kernel.
constexpr int PF_TAPS = 16;
template <int MODE>
__global__ void __launch_bounds__(256) posconv_fold_tiled_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                                 const float* __restrict__ norm, __half* __restrict__ w16,
                                                                 int H, int G, int K, int U) {
  extern __shared__ float tile[];                         // [A][8][PF_TAPS + 1]
  const int O = H / G, I = H / G;
  const int kchunks = K / PF_TAPS, planes = I / 8;
  const int kc = blockIdx.x % kchunks;
  const int c = (blockIdx.x / kchunks) % planes;
  const int grp = blockIdx.x / (kchunks * planes);
  const int A = MODE == 0 ? O : I;                        // the channel index that becomes `o` of the output layout
  constexpr int TP = PF_TAPS + 1;
  // gather: element (a, e, kk); mode 0: v[grp*O + a][c*8 + e][k], mode 1: v[grp*O + c*8 + e][a][K-1-k], k = kc*16 + kk
  for (int idx = threadIdx.x; idx < A * 8 * PF_TAPS; idx += 256) {
    const int kk = idx % PF_TAPS, e = (idx / PF_TAPS) % 8, a = idx / (8 * PF_TAPS);
    const int k = kc * PF_TAPS + kk;
    float val;
    if (MODE == 0) val = v[(int64_t(grp * O + a) * I + c * 8 + e) * K + k] * (g[k] / norm[k]);
    else {
      // consecutive kk walk the source taps backwards; read them forwards so that a warp still covers whole sectors
      const int ks = K - 1 - (kc * PF_TAPS + (PF_TAPS - 1 - kk));      // = source tap of output tap kc*16 + (15 - kk)
      val = v[(int64_t(grp * O + c * 8 + e) * I + a) * K + ks] * (g[ks] / norm[ks]);
      tile[(a * 8 + e) * TP + (PF_TAPS - 1 - kk)] = val;
      continue;
    }
    tile[(a * 8 + e) * TP + kk] = val;
  }
  __syncthreads();
  // scatter: 8 halves (e = 0..7) per (tap, a)
  for (int idx = threadIdx.x; idx < PF_TAPS * A; idx += 256) {
    const int a = idx % A, kk = idx / A;
    const int k = kc * PF_TAPS + kk;
    const int jp = k / U, u = k % U;
    uint32_t pk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      pk[q] = pack_half2(tile[(a * 8 + 2 * q) * TP + kk], tile[(a * 8 + 2 * q + 1) * TP + kk]);
    const int64_t dst = ((((int64_t(grp) * (K / U) + jp) * (I / 8) + c) * U + u) * O + a) * 8;
    *reinterpret_cast<uint4*>(w16 + dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

static inline int grid_for(int64_t n, int per_block) {
  int64_t g = (n + per_block - 1) / per_block;
  const int64_t cap = int64_t(device_sm_count()) * 16;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace w2v2

using namespace w2v2;

extern "C" {

const char* w2v2_last_error(void) { return w2v2::g_err; }
int w2v2_abi_version(void) { return 1; }
int w2v2_sm_count(void) { return device_sm_count(); }
int64_t w2v2_launch_count(void) { return (int64_t)__atomic_load_n(&w2v2::g_launches, __ATOMIC_RELAXED); }

int w2v2_cast_f16(const float* x, void* y16, int64_t n, float scale, void* stream) {
  W2V2_REQUIRE(n >= 0, "w2v2_cast_f16: negative n");
  if (n == 0) return 0;
  cast_f16_kernel<<<grid_for(n, 256 * 8), 256, 0, (cudaStream_t)stream>>>(x, (__half*)y16, n, scale);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_cast_f16_rowmask(const float* x, void* y16, int B, int T, int H, const int* lens, void* stream) {
  W2V2_REQUIRE(H % 4 == 0 && lens != nullptr, "w2v2_cast_f16_rowmask: H %% 4 == 0 and lens are required");
  const int64_t n4 = int64_t(B) * T * (H / 4);
  if (n4 == 0) return 0;
  cast_f16_rowmask_kernel<<<grid_for(n4, 256), 256, 0, (cudaStream_t)stream>>>(x, (__half*)y16, n4, T, H / 4, lens);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_conv_weight_tapmajor(const float* w, void* w16, int cout, int cin, int k, void* stream) {
  const int64_t n = int64_t(cout) * cin * k;
  conv_weight_tapmajor_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(w, (__half*)w16, cout, cin, k);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
struct Conv0Ws { int64_t scale, shift, mom, w, a, total; };
static Conv0Ws conv0_ws(int B, int N, int C) {
  const int64_t L = (N - C0_K) / C0_S + 1;
  Conv0Ws o;
  o.scale = 0;
  o.shift = align_up(o.scale + int64_t(B) * C * 4, 256);
  o.mom = align_up(o.shift + int64_t(B) * C * 4, 256);
  o.w = align_up(o.mom + int64_t(B) * C0_MOMS * 8, 256);
  o.a = align_up(o.w + int64_t(C) * C0_KP * 2, 256);
  o.total = align_up(o.a + int64_t(B) * L * C0_KP * 2, 256);
  return o;
}

int64_t w2v2_conv0_workspace_bytes(int B, int N, int C) { return N >= C0_K ? conv0_ws(B, N, C).total : 0; }

int w2v2_conv0_workspace_offsets(int B, int N, int C, int64_t* scale_off, int64_t* shift_off, int64_t* im2col_off) {
  W2V2_REQUIRE(N >= C0_K, "w2v2_conv0_workspace_offsets: N=%d too short", N);
  const Conv0Ws ws = conv0_ws(B, N, C);
  if (scale_off) *scale_off = ws.scale;
  if (shift_off) *shift_off = ws.shift;
  if (im2col_off) *im2col_off = ws.a;
  return 0;
}

int w2v2_conv0_gn_gelu(const float* wav, int B, int N, const float* w, const float* gamma, const float* beta, float eps,
                       void* workspace, void* out_f16, int C, void* stream_) {
  return w2v2_conv0_gn_ex(wav, B, N, w, gamma, beta, eps, workspace, out_f16, C, 1, stream_);
}

int w2v2_conv0_gn_ex(const float* wav, int B, int N, const float* w, const float* gamma, const float* beta, float eps,
                     void* workspace, void* out_f16, int C, int act, void* stream_) {
  return w2v2_conv0_gn_lens(wav, B, N, nullptr, w, gamma, beta, eps, workspace, out_f16, C, act, stream_);
}

int w2v2_conv0_gn_lens(const float* wav, int B, int N, const int* lens, const float* w, const float* gamma, const float* beta,
                       float eps, void* workspace, void* out_f16, int C, int act, void* stream_) {
  return w2v2_conv0_raw(wav, 1, 0, B, N, lens, w, gamma, beta, eps, workspace, out_f16, C, act, stream_);
}

int w2v2_conv0_raw(const void* wav, int in_dtype, int normalize, int B, int N, const int* lens, const float* w,
                   const float* gamma, const float* beta, float eps, void* workspace, void* out_f16, int C, int act,
                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  W2V2_REQUIRE(in_dtype == 0 || in_dtype == 1, "w2v2_conv0_raw: in_dtype 0 = int16 PCM, 1 = float32");
  W2V2_REQUIRE(act == 0 || act == 1, "w2v2_conv0_gn_ex: act must be 0 (GroupNorm output) or 1 (+ GELU)");
  W2V2_REQUIRE(B > 0 && N >= C0_K, "w2v2_conv0_gn_gelu: need B>0 and N>=10 (got B=%d N=%d)", B, N);
  W2V2_REQUIRE(C > 128 && C % 8 == 0, "w2v2_conv0_gn_gelu: C=%d must be a multiple of 8 and > 128", C);
  W2V2_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "w2v2_conv0_gn_gelu: workspace must be 256-byte aligned");
  const int L = (N - C0_K) / C0_S + 1;
  const Conv0Ws ws = conv0_ws(B, N, C);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  float* scale = reinterpret_cast<float*>(base + ws.scale);
  float* shift = reinterpret_cast<float*>(base + ws.shift);
  double* mom = reinterpret_cast<double*>(base + ws.mom);
  __half* w16 = reinterpret_cast<__half*>(base + ws.w);
  __half* a16 = reinterpret_cast<__half*>(base + ws.a);
  W2V2_CHECK_CUDA(cudaMemsetAsync(mom, 0, sizeof(double) * B * C0_MOMS, stream));
  dim3 g1((L + 256 * 8 - 1) / (256 * 8), B);
  if (in_dtype == 0) {
    if (normalize) conv0_moments_kernel<true, true><<<g1, 256, 0, stream>>>(wav, N, L, mom, lens);
    else conv0_moments_kernel<true, false><<<g1, 256, 0, stream>>>(wav, N, L, mom, lens);
  } else {
    if (normalize) conv0_moments_kernel<false, true><<<g1, 256, 0, stream>>>(wav, N, L, mom, lens);
    else conv0_moments_kernel<false, false><<<g1, 256, 0, stream>>>(wav, N, L, mom, lens);
  }
  dim3 g2((C + 127) / 128, B);
  conv0_stats_kernel<<<g2, 128, 0, stream>>>(mom, w, gamma, beta, C, L, eps, scale, shift, lens, N, normalize);
  conv0_weight_split_kernel<<<(C + 127) / 128, 128, 0, stream>>>(w, w16, C);
  dim3 g3((L + 255) / 256, B);
  if (in_dtype == 0) conv0_im2col_kernel<true><<<g3, 256, 0, stream>>>(wav, N, L, a16);
  else conv0_im2col_kernel<false><<<g3, 256, 0, stream>>>(wav, N, L, a16);
  count_launches(4);
  W2V2_CHECK_CUDA(cudaGetLastError());
  // y = GELU(conv * scale[b,c] + shift[b,c]) on the tensor cores, fp16 channels-last out
  return gemm_f16_impl(a16, L, C0_KP, int64_t(L) * C0_KP, B, 1, 0, C0_KP, w16, C0_KP, C, scale, shift, act, out_f16, 0, C,
                       int64_t(L) * C, stream);
}

int w2v2_layernorm(const void* x, int x_dtype, const float* bias, const float* residual, const float* gamma,
                   const float* beta, float eps, float* y32, void* y16, int64_t rows, int H, void* stream) {
  return w2v2_layernorm_ex(x, x_dtype, bias, residual, gamma, beta, eps, y32, y16, rows, H, 0.f, 0, stream);
}

int w2v2_layernorm_ex(const void* x, int x_dtype, const float* bias, const float* residual, const float* gamma,
                      const float* beta, float eps, float* y32, void* y16, int64_t rows, int H, float drop_p,
                      uint64_t drop_seed, void* stream) {
  return w2v2_layernorm_ex2(x, x_dtype, bias, residual, gamma, beta, eps, y32, y16, nullptr, rows, H, drop_p, drop_seed, stream);
}

int w2v2_layernorm_ex2(const void* x, int x_dtype, const float* bias, const float* residual, const float* gamma,
                       const float* beta, float eps, float* y32, void* y16, float* rstd_out, int64_t rows, int H,
                       float drop_p, uint64_t drop_seed, void* stream) {
  W2V2_REQUIRE(H % 4 == 0 && H <= 1024, "w2v2_layernorm: H=%d must be a multiple of 4 and <= 1024", H);
  W2V2_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "w2v2_layernorm: drop_p=%f out of [0,1)", drop_p);
  if (rows == 0) return 0;
  const uint32_t thr = uint32_t(drop_p * 65536.0f + 0.5f);
  const float inv_keep = 1.0f / (1.0f - float(thr) / 65536.0f);
  const int grid = int((rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
#define W2V2_LN(F32, NV, EX) launch_k(layernorm_kernel<F32, NV, EX>, dim3(grid), dim3(256), 0, st, 1, x, bias, residual, gamma, beta, eps, y32, (__half*)y16, rows, H, thr, inv_keep, drop_seed, rstd_out)
  if (x_dtype == 1) {
    if (H == 512) W2V2_LN(true, 4, true); else if (H == 768) W2V2_LN(true, 6, true);
    else if (H == 1024) W2V2_LN(true, 8, true); else W2V2_LN(true, 8, false);
  } else {
    if (H == 512) W2V2_LN(false, 4, true); else if (H == 768) W2V2_LN(false, 6, true);
    else if (H == 1024) W2V2_LN(false, 8, true); else W2V2_LN(false, 8, false);
  }
#undef W2V2_LN
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_stat_pool(const float* x, float* out, int B, int T, int H, int mode, void* stream) {
  return w2v2_stat_pool_lens(x, out, B, T, H, mode, nullptr, stream);
}
int w2v2_stat_pool_lens(const float* x, float* out, int B, int T, int H, int mode, const int* lens, void* stream) {
  W2V2_REQUIRE(H % POOL_CH == 0, "w2v2_stat_pool: H=%d must be a multiple of %d", H, POOL_CH);
  W2V2_REQUIRE(mode >= 0 && mode <= 3, "w2v2_stat_pool: unknown mode %d", mode);
  W2V2_REQUIRE(T >= 1 && (mode != 1 || T >= 2), "w2v2_stat_pool: T=%d too short", T);
  dim3 g(H / POOL_CH, B);
  stat_pool_kernel<<<g, POOL_TS * POOL_CH, 0, (cudaStream_t)stream>>>(x, out, T, H, mode, lens);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_asp_concat(const float* x, void* cat16, int B, int T, int H, void* stream) {
  return w2v2_asp_concat_lens(x, cat16, B, T, H, nullptr, stream);
}
int w2v2_asp_concat_lens(const float* x, void* cat16, int B, int T, int H, const int* lens, void* stream) {
  W2V2_REQUIRE(H % POOL_CH == 0, "w2v2_asp_concat: H=%d must be a multiple of %d", H, POOL_CH);
  dim3 g(H / POOL_CH, B);
  asp_concat_kernel<<<g, POOL_TS * POOL_CH, 0, (cudaStream_t)stream>>>(x, (__half*)cat16, T, H, lens);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_asp_relu_bn_tanh(const float* z, const float* scale, const float* shift, void* y16, int64_t rows, int A,
                          void* stream) {
  return w2v2_asp_relu_bn_tanh_ubias(z, scale, shift, nullptr, 1, y16, rows, A, stream);
}
int w2v2_asp_relu_bn_tanh_ubias(const float* z, const float* scale, const float* shift, const float* ubias, int rows_per_utt,
                                void* y16, int64_t rows, int A, void* stream) {
  W2V2_REQUIRE(rows_per_utt >= 1 && rows % rows_per_utt == 0, "w2v2_asp_relu_bn_tanh: rows=%lld not a multiple of %d",
               (long long)rows, rows_per_utt);
  const int64_t n = rows * A;
  asp_relu_bn_tanh_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(z, scale, shift, (__half*)y16, n, A, ubias,
                                                                              int64_t(rows_per_utt) * A);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_asp_pool(const float* x, const float* logits, float* out, int B, int T, int H, void* stream) {
  return w2v2_asp_pool_lens(x, logits, out, B, T, H, nullptr, stream);
}
int w2v2_asp_pool_lens(const float* x, const float* logits, float* out, int B, int T, int H, const int* lens, void* stream) {
  W2V2_REQUIRE(H % POOL_CH == 0, "w2v2_asp_pool: H=%d must be a multiple of %d", H, POOL_CH);
  dim3 g(H / POOL_CH, B);
  asp_pool_kernel<<<g, POOL_TS * POOL_CH, 0, (cudaStream_t)stream>>>(x, logits, out, T, H, lens);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_softmax_ce(const float* logits, int64_t ldl, const int64_t* labels, float* prob, float* loss_rows,
                    int32_t* argmax, int B, int S, void* stream) {
  softmax_ce_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(const_cast<float*>(logits), ldl, labels, 0, 0.f, 0.f, 0.f, 0.f,
                                                         1.f, 0, prob, loss_rows, argmax, nullptr, S);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_aam_softmax_ce_ex(float* cosine, int64_t ldl, const int64_t* labels, float margin, float scale, int easy_margin,
                           float* prob, float* loss_rows, int32_t* argmax, float* cos_label, int B, int S, void* stream) {
  const double m = margin;
  const double pi = 3.14159265358979323846;
  softmax_ce_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(cosine, ldl, labels, 1, float(cos(m)), float(sin(m)),
                                                         float(cos(pi - m)), float(sin(pi - m) * m), scale, easy_margin,
                                                         prob, loss_rows, argmax, cos_label, S);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_aam_softmax_ce(float* cosine, int64_t ldl, const int64_t* labels, float margin, float scale, int easy_margin,
                        float* prob, float* loss_rows, int32_t* argmax, int B, int S, void* stream) {
  return w2v2_aam_softmax_ce_ex(cosine, ldl, labels, margin, scale, easy_margin, prob, loss_rows, argmax, nullptr, B, S,
                                stream);
}

int w2v2_l2norm_rows_f16(const float* x, void* y16, int64_t rows, int E, void* stream) {
  l2norm_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, (__half*)y16, E, 0, 0);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_l2norm_rows_split3(const float* x, void* y16, int64_t rows, int E, int which, void* stream) {
  l2norm_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, (__half*)y16, E, 1, which);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_split3_rows(const float* x, void* y16, int64_t rows, int E, int which, void* stream) {
  split3_rows_kernel<<<grid_for(rows * E, 256), 256, 0, (cudaStream_t)stream>>>(x, (__half*)y16, rows, E, which);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_mean_rows(const float* x, float* out, int n, void* stream) {
  mean_rows_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(x, out, n);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_posconv_fold_weight(const float* v, const float* g, void* w16, int H, int groups, int K, int U, int mode,
                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  // scratch for the K norms: reuse the head of w16?  No -- keep a tiny static device buffer per call site:
  // the caller passes w16 sized H*(H/groups)*K halfs + K floats; norms live behind the weights.
  float* norm = reinterpret_cast<float*>(static_cast<__half*>(w16) + int64_t(H) * (H / groups) * K);
  posconv_tap_norms(v, norm, H, H / groups, K, stream);
  const int64_t n = int64_t(H) * (H / groups) * K;
  W2V2_REQUIRE(U >= 1 && K % U == 0, "w2v2_posconv_fold_weight: taps per MMA U=%d must divide K=%d", U, K);
  const int I = H / groups;
  if (K % PF_TAPS == 0 && I % 8 == 0 && (mode == 0 || mode == 1)) {
    const int blocks = groups * (I / 8) * (K / PF_TAPS);
    const size_t smem = size_t(I) * 8 * (PF_TAPS + 1) * sizeof(float);
    if (mode == 0) posconv_fold_tiled_kernel<0><<<blocks, 256, smem, stream>>>(v, g, norm, (__half*)w16, H, groups, K, U);
    else posconv_fold_tiled_kernel<1><<<blocks, 256, smem, stream>>>(v, g, norm, (__half*)w16, H, groups, K, U);
  } else {
    posconv_fold_kernel<<<grid_for(n, 256), 256, 0, stream>>>(v, g, norm, (__half*)w16, H, groups, K, U, mode);
  }
  count_launches(3);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
