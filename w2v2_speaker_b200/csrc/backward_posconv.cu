// Training-only helpers around the positional conv and the GELU activations:
//   * standalone GELU forward (training keeps the pre-activation, so GEMM / posconv run without the fused GELU)
//   * im2col of one channel group for the positional-conv weight gradient (then a plain wgrad GEMM)
//   * weight-norm backward (dW of the folded weight -> dg, dv of the parametrisation, HF:340-358)
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int device_sm_count();
void posconv_tap_norms(const float* v, float* norm, int H, int I, int K, cudaStream_t stream);
static inline int pgrid(int64_t n, int per_block, int per_sm) {
  int64_t g = (n + per_block - 1) / per_block;
  const int64_t cap = int64_t(device_sm_count()) * per_sm;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

template <bool IN_F32, bool OUT_F32>
__global__ void gelu_fwd_kernel(const void* __restrict__ x_, void* __restrict__ y_, __half* __restrict__ x16_copy,
                                int64_t n) {
  pdl_trigger();
  pdl_wait();
  // n % 4 == 0.  Optionally also writes the (rounded) input as f16 (saved pre-activation).
  int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x * 4;
  for (; i < n; i += stride) {
    float v[4];
    if constexpr (IN_F32) {
      const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(x_) + i);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else {
      const uint2 q = *reinterpret_cast<const uint2*>(static_cast<const __half*>(x_) + i);
      const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
      const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
      v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
    }
    if (x16_copy != nullptr) {
      uint2 q;
      q.x = pack_half2(v[0], v[1]);
      q.y = pack_half2(v[2], v[3]);
      *reinterpret_cast<uint2*>(x16_copy + i) = q;
    }
    gelu_erf2(v[0], v[1]);
    gelu_erf2(v[2], v[3]);
    if constexpr (OUT_F32) {
      *reinterpret_cast<float4*>(static_cast<float*>(y_) + i) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      uint2 q;
      q.x = pack_half2(v[0], v[1]);
      q.y = pack_half2(v[2], v[3]);
      *reinterpret_cast<uint2*>(static_cast<__half*>(y_) + i) = q;
    }
  }
}

// X_g[(b,t), j*I + i] = x[b, t + j - K/2, g*I + i]  (0 outside the utterance)
__global__ void posconv_im2col_kernel(const __half* __restrict__ x, __half* __restrict__ xg, int B, int T, int H, int I,
                                      int K, int g) {
  const int chunks = I / 8;
  const int64_t n = int64_t(B) * T * K * chunks;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < n; idx += int64_t(gridDim.x) * blockDim.x) {
    const int c = idx % chunks;
    const int j = (idx / chunks) % K;
    const int64_t row = idx / (int64_t(chunks) * K);
    const int t = row % T;
    const int b = row / T;
    const int ts = t + j - K / 2;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ts >= 0 && ts < T) v = *reinterpret_cast<const uint4*>(x + (int64_t(b) * T + ts) * H + g * I + c * 8);
    *reinterpret_cast<uint4*>(xg + row * (int64_t(K) * I) + j * I + c * 8) = v;
  }
}

// S_k = sum_{o,i} dW[o][k][i] * v[o][i][k]   with dW laid out [H][K][I] (row o of the wgrad output = [K*I])
__global__ void wnorm_bwd_reduce_kernel(const float* __restrict__ dw, const float* __restrict__ v, float* __restrict__ S,
                                        int H, int I, int K) {
  __shared__ float sm[8];
  const int k = blockIdx.x;
  float s = 0.f;
  for (int64_t idx = threadIdx.x; idx < int64_t(H) * I; idx += blockDim.x) {
    const int i = idx % I;
    const int o = idx / I;
    s = fmaf(dw[(int64_t(o) * K + k) * I + i], v[(int64_t(o) * I + i) * K + k], s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) t += sm[w];
    S[k] = t;
  }
}
// dg[k] += scale * S_k / n_k ;  dv[o,i,k] += scale * g_k / n_k * (dW[o,k,i] - v[o,i,k] * S_k / n_k^2)
__global__ void wnorm_bwd_apply_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                       const float* __restrict__ g, const float* __restrict__ norm,
                                       const float* __restrict__ S, float scale, float* __restrict__ dv,
                                       float* __restrict__ dg, int H, int I, int K) {
  const int64_t n = int64_t(H) * I * K;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < n; idx += int64_t(gridDim.x) * blockDim.x) {
    const int k = idx % K;
    const int i = (idx / K) % I;
    const int o = idx / (int64_t(K) * I);
    const float nk = norm[k];
    const float d = dw[(int64_t(o) * K + k) * I + i];
    dv[idx] += scale * (g[k] / nk) * (d - v[idx] * S[k] / (nk * nk));
    if (idx < K) dg[idx] += scale * S[idx] / norm[idx];
  }
}

// Coalesced forms of the two kernels above.  dW arrives as [H][K][I] (taps outer), v and dv are [H][I][K] (taps
// inner): one block per output channel o stages dW[o] transposed in shared memory (padded rows), so both global
// streams are read / written along their contiguous axis.
float* posconv_partial_buffer();
constexpr int WN_BLOCKS = 296;     // <= PN_BLOCKS (shared partial buffer), 2 per SM

__global__ void __launch_bounds__(256) wnorm_bwd_reduce_tiled_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                                     float* __restrict__ partial, int H, int I, int K) {
  extern __shared__ float dwt[];                          // [I][K + 1]
  __shared__ float red[256];
  const int k = threadIdx.x % K, hl = threadIdx.x / K, nhl = 256 / K;
  float s = 0.f;
  for (int o = blockIdx.x; o < H; o += gridDim.x) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < K * I; idx += 256) {       // dW[o][kk][i], i fastest
      const int i = idx % I, kk = idx / I;
      dwt[i * (K + 1) + kk] = dw[int64_t(o) * K * I + idx];
    }
    __syncthreads();
    for (int i = hl; i < I; i += nhl) s = fmaf(dwt[i * (K + 1) + k], v[(int64_t(o) * I + i) * K + k], s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (hl == 0) {
    for (int j = 1; j < nhl; ++j) s += red[j * K + k];
    partial[blockIdx.x * K + k] = s;
  }
}
__global__ void __launch_bounds__(256) posconv_partial_sum_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                                  int nblocks, int K) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // one warp per tap
  if (k >= K) return;
  float s = 0.f;
  for (int b = threadIdx.x & 31; b < nblocks; b += 32) s += partial[b * K + k];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) out[k] = s;
}
__global__ void __launch_bounds__(256) wnorm_bwd_apply_tiled_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                                    const float* __restrict__ g, const float* __restrict__ norm,
                                                                    const float* __restrict__ S, float scale,
                                                                    float* __restrict__ dv, float* __restrict__ dg, int H, int I,
                                                                    int K) {
  extern __shared__ float dwt[];                          // [I][K + 1]
  const int o = blockIdx.x;
  for (int idx = threadIdx.x; idx < K * I; idx += 256) {
    const int i = idx % I, kk = idx / I;
    dwt[i * (K + 1) + kk] = dw[int64_t(o) * K * I + idx];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < I * K; idx += 256) {       // dv[o][i][k], k fastest
    const int k = idx % K, i = idx / K;
    const float nk = norm[k];
    const int64_t at = (int64_t(o) * I + i) * K + k;
    dv[at] += scale * (g[k] / nk) * (dwt[i * (K + 1) + k] - v[at] * S[k] / (nk * nk));
  }
  if (o == 0 && int(threadIdx.x) < K) dg[threadIdx.x] += scale * S[threadIdx.x] / norm[threadIdx.x];
}

}  // namespace w2v2

using namespace w2v2;

extern "C" {

int w2v2_gelu_fwd(const void* x, int x_dtype, void* y, int y_dtype, void* x16_copy, int64_t n, void* stream) {
  W2V2_REQUIRE(n % 4 == 0, "w2v2_gelu_fwd: n=%lld must be a multiple of 4", (long long)n);
  if (n == 0) return 0;
  const int grid = pgrid(n / 4, 256, 8);
  cudaStream_t st = (cudaStream_t)stream;
  __half* c = (__half*)x16_copy;
  if (x_dtype == 1 && y_dtype == 1) launch_k(gelu_fwd_kernel<true, true>, dim3(grid), dim3(256), 0, st, 1, x, y, c, n);
  else if (x_dtype == 1) launch_k(gelu_fwd_kernel<true, false>, dim3(grid), dim3(256), 0, st, 1, x, y, c, n);
  else if (y_dtype == 1) launch_k(gelu_fwd_kernel<false, true>, dim3(grid), dim3(256), 0, st, 1, x, y, c, n);
  else launch_k(gelu_fwd_kernel<false, false>, dim3(grid), dim3(256), 0, st, 1, x, y, c, n);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_posconv_im2col(const void* x16, void* xg16, int B, int T, int H, int groups, int K, int g, void* stream) {
  const int I = H / groups;
  W2V2_REQUIRE(I % 8 == 0 && g >= 0 && g < groups, "w2v2_posconv_im2col: bad group %d / channels %d", g, I);
  const int64_t n = int64_t(B) * T * K * (I / 8);
  posconv_im2col_kernel<<<pgrid(n, 256, 16), 256, 0, (cudaStream_t)stream>>>((const __half*)x16, (__half*)xg16, B, T, H, I,
                                                                              K, g);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_weight_norm_bwd(const float* dw_hki, const float* v, const float* g, float* scratch_2k, float scale, float* dv,
                         float* dg, int H, int I, int K, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  float* norm = scratch_2k;
  float* S = scratch_2k + K;
  posconv_tap_norms(v, norm, H, I, K, stream);      // ||v[:,:,k]||
  const size_t smem = size_t(I) * (K + 1) * sizeof(float);
  if (K <= 256 && 256 % K == 0 && smem <= 48 * 1024) {
    float* partial = posconv_partial_buffer();      // the norm kernels are done with it (same stream)
    const int nb = H < WN_BLOCKS ? H : WN_BLOCKS;
    wnorm_bwd_reduce_tiled_kernel<<<nb, 256, smem, stream>>>(dw_hki, v, partial, H, I, K);
    posconv_partial_sum_kernel<<<(K + 7) / 8, 256, 0, stream>>>(partial, S, nb, K);
    wnorm_bwd_apply_tiled_kernel<<<H, 256, smem, stream>>>(dw_hki, v, g, norm, S, scale, dv, dg, H, I, K);
  } else {
    wnorm_bwd_reduce_kernel<<<K, 256, 0, stream>>>(dw_hki, v, S, H, I, K);
    const int64_t n = int64_t(H) * I * K;
    wnorm_bwd_apply_kernel<<<pgrid(n, 256, 8), 256, 0, stream>>>(dw_hki, v, g, norm, S, scale, dv, dg, H, I, K);
  }
  count_launches(5);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
