// Small elementwise helpers of the backward pass (gradient plumbing between the tensor-core kernels).
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int device_sm_count();
static inline int mgrid(int64_t n, int per_block, int per_sm) {
  int64_t g = (n + per_block - 1) / per_block;
  const int64_t cap = int64_t(device_sm_count()) * per_sm;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

// out16 / out32 = a + b   (b may be NULL); n % 4 == 0
__global__ void add2_cast_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o32,
                                 __half* __restrict__ o16, int64_t n) {
  int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x * 4;
  for (; i < n; i += stride) {
    float4 x = *reinterpret_cast<const float4*>(a + i);
    if (b != nullptr) {
      const float4 y = *reinterpret_cast<const float4*>(b + i);
      x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    }
    if (o32 != nullptr) *reinterpret_cast<float4*>(o32 + i) = x;
    if (o16 != nullptr) {
      uint2 q;
      q.x = pack_half2(x.x, x.y);
      q.y = pack_half2(x.z, x.w);
      *reinterpret_cast<uint2*>(o16 + i) = q;
    }
  }
}

// y16[r, 0:ldy] = half(x[r, 0:cols] * scale), zero filled to ldy
__global__ void cast_f16_rows_kernel(const float* __restrict__ x, int64_t ldx, __half* __restrict__ y, int64_t ldy,
                                     int64_t rows, int cols, float scale) {
  const int64_t n = rows * ldy;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / ldy;
    const int c = int(i % ldy);
    y[i] = __float2half_rn(c < cols ? x[r * ldx + c] * scale : 0.f);
  }
}

__global__ void scale_f32_kernel(float* __restrict__ x, int64_t n, float s) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) x[i] *= s;
}

// y = x * s (out of place: autograd's gradient tensors are not ours to overwrite); 16-byte vectors + scalar tail
__global__ void scale_copy_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float s, int vec) {
  pdl_trigger();
  pdl_wait();
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = int64_t(gridDim.x) * blockDim.x;
  const int64_t n4 = vec ? n / 4 : 0;
  for (int64_t i = tid; i < n4; i += nth) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    reinterpret_cast<float4*>(y)[i] = v;
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += nth) y[i] = x[i] * s;
}

// dlogits (f32) = (prob - onehot) * coef
__global__ void softmax_ce_bwd_f32_kernel(const float* __restrict__ prob, const int64_t* __restrict__ labels,
                                          const float* __restrict__ coef_ptr, float coef, float* __restrict__ dl, int S) {
  const int b = blockIdx.x;
  const int label = int(labels[b]);
  const float c = coef * (coef_ptr != nullptr ? coef_ptr[0] : 1.0f);
  for (int i = threadIdx.x; i < S; i += blockDim.x)
    dl[int64_t(b) * S + i] = (prob[int64_t(b) * S + i] - (i == label ? 1.0f : 0.0f)) * c;
}

// ---- gradient entry of an autograd Function: loss-scaled working copy whose magnitude is made safe for fp16 operands ----
// Stage 1: amax = max |x| as a bit pattern (non-negative floats order like their bits; NaN / Inf give a pattern above every
// finite one and disable the normalisation: they must propagate so that an outer GradScaler sees them).
__global__ void amax_f32_kernel(const float* __restrict__ x, int64_t n, unsigned int* __restrict__ amax_bits) {
  pdl_trigger();
  pdl_wait();
  unsigned int m = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    m = max(m, __float_as_uint(x[i]) & 0x7fffffffu);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m != 0) atomicMax(amax_bits, m);
}
// The power of two k that brings a = amax * s back to `mid` when a left the window [lo, hi]; 1 inside the window (the
// normal case: nothing changes), for an all-zero or a non-finite input.
__device__ __forceinline__ float entry_norm_factor(unsigned int amax_bits, float s, float lo, float hi, float mid) {
  const float a = __uint_as_float(amax_bits) * s;
  if (!(a > 0.f) || !(a < 3.0e38f) || (a >= lo && a <= hi)) return 1.0f;
  return exp2f(floorf(log2f(mid / a)));
}
// Stage 2: y = x * s * k; block 0 publishes 1 / k for the Function's outputs.
__global__ void scale_copy_norm_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float s,
                                           const unsigned int* __restrict__ amax_bits, float lo, float hi, float mid,
                                           float* __restrict__ inv_out, int vec) {
  pdl_trigger();
  pdl_wait();
  const float k = entry_norm_factor(*amax_bits, s, lo, hi, mid);
  if (blockIdx.x == 0 && threadIdx.x == 0) *inv_out = 1.0f / k;
  const float sk = s * k;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = int64_t(gridDim.x) * blockDim.x;
  const int64_t n4 = vec ? n / 4 : 0;
  for (int64_t i = tid; i < n4; i += nth) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    v.x *= sk; v.y *= sk; v.z *= sk; v.w *= sk;
    reinterpret_cast<float4*>(y)[i] = v;
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += nth) y[i] = x[i] * sk;
}
// x *= s * dev[0]  (leaving a Function: undo the loss scale and the device-chosen normalisation)
__global__ void scale_f32_dev_kernel(float* __restrict__ x, int64_t n, float s, const float* __restrict__ dev) {
  pdl_trigger();
  pdl_wait();
  const float sk = s * dev[0];
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) x[i] *= sk;
}

}  // namespace w2v2

using namespace w2v2;

extern "C" {

int w2v2_add2_cast(const float* a, const float* b, float* out32, void* out16, int64_t n, void* stream) {
  W2V2_REQUIRE(n % 4 == 0, "w2v2_add2_cast: n=%lld must be a multiple of 4", (long long)n);
  if (n == 0) return 0;
  add2_cast_kernel<<<mgrid(n / 4, 256, 8), 256, 0, (cudaStream_t)stream>>>(a, b, out32, (__half*)out16, n);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_cast_f16_rows(const float* x, int64_t ldx, void* y16, int64_t ldy, int64_t rows, int cols, float scale,
                       void* stream) {
  W2V2_REQUIRE(ldy >= cols, "w2v2_cast_f16_rows: ldy < cols");
  if (rows == 0) return 0;
  cast_f16_rows_kernel<<<mgrid(rows * ldy, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, ldx, (__half*)y16, ldy, rows, cols, scale);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_scale_f32(float* x, int64_t n, float s, void* stream) {
  if (n == 0) return 0;
  W2V2_CHECK_CUDA(launch_k(scale_f32_kernel, dim3(mgrid(n, 256, 8)), dim3(256), 0, (cudaStream_t)stream, 1, x, n, s));
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_scale_copy_f32(const float* x, float* y, int64_t n, float s, void* stream) {
  if (n == 0) return 0;
  const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  W2V2_CHECK_CUDA(launch_k(scale_copy_f32_kernel, dim3(mgrid((n + 3) / 4, 256, 8)), dim3(256), 0, (cudaStream_t)stream, 1, x, y,
                           n, s, vec));
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_grad_entry_scale(const float* x, float* y, int64_t n, float s, float lo, float hi, float mid, float* state,
                          void* stream) {
  W2V2_REQUIRE(state != nullptr && lo > 0.f && hi > lo && mid >= lo && mid <= hi, "w2v2_grad_entry_scale: bad window / state");
  cudaStream_t st = (cudaStream_t)stream;
  W2V2_CHECK_CUDA(cudaMemsetAsync(state, 0, sizeof(float), st));
  if (n > 0) W2V2_CHECK_CUDA(launch_k(amax_f32_kernel, dim3(mgrid(n, 256, 4)), dim3(256), 0, st, 1, x, n, reinterpret_cast<unsigned int*>(state)));
  const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  W2V2_CHECK_CUDA(launch_k(scale_copy_norm_f32_kernel, dim3(mgrid((n + 3) / 4 + 1, 256, 8)), dim3(256), 0, st, 1, x, y, n, s,
                           reinterpret_cast<const unsigned int*>(state), lo, hi, mid, state + 1, vec));
  count_launches(2);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_scale_f32_dev(float* x, int64_t n, float s, const float* dev_scale, void* stream) {
  W2V2_REQUIRE(dev_scale != nullptr, "w2v2_scale_f32_dev: dev_scale is required");
  if (n == 0) return 0;
  W2V2_CHECK_CUDA(launch_k(scale_f32_dev_kernel, dim3(mgrid(n, 256, 8)), dim3(256), 0, (cudaStream_t)stream, 1, x, n, s, dev_scale));
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_softmax_ce_bwd_f32(const float* prob, const int64_t* labels, const float* dloss, float coef, float* dlogits,
                            int B, int S, void* stream) {
  softmax_ce_bwd_f32_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(prob, labels, dloss, coef, dlogits, S);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
