// Train-mode regularisation of the wav2vec2 encoder (HF:433, 693, 546-600, 1280-1324):
// counter-based dropout masks (regenerated in the backward from (seed, element index), never stored)
// and the SpecAugment time mask.  mask(seed, idx): one 64-bit mix per element PAIR, 16 random bits per
// element, keep <=> bits >= round(p * 65536).  tests/test_gpu_regularise.py holds the numpy replica.
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int device_sm_count();
static inline int rgrid(int64_t n, int per_block, int per_sm) {
  int64_t g = (n + per_block - 1) / per_block;
  const int64_t cap = int64_t(device_sm_count()) * per_sm;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

// y = keep ? (x + bias[col]) / (1 - p) : 0     (x, y f32 or f16; optional second f16 output)
template <bool F32>
__global__ void dropout_kernel(const void* __restrict__ x_, const float* __restrict__ bias, int H, void* __restrict__ y_,
                               __half* __restrict__ y16, int64_t n, uint32_t thr, float inv_keep, uint64_t seed) {
  pdl_trigger();
  pdl_wait();
  // two elements (one hash) per thread iteration; n % 2 == 0
  const DropKeys dkeys = drop_keys(seed);
  for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 2; i < n;
       i += int64_t(gridDim.x) * blockDim.x * 2) {
    const uint32_t h = dropout_hash_k(dkeys, uint64_t(i >> 1));
    float a, b;
    if constexpr (F32) {
      const float2 v = *reinterpret_cast<const float2*>(static_cast<const float*>(x_) + i);
      a = v.x; b = v.y;
    } else {
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(static_cast<const __half*>(x_) + i));
      a = v.x; b = v.y;
    }
    if (bias != nullptr) {
      const int c = int(i % H);
      a += bias[c];
      b += bias[c + 1];
    }
    a = (h & 0xffffu) >= thr ? a * inv_keep : 0.f;
    b = (h >> 16) >= thr ? b * inv_keep : 0.f;
    if constexpr (F32) *reinterpret_cast<float2*>(static_cast<float*>(y_) + i) = make_float2(a, b);
    else *reinterpret_cast<uint32_t*>(static_cast<__half*>(y_) + i) = pack_half2(a, b);
    if (y16 != nullptr) *reinterpret_cast<uint32_t*>(y16 + i) = pack_half2(a, b);
  }
}

// SpecAugment along the feature axis (HF:1312-1322): columns with mask[b, c] != 0 are zeroed in every frame of utterance b.
// Only masked columns are written (a few percent of the tensor): the pass reads the mask and stores zeros.
__global__ void feature_mask_kernel(float4* __restrict__ h, const uint8_t* __restrict__ mask, int T, int H4, int64_t n4) {
  const int64_t per_utt = int64_t(T) * H4;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = i / per_utt;
    const int c4 = int(i % H4);
    const uint32_t m = *reinterpret_cast<const uint32_t*>(mask + (b * H4 + c4) * 4);
    if (m == 0) continue;
    float4 v = h[i];
    if (m & 0x000000ffu) v.x = 0.f;
    if (m & 0x0000ff00u) v.y = 0.f;
    if (m & 0x00ff0000u) v.z = 0.f;
    if (m & 0xff000000u) v.w = 0.f;
    h[i] = v;
  }
}

// SpecAugment: rows with mask != 0 are overwritten by the learned embedding (HF:1301-1310)
__global__ void time_mask_apply_kernel(float* __restrict__ h, const uint8_t* __restrict__ mask,
                                       const float* __restrict__ embed, int64_t rows, int H) {
  const int64_t n = rows * H;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / H;
    if (mask[r]) h[i] = embed[i % H];
  }
}
// backward: d_embed += sum of masked rows of dh ; dh[masked rows] = 0
// A block owns a contiguous chunk of rows, a thread four columns: the mask byte of a row is one broadcast load, only
// the ~10 % masked rows touch dh, and the column sums leave the block as one atomic per column (two blocks per SM:
// ~0.2 M atomics instead of the ~0.9 M of a thread-per-(column, row-lane) layout, 58 us -> ~10 us at cfg1).
__global__ void __launch_bounds__(256) time_mask_bwd_kernel(float* __restrict__ dh, const uint8_t* __restrict__ mask,
                                                            float* __restrict__ dembed, int64_t rows, int H, float scale) {
  const int c = threadIdx.x * 4;
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = int64_t(blockIdx.x) * per, r1 = (r0 + per < rows) ? r0 + per : rows;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  // the chunk's mask bytes in one parallel load (a serial walk pays a dependent global load per row)
  __shared__ uint8_t flags[256];
  for (int64_t rb = r0; rb < r1; rb += 256) {
    __syncthreads();
    if (rb + threadIdx.x < r1) flags[threadIdx.x] = mask[rb + threadIdx.x];
    __syncthreads();
    const int64_t re = (rb + 256 < r1) ? rb + 256 : r1;
    for (int64_t r = rb; r < re; ++r) {
      if (flags[r - rb] == 0) continue;              // block-uniform
      for (int cc = c; cc < H; cc += blockDim.x * 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (cc + j < H) {
            s[j] += dh[r * H + cc + j];
            dh[r * H + cc + j] = 0.f;
          }
        }
      }
    }
  }
  // (a thread visits columns c, c + 4 blockDim, ...; with H <= 4 blockDim -- every model here -- that is one group)
  if (H <= int(blockDim.x) * 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < H && s[j] != 0.f) atomicAdd(dembed + c + j, s[j] * scale);
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" {

int w2v2_dropout(const void* x, int dtype, const float* bias, int H, void* y, void* y16, int64_t n, float p,
                 uint64_t seed, void* stream) {
  W2V2_REQUIRE(n % 2 == 0 && (bias == nullptr || H % 2 == 0), "w2v2_dropout: n (and H) must be even");
  W2V2_REQUIRE(p >= 0.f && p < 1.f, "w2v2_dropout: p=%f out of [0,1)", p);
  if (n == 0) return 0;
  const uint32_t thr = uint32_t(p * 65536.0f + 0.5f);
  const float inv_keep = 1.0f / (1.0f - float(thr) / 65536.0f);
  const int grid = rgrid(n / 2, 256, 8);
  if (dtype == 1) launch_k(dropout_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 1, x, bias, H, y, (__half*)y16, n, thr, inv_keep, seed);
  else launch_k(dropout_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 1, x, bias, H, y, (__half*)y16, n, thr, inv_keep, seed);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// SpecAugment along the feature axis (HF:1312-1322, `mask_feature_prob`): h[b, t, c] = 0 where mask[b, c]; the same
// call zeroes the gradient of those columns in the backward.  One float4 per thread, mask bytes read as one word.
int w2v2_feature_mask(float* h, const uint8_t* mask, int B, int T, int H, void* stream) {
  W2V2_REQUIRE(H % 4 == 0, "w2v2_feature_mask: H=%d must be a multiple of 4", H);
  const int64_t n4 = int64_t(B) * T * (H / 4);
  if (n4 == 0) return 0;
  feature_mask_kernel<<<rgrid(n4, 256, 8), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(h), mask, T, H / 4, n4);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_time_mask_apply(float* h, const uint8_t* mask, const float* embed, int64_t rows, int H, void* stream) {
  if (rows == 0) return 0;
  time_mask_apply_kernel<<<rgrid(rows * H, 256, 8), 256, 0, (cudaStream_t)stream>>>(h, mask, embed, rows, H);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_time_mask_bwd(float* dh, const uint8_t* mask, float* dembed, int64_t rows, int H, float scale, void* stream) {
  if (rows == 0) return 0;
  W2V2_REQUIRE(H <= 1024, "w2v2_time_mask_bwd: H=%d exceeds 1024", H);
  const int64_t blocks = rows < 296 ? rows : 296;
  time_mask_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dh, mask, dembed, rows, H, scale);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
