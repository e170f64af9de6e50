// The step in front of the hot path: per-utterance standardisation of the waveform,
// InputNormalizer2D.normalize(channel_wise=False) of R:src/data/preprocess/input_normalisation.py:53-67:
//     y = (x - mean(x)) / (std_unbiased(x) + 1e-5)
// done on the device, optionally straight from 16-bit PCM (x = pcm / 32768, what torchaudio.load returns for a
// 16-bit wav): the host then uploads half the bytes and runs no arithmetic.  One block per utterance, two passes
// over a row that lives in L2 after the first; sums in double.
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

template <bool PCM16>
__global__ void __launch_bounds__(1024) normalize_wav_kernel(const void* __restrict__ in_, float* __restrict__ out,
                                                             float* __restrict__ mean_out, float* __restrict__ std_out, int N) {
  __shared__ double red[2][32];
  __shared__ float stat[2];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto load = [&](int i) -> float {
    if constexpr (PCM16) return float(static_cast<const int16_t*>(in_)[int64_t(b) * N + i]) * (1.0f / 32768.0f);
    else return static_cast<const float*>(in_)[int64_t(b) * N + i];
  };
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double v = load(i);
    s += v;
    q += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) { red[0][warp] = s; red[1][warp] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) { ts += red[0][w]; tq += red[1][w]; }
    const double mean = ts / N;
    double var = N > 1 ? (tq - ts * mean) / double(N - 1) : 0.0;       // unbiased, like torch.std_mean
    if (var < 0.0) var = 0.0;
    stat[0] = float(mean);
    stat[1] = float(sqrt(var));
    if (mean_out != nullptr) mean_out[b] = stat[0];
    if (std_out != nullptr) std_out[b] = stat[1];
  }
  __syncthreads();
  const float mean = stat[0], inv = 1.0f / (stat[1] + 1e-5f);
  for (int i = threadIdx.x; i < N; i += blockDim.x) out[int64_t(b) * N + i] = (load(i) - mean) * inv;
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_normalize_wav(const void* in, int in_dtype, float* out, float* mean, float* stdv, int B, int N,
                                  void* stream) {
  W2V2_REQUIRE(B >= 1 && N >= 1, "w2v2_normalize_wav: empty batch");
  W2V2_REQUIRE(in_dtype == 0 || in_dtype == 1, "w2v2_normalize_wav: in_dtype 0 = int16 PCM, 1 = float32");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_dtype == 0) normalize_wav_kernel<true><<<B, 1024, 0, st>>>(in, out, mean, stdv, N);
  else normalize_wav_kernel<false><<<B, 1024, 0, st>>>(in, out, mean, stdv, N);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}
