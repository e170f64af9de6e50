// Batched weight preparation: after an optimizer step every trainable matrix needs its fp16 copy
// (the W operand of the forward GEMMs) and its transposed fp16 copy (the W^T operand of the
// data-gradient GEMMs), some with a folded scalar (the d^-0.5 of the q projection, HF:528), and the
// fused bias vectors need re-assembling.  One launch walks a device-resident job table instead of
// ~200 small cast / transpose / cat launches.
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

static_assert(sizeof(w2v2_prep_job) == 64, "w2v2_prep_job is a 64-byte record");

// A bounded number of resident blocks walks the tiles (3 per SM): the pass is HBM-bound and shares the machine
// with the next step's CNN forward (see trainer.py), whose GEMM CTAs must fit next to it.
__global__ void __launch_bounds__(256) prepare_weights_kernel(const w2v2_prep_job* __restrict__ jobs, int njobs,
                                                              long long total_tiles) {
  __shared__ float tile[32][33];
  __shared__ int sjob;
  for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
  __syncthreads();                       // the previous tile's transposed writes are done with `tile`
  // locate the job of this 32x32 tile (tile_begin is an ascending prefix sum)
  if (threadIdx.x == 0) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
    }
    sjob = lo;
  }
  __syncthreads();
  const w2v2_prep_job j = jobs[sjob];
  const int tiles_c = (j.C + 31) / 32;
  const int local = int(t - j.tile_begin);
  const int r0 = (local / tiles_c) * 32, c0 = (local % tiles_c) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const float* src = static_cast<const float*>(j.src);
  __half* d16 = static_cast<__half*>(j.dst16);
  __half* dT = static_cast<__half*>(j.dstT16);
  float* d32 = static_cast<float*>(j.dst32);
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < j.R && c < j.C) {
      const float raw = src[int64_t(r) * j.C + c];
      v = raw * j.scale;
      if (d16 != nullptr) d16[int64_t(r) * j.ld + c] = __float2half_rn(v);
      if (d32 != nullptr) d32[int64_t(r) * j.ld + c] = v;
      v = raw * j.scale_t;
    }
    tile[i][tx] = v;
  }
  if (dT == nullptr) continue;           // block-uniform: every thread reads the same job record
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < j.C && r < j.R) dT[int64_t(c) * j.ldt + r] = __float2half_rn(tile[tx][i]);
  }
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_prepare_weights(const w2v2_prep_job* jobs_dev, int njobs, int64_t total_tiles, void* stream) {
  W2V2_REQUIRE(njobs >= 0 && total_tiles >= 0 && total_tiles < (int64_t(1) << 31), "w2v2_prepare_weights: bad job table");
  if (njobs == 0 || total_tiles == 0) return 0;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t cap = int64_t(sms > 0 ? sms : 148) * 3;
  prepare_weights_kernel<<<unsigned(total_tiles < cap ? total_tiles : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      jobs_dev, njobs, total_tiles);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}
