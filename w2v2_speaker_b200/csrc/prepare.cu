// Batched weight preparation: after an optimizer step every trainable matrix needs its fp16 copy
// (the W operand of the forward GEMMs) and its transposed fp16 copy (the W^T operand of the
// data-gradient GEMMs), some with a folded scalar (the d^-0.5 of the q projection, HF:528), and the
// fused bias vectors need re-assembling.  One launch walks a device-resident job table instead of
// ~200 small cast / transpose / cat launches.
#include <stdlib.h>

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

// Second form, the default since round 2 (W2V2_PREP_V2=0 restores the first; validated on B200: outputs bit-identical,
// tests/test_gpu_training.py::test_inplace_weight_refresh_equals_rebuild, train step 11.07 -> 10.89 ms).
// The ncu launch list puts the kernel above at 608 us for ~0.7 GB of traffic (18 % of the copy bandwidth): every
// 32x32 tile pays a binary search through the job table in GLOBAL memory by one thread (8 dependent loads) plus two
// block barriers, with 4 KB in flight per block.  Here the job table is staged in shared memory once per block, a
// block takes PV2_GROUP consecutive tiles per iteration (four searches in parallel, all loads of the group issued
// before the first use), and the arithmetic -- the same multiplies and round-to-nearest conversions -- is unchanged,
// so the outputs are bit-identical.
constexpr int PV2_GROUP = 4;
constexpr int PV2_MAX_JOBS = 512;

__global__ void __launch_bounds__(256) prepare_weights_v2_kernel(const w2v2_prep_job* __restrict__ jobs, int njobs,
                                                                 long long total_tiles) {
  extern __shared__ __align__(16) unsigned char pv2_smem[];
  w2v2_prep_job* sjobs = reinterpret_cast<w2v2_prep_job*>(pv2_smem);
  float (*tile)[32][33] = reinterpret_cast<float (*)[32][33]>(pv2_smem + sizeof(w2v2_prep_job) * njobs);
  __shared__ int sjob[PV2_GROUP];
  {
    const uint4* g = reinterpret_cast<const uint4*>(jobs);
    uint4* d = reinterpret_cast<uint4*>(sjobs);
    for (int i = threadIdx.x; i < njobs * 4; i += blockDim.x) d[i] = g[i];      // 64-byte records = 4 x uint4
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (long long t0 = (long long)blockIdx.x * PV2_GROUP; t0 < total_tiles; t0 += (long long)gridDim.x * PV2_GROUP) {
    __syncthreads();                     // job table staged / previous group's transposed writes done with `tile`
    if (threadIdx.x < PV2_GROUP) {
      const long long t = t0 + threadIdx.x;
      int lo = 0, hi = njobs - 1;
      if (t < total_tiles) {
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (sjobs[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
        }
      }
      sjob[threadIdx.x] = t < total_tiles ? lo : -1;
    }
    __syncthreads();
    float raw[PV2_GROUP][4];
    // load phase: every element of the group in flight before anything is used
#pragma unroll
    for (int g = 0; g < PV2_GROUP; ++g) {
      const int ji = sjob[g];
#pragma unroll
      for (int k = 0; k < 4; ++k) raw[g][k] = 0.f;
      if (ji < 0) continue;
      const w2v2_prep_job& j = sjobs[ji];
      const int tiles_c = (j.C + 31) / 32;
      const int local = int(t0 + g - j.tile_begin);
      const int r0 = (local / tiles_c) * 32, c = (local % tiles_c) * 32 + tx;
      const float* src = static_cast<const float*>(j.src);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k;
        if (r < j.R && c < j.C) raw[g][k] = __ldg(src + int64_t(r) * j.C + c);
      }
    }
    // plain copies + staging for the transposed ones
    bool any_t = false;
#pragma unroll
    for (int g = 0; g < PV2_GROUP; ++g) {
      const int ji = sjob[g];
      if (ji < 0) continue;
      const w2v2_prep_job& j = sjobs[ji];
      const int tiles_c = (j.C + 31) / 32;
      const int local = int(t0 + g - j.tile_begin);
      const int r0 = (local / tiles_c) * 32, c = (local % tiles_c) * 32 + tx;
      __half* d16 = static_cast<__half*>(j.dst16);
      float* d32 = static_cast<float*>(j.dst32);
      any_t |= j.dstT16 != nullptr;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = ty + 8 * k, r = r0 + i;
        float vt = 0.f;
        if (r < j.R && c < j.C) {
          const float v = raw[g][k] * j.scale;
          if (d16 != nullptr) d16[int64_t(r) * j.ld + c] = __float2half_rn(v);
          if (d32 != nullptr) d32[int64_t(r) * j.ld + c] = v;
          vt = raw[g][k] * j.scale_t;
        }
        tile[g][i][tx] = vt;
      }
    }
    if (!any_t) continue;                // block-uniform: derived from shared state only
    __syncthreads();
#pragma unroll
    for (int g = 0; g < PV2_GROUP; ++g) {
      const int ji = sjob[g];
      if (ji < 0) continue;
      const w2v2_prep_job& j = sjobs[ji];
      __half* dT = static_cast<__half*>(j.dstT16);
      if (dT == nullptr) continue;
      const int tiles_c = (j.C + 31) / 32;
      const int local = int(t0 + g - j.tile_begin);
      const int r0 = (local / tiles_c) * 32, c0 = (local % tiles_c) * 32;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = ty + 8 * k;
        const int c = c0 + i, r = r0 + tx;
        if (c < j.C && r < j.R) dT[int64_t(c) * j.ldt + r] = __float2half_rn(tile[g][tx][i]);
      }
    }
  }
}

// Third form (default; W2V2_PREP_V3=0 restores the second).  The second form moves ~0.76 GB in 337 us -- a third of the
// copy bandwidth: 4-byte loads, 2-byte stores (64-byte segments), two block barriers and a one-thread search per group of
// four 4 KB tiles.  Here a block owns one 64 x 64 tile per iteration: every thread issues four 16-byte loads (a warp reads
// two full 256-byte row segments) before anything is used, the plain copies leave as 8-byte (fp16) / 16-byte (fp32)
// stores, the transposed copy goes through an fp32 shared-memory tile and leaves as 16-byte stores (8 consecutive source
// rows of one column), the job search runs in every thread on the shared-memory copy of the table (no barrier), and four
// blocks per SM keep 64 KB of loads in flight per SM.  Same multiplies, same round-to-nearest conversions: bit-identical
// outputs.  tile_begin counts 64 x 64 tiles for this form (w2v2_prepare_tile_edge() tells the host which).
constexpr int PV3_EDGE = 64;
constexpr int PV3_PITCH = 65;

__device__ __forceinline__ uint32_t pv3_pack(float a, float b) {
  const __half2 h = __halves2half2(__float2half_rn(a), __float2half_rn(b));
  return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(256, 4) prepare_weights_v3_kernel(const w2v2_prep_job* __restrict__ jobs, int njobs,
                                                                    long long total_tiles) {
  extern __shared__ __align__(16) unsigned char pv3_smem[];
  float* tile = reinterpret_cast<float*>(pv3_smem);                                       // [64][65]
  w2v2_prep_job* sjobs = reinterpret_cast<w2v2_prep_job*>(pv3_smem + sizeof(float) * PV3_EDGE * PV3_PITCH + 16);
  {
    const uint4* g = reinterpret_cast<const uint4*>(jobs);
    uint4* d = reinterpret_cast<uint4*>(sjobs);
    for (int i = threadIdx.x; i < njobs * 4; i += blockDim.x) d[i] = g[i];      // 64-byte records = 4 x uint4
  }
  __syncthreads();
  const int c4 = threadIdx.x & 15, rr = threadIdx.x >> 4;      // load map: 16 float4 per row, 16 rows per pass
  const int r8 = threadIdx.x & 7, cc = threadIdx.x >> 3;       // transposed-store map: 8 chunks of 8 rows, 32 columns per pass
  for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (sjobs[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
    }
    const w2v2_prep_job& j = sjobs[lo];
    const int R = j.R, C = j.C;
    const int tiles_c = (C + PV3_EDGE - 1) / PV3_EDGE;
    const int local = int(t - j.tile_begin);
    const int r0 = (local / tiles_c) * PV3_EDGE, c0 = (local % tiles_c) * PV3_EDGE;
    const float* src = static_cast<const float*>(j.src);
    __half* d16 = static_cast<__half*>(j.dst16);
    float* d32 = static_cast<float*>(j.dst32);
    __half* dT = static_cast<__half*>(j.dstT16);
    const bool vec = (C & 3) == 0 && (j.ld & 3) == 0;            // 16-byte source rows / 8-byte fp16 destination rows
    const int c = c0 + 4 * c4;
    float4 raw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + rr + 16 * k;
      raw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < R) {
        const float* s = src + int64_t(r) * C + c;
        if (vec && c + 3 < C) raw[k] = __ldg(reinterpret_cast<const float4*>(s));
        else {
          if (c < C) raw[k].x = __ldg(s);
          if (c + 1 < C) raw[k].y = __ldg(s + 1);
          if (c + 2 < C) raw[k].z = __ldg(s + 2);
          if (c + 3 < C) raw[k].w = __ldg(s + 3);
        }
      }
    }
    if (dT != nullptr) __syncthreads();          // the previous tile's transposed reads are done with `tile` (block-uniform)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = rr + 16 * k, r = r0 + i;
      if (r < R && c < C) {
        const float v0 = raw[k].x * j.scale, v1 = raw[k].y * j.scale, v2 = raw[k].z * j.scale, v3 = raw[k].w * j.scale;
        if (vec && c + 3 < C) {
          if (d16 != nullptr) *reinterpret_cast<uint2*>(d16 + int64_t(r) * j.ld + c) = make_uint2(pv3_pack(v0, v1), pv3_pack(v2, v3));
          if (d32 != nullptr) *reinterpret_cast<float4*>(d32 + int64_t(r) * j.ld + c) = make_float4(v0, v1, v2, v3);
        } else {
          const float vv[4] = {v0, v1, v2, v3};
          for (int e = 0; e < 4 && c + e < C; ++e) {
            if (d16 != nullptr) d16[int64_t(r) * j.ld + c + e] = __float2half_rn(vv[e]);
            if (d32 != nullptr) d32[int64_t(r) * j.ld + c + e] = vv[e];
          }
        }
      }
      if (dT != nullptr) {
        float* trow = tile + i * PV3_PITCH + 4 * c4;
        trow[0] = raw[k].x * j.scale_t;
        trow[1] = raw[k].y * j.scale_t;
        trow[2] = raw[k].z * j.scale_t;
        trow[3] = raw[k].w * j.scale_t;
      }
    }
    if (dT == nullptr) continue;                 // block-uniform
    __syncthreads();
    const bool vec_t = (j.ldt & 7) == 0 && (reinterpret_cast<uintptr_t>(dT) & 15) == 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int ci = cc + 32 * k, col = c0 + ci, row = r0 + 8 * r8;
      if (col >= C || row >= R) continue;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = tile[(8 * r8 + e) * PV3_PITCH + ci];
      __half* dst = dT + int64_t(col) * j.ldt + row;
      if (vec_t && row + 7 < R) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(pv3_pack(v[0], v[1]), pv3_pack(v[2], v[3]), pv3_pack(v[4], v[5]), pv3_pack(v[6], v[7]));
      } else {
        for (int e = 0; e < 8 && row + e < R; ++e) dst[e] = __float2half_rn(v[e]);
      }
    }
  }
}

static bool prep_v3() {
  static const bool on = []() { const char* e = getenv("W2V2_PREP_V3"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}

}  // namespace w2v2

using namespace w2v2;

// Edge of the square tiles `tile_begin` of the job table counts (the host builds the table with it).
extern "C" int w2v2_prepare_tile_edge(void) { return prep_v3() ? PV3_EDGE : 32; }

extern "C" int w2v2_prepare_weights(const w2v2_prep_job* jobs_dev, int njobs, int64_t total_tiles, void* stream) {
  W2V2_REQUIRE(njobs >= 0 && total_tiles >= 0 && total_tiles < (int64_t(1) << 31), "w2v2_prepare_weights: bad job table");
  if (njobs == 0 || total_tiles == 0) return 0;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t cap = int64_t(sms > 0 ? sms : 148) * 3;
  if (prep_v3()) {
    W2V2_REQUIRE(njobs <= PV2_MAX_JOBS, "w2v2_prepare_weights: %d jobs exceed the shared-memory table (%d)", njobs, PV2_MAX_JOBS);
    const size_t smem3 = sizeof(float) * PV3_EDGE * PV3_PITCH + 16 + sizeof(w2v2_prep_job) * size_t(njobs);
    static size_t configured3 = 0;
    if (smem3 > configured3) {
      W2V2_CHECK_CUDA(cudaFuncSetAttribute(prepare_weights_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem3)));
      configured3 = smem3;
    }
    const int64_t cap3 = int64_t(sms > 0 ? sms : 148) * 4;
    prepare_weights_v3_kernel<<<unsigned(total_tiles < cap3 ? total_tiles : cap3), 256, smem3, static_cast<cudaStream_t>(stream)>>>(
        jobs_dev, njobs, total_tiles);
    count_launches(1);
    W2V2_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  static const bool v2 = []() { const char* e = getenv("W2V2_PREP_V2"); return !(e != nullptr && e[0] == '0'); }();
  if (v2 && njobs <= PV2_MAX_JOBS) {
    const size_t smem = sizeof(w2v2_prep_job) * size_t(njobs) + sizeof(float) * PV2_GROUP * 32 * 33;
    static bool configured = false;
    if (!configured) {
      W2V2_CHECK_CUDA(cudaFuncSetAttribute(prepare_weights_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           int(sizeof(w2v2_prep_job) * PV2_MAX_JOBS + sizeof(float) * PV2_GROUP * 32 * 33)));
      configured = true;
    }
    const int64_t groups = (total_tiles + PV2_GROUP - 1) / PV2_GROUP;
    prepare_weights_v2_kernel<<<unsigned(groups < cap ? groups : cap), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        jobs_dev, njobs, total_tiles);
    count_launches(1);
    W2V2_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  prepare_weights_kernel<<<unsigned(total_tiles < cap ? total_tiles : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      jobs_dev, njobs, total_tiles);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}
