// Weight gradient of the positional convolution (grouped Conv1d, k = 128, pad = 64, groups = 16; HF:340-368)
// on tcgen05, without an im2col:
//
//     dW[g*I + o][k][i] += sum_{b,t} dz[b, t, g*I + o] * x[b, t + k - 64, g*I + i]
//
// (torch autograd reaches the same numbers through cuDNN's grouped-conv backward-filter.)
//
// Per (group, utterance, time chunk) two activation slabs sit in shared memory in the no-swizzle UMMA
// layout [8-channel plane][time row][16 B]; a slab read MN-major has the channels on M / N and the time
// rows on K, and "the same slab `r` time steps later" is the descriptor start address + 16*r bytes.
// So for tap k the MMA is   D_k[o, i] += A[o, tau] * B_k[i, tau]   with A = dz slab, B_k = x slab + 16*k.
//
// M = 128 holds TWO copies of the dz slab, the second one delayed by 8 time steps (planes 8..15):
// accumulator rows 0..I-1 collect tap k0+u, rows 64..64+I-1 tap k0+8+u -- the tensor core has no M = 48.
// A unit = (group, 16 consecutive taps) keeps its 8 accumulators ([128 x I] fp32 each) in TMEM across
// all utterances; the 148 persistent CTAs split the (unit, utterance) list evenly and flush with
// red.global.add at unit boundaries (at most one unit is shared by two CTAs).
//   warps 0-3: slab producers (cp.async 16 B pieces, zero fill outside the utterance)
//   warps 4-7: flush (TMEM -> red.global.add.v4.f32), one TMEM lane quarter each
//   warp  8  : MMA issuer
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int device_sm_count();

constexpr int PW_LOADERS = 128;
constexpr int PW_THREADS = 288;
constexpr int PW_TAPS = 8;         // taps per accumulator half; a unit covers 2 * PW_TAPS taps
constexpr int PW_MAX_STAGES = 6;

struct PosWgradParams {
  const __half* dz;   // [B, T, H]
  const __half* x;    // [B, T, H]
  float* dw;          // [H][K][I]
  int B, T, H, K;
  int TC, nchunks;    // time rows per chunk (multiple of 16), chunks per utterance (cover T + 8 rows)
  int stages;
  int nunits;         // groups * K / 16
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int I>
__global__ void __launch_bounds__(PW_THREADS, 1) posconv_wgrad_kernel(const PosWgradParams p) {
  constexpr int OB = I / 8;                       // 8-channel planes per group
  extern __shared__ __align__(128) uint8_t smem[];
  const int TC = p.TC, XR = TC + PW_TAPS;
  const int pitch_a = TC * 16, pitch_x = XR * 16;
  const int a_bytes = 16 * pitch_a;               // 16 planes: [0, OB) = dz, [8, 8 + OB) = dz delayed by 8 rows
  const int stage_bytes = a_bytes + OB * pitch_x;
  const int S = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
  uint64_t* empty_bar = full_bar + PW_MAX_STAGES;
  uint64_t* acc_full = empty_bar + PW_MAX_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's share of the (unit, utterance) list, unit-major
  const int64_t items = int64_t(p.nunits) * p.B;
  const int64_t it0 = items * blockIdx.x / gridDim.x, it1 = items * (blockIdx.x + 1) / gridDim.x;
  const int units_per_group = p.K / (2 * PW_TAPS);

  if (warp == 8 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], PW_LOADERS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp < 4) {
    // the planes no producer writes (OB..7, 8+OB..15) feed accumulator rows nobody reads, but keep them finite
    for (int i = threadIdx.x; i < S * stage_bytes / 16; i += PW_LOADERS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ slab producers
    const int tid = threadIdx.x;
    int stage = 0, pending = -1;
    uint32_t phase = 0;
    for (int64_t it = it0; it < it1; ++it) {
      const int unit = int(it / p.B), b = int(it % p.B);
      const int g = unit / units_per_group, k0 = (unit % units_per_group) * 2 * PW_TAPS;
      const __half* dzb = p.dz + int64_t(b) * p.T * p.H + g * I;
      const __half* xb = p.x + int64_t(b) * p.T * p.H + g * I;
      for (int c = 0; c < p.nchunks; ++c) {
        const int t0 = c * TC;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint32_t sx = sa + a_bytes;
        // dz, both copies: piece (copy, tau, plane) <- dz[t0 + tau - 8 * copy, plane]
        for (int q = tid; q < 2 * OB * TC; q += PW_LOADERS) {
          const int plane = q % OB, rest = q / OB;
          const int tau = rest % TC, copy = rest / TC;
          const int t = t0 + tau - PW_TAPS * copy;
          const bool ok = t >= 0 && t < p.T;
          cp_async16_zfill(sa + (copy * 8 + plane) * pitch_a + tau * 16, dzb + int64_t(ok ? t : 0) * p.H + plane * 8, ok ? 16u : 0u);
        }
        // x: row r <- x[t0 + k0 - 64 + r]
        for (int q = tid; q < OB * XR; q += PW_LOADERS) {
          const int plane = q % OB, r = q / OB;
          const int t = t0 + k0 - p.K / 2 + r;
          const bool ok = t >= 0 && t < p.T;
          cp_async16_zfill(sx + plane * pitch_x + r * 16, xb + int64_t(ok ? t : 0) * p.H + plane * 8, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (pending >= 0) {               // the previous stage's copies have landed: publish it
          cp_async_wait<1>();
          fence_proxy_async_smem();
          mbar_arrive(&full_bar[pending]);
        }
        pending = stage;
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
    if (pending >= 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(&full_bar[pending]);
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, I, 1, 1);      // both operands MN-major
      int stage = 0;
      uint32_t phase = 0, flush_phase = 0;
      int cur_unit = -1;
      bool fresh = true;
      for (int64_t it = it0; it < it1; ++it) {
        const int unit = int(it / p.B);
        if (unit != cur_unit) {
          if (cur_unit >= 0) {
            umma_commit(acc_full);                   // previous unit complete -> flush warps
            mbar_wait(acc_empty, flush_phase);       // ... and wait until they have drained TMEM
            flush_phase ^= 1;
            tc_fence_after();
          }
          cur_unit = unit;
          fresh = true;
        }
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sx = sa + a_bytes;
          for (int ks = 0; ks < TC / 16; ++ks) {
            const uint64_t adesc = make_smem_desc(sa + ks * 256, 128, pitch_a, 0);
#pragma unroll
            for (int u = 0; u < PW_TAPS; ++u) {
              const uint64_t bdesc = make_smem_desc(sx + (ks * 16 + u) * 16, 128, pitch_x, 0);
              umma_f16(tmem + u * I, adesc, bdesc, idesc, (fresh && ks == 0) ? 0u : 1u);
            }
          }
          fresh = false;
          umma_commit(&empty_bar[stage]);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      if (cur_unit >= 0) umma_commit(acc_full);
    }
  } else {
    // ------------------------------------------------------------------ flush warps (4..7): quarter = warp & 3
    const int quarter = warp & 3;
    const int half = quarter >> 1;                       // 0: taps k0 + u, 1: taps k0 + 8 + u
    const int o = (quarter & 1) * 32 + lane;             // output channel inside the group
    uint32_t full_phase = 0;
    int64_t it = it0;
    while (it < it1) {
      const int unit = int(it / p.B);
      const int64_t next = min(it1, int64_t(unit + 1) * p.B);   // first item of the next unit
      const int g = unit / units_per_group, k0 = (unit % units_per_group) * 2 * PW_TAPS;
      mbar_wait(acc_full, full_phase);
      full_phase ^= 1;
      tc_fence_after();
      float* drow = p.dw + (int64_t(g) * I + o) * (int64_t(p.K) * I) + int64_t(k0 + half * PW_TAPS) * I;
#pragma unroll 1
      for (int u = 0; u < PW_TAPS; ++u) {
#pragma unroll
        for (int c16 = 0; c16 < I / 16; ++c16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(tmem + (uint32_t(quarter * 32) << 16) + u * I + c16 * 16, r);
          tmem_ld_wait();
          if (o < I) {
            float* d = drow + u * I + c16 * 16;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              red_add_v4(d + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                         __uint_as_float(r[j + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      it = next;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_posconv_wgrad(const void* dz16, const void* x16, float* dw_hki, int B, int T, int H, int groups, int K,
                                  void* stream) {
  W2V2_REQUIRE(groups > 0 && H % groups == 0, "w2v2_posconv_wgrad: H=%d not divisible by groups=%d", H, groups);
  const int I = H / groups;
  W2V2_REQUIRE(I == 48 || I == 64, "w2v2_posconv_wgrad: %d channels per group (built for 48 and 64)", I);
  W2V2_REQUIRE(K % (2 * PW_TAPS) == 0, "w2v2_posconv_wgrad: K=%d must be a multiple of %d", K, 2 * PW_TAPS);
  W2V2_REQUIRE(B >= 1 && T >= 1, "w2v2_posconv_wgrad: empty problem");
  PosWgradParams p;
  p.dz = static_cast<const __half*>(dz16);
  p.x = static_cast<const __half*>(x16);
  p.dw = dw_hki;
  p.B = B; p.T = T; p.H = H; p.K = K;
  const int TT = T + PW_TAPS;                      // the delayed dz copy needs 8 more time rows
  p.nchunks = (TT + 127) / 128;
  p.TC = (((TT + p.nchunks - 1) / p.nchunks) + 15) / 16 * 16;
  const int OB = I / 8;
  const int stage_bytes = 16 * p.TC * 16 + OB * (p.TC + PW_TAPS) * 16;
  int stages = (220 * 1024) / stage_bytes;
  if (stages > PW_MAX_STAGES) stages = PW_MAX_STAGES;
  W2V2_REQUIRE(stages >= 2, "w2v2_posconv_wgrad: stage of %d bytes does not fit twice", stage_bytes);
  p.stages = stages;
  p.nunits = groups * (K / (2 * PW_TAPS));
  const int smem_bytes = stages * stage_bytes + (2 * PW_MAX_STAGES + 2) * 8 + 16;
  const int64_t items = int64_t(p.nunits) * B;
  const int sms = device_sm_count();
  const int grid = int(items < sms ? items : sms);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (I == 48) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(posconv_wgrad_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    posconv_wgrad_kernel<48><<<grid, PW_THREADS, smem_bytes, st>>>(p);
  } else {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(posconv_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    posconv_wgrad_kernel<64><<<grid, PW_THREADS, smem_bytes, st>>>(p);
  }
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}
