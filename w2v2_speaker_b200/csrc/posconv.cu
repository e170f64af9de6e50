// Positional convolution embedding on tcgen05 (HF:326-379):
//     y[b,t,o] = GELU( bias[o] + sum_{j<K, i<I} w[g(o)][j][o][i] * x[b, t + j - K/2, g(o)*I + i] )
// i.e. a grouped Conv1d(H->H, k=K=128, pad=64, groups=16) whose last output frame is dropped.
//
// cuDNN treats this as a grouped conv; here it is K shifted GEMMs per group accumulated in TMEM.
// One CTA owns one group g and NB batch elements.  The [rows, I] activation slab of those batch
// elements (zero padded, 64 rows between consecutive utterances so one pad serves both neighbours)
// sits in shared memory in the *no-swizzle K-major* UMMA layout with a 16-byte row pitch per
// 8-channel plane.  In that layout "the A tile shifted by j rows" is just the descriptor start
// address + 16*j bytes, so all K taps read the same slab -- no im2col, no re-staging.
// The per-tap [O x I] weight blocks (pre-folded weight norm, fp16, already in UMMA core-matrix
// order) stream through a 4-stage ring with cp.async.bulk + mbarriers.
//   warp 0: weight producer, warp 1: MMA issuer, warps 2-5: slab fill, then epilogue
//   (TMEM -> +bias -> GELU -> fp32 store).
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

constexpr int PC_TAPS_PER_STAGE = 4;
constexpr int PC_STAGES = 4;
constexpr int PC_THREADS = 192;
constexpr int PC_PAD = 64;       // = K/2 rows of zeros in front of every utterance

struct PosconvParams {
  const __half* x;      // [B, T, H]
  const __half* w;      // [G][K][I/8][O][8]
  const float* bias;    // [H]
  float* out;           // [B, T, H]
  int B, T, H, G, K, I, O;
  int NB, P, ntiles, R; // batch per CTA, row pitch per utterance (T+64), M tiles, slab rows
  int tmem_cols;
};

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__global__ void __launch_bounds__(PC_THREADS, 1) posconv_kernel(const PosconvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int I = p.I, O = p.O, R = p.R;
  const int planes = I / 8;
  const int plane_bytes = R * 16;
  const int tap_bytes = I * O * 2;
  const int stage_bytes = PC_TAPS_PER_STAGE * tap_bytes;
  uint8_t* slab = smem;
  uint8_t* wring = slab + planes * plane_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(wring + PC_STAGES * stage_bytes);
  uint64_t* empty_bar = full_bar + PC_STAGES;
  uint64_t* acc_bar = empty_bar + PC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x;
  const int b0 = blockIdx.y * p.NB;
  const int nb = min(p.NB, p.B - b0);
  const int nstages_total = p.K / PC_TAPS_PER_STAGE;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  // ---- fill the slab (all threads): rows b*P + 64 + t hold x[b0+b, t, g*I : (g+1)*I], everything else 0
  {
    const int total = R * planes;
    for (int idx = threadIdx.x; idx < total; idx += PC_THREADS) {
      const int s = idx % R, c = idx / R;
      const int bl = s / p.P, r = s % p.P;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (bl < nb && r >= PC_PAD) {
        const __half* src = p.x + (int64_t(b0 + bl) * p.T + (r - PC_PAD)) * p.H + g * I + c * 8;
        v = *reinterpret_cast<const uint4*>(src);
      }
      *reinterpret_cast<uint4*>(slab + c * plane_bytes + s * 16) = v;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint8_t* wg = reinterpret_cast<const uint8_t*>(p.w) + int64_t(g) * p.K * tap_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nstages_total; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        bulk_copy_g2s(wring + stage * stage_bytes, wg + int64_t(it) * stage_bytes, stage_bytes, &full_bar[stage]);
        if (++stage == PC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, O);
      const uint32_t slab_a = smem_u32(slab);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nstages_total; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t wbase = smem_u32(wring + stage * stage_bytes);
        for (int tp = 0; tp < PC_TAPS_PER_STAGE; ++tp) {
          const int j = it * PC_TAPS_PER_STAGE + tp;
          for (int ks = 0; ks < I / 16; ++ks) {
            // B: [O x 16] block of tap j: planes of O rows x 16 B; LBO = plane pitch, SBO = 8 rows
            const uint64_t bdesc = make_smem_desc(wbase + tp * tap_bytes + (2 * ks) * (O * 16), O * 16, 128, 0);
            for (int mt = 0; mt < p.ntiles; ++mt) {
              // A: rows (mt*128 + j) .. +127 of the slab, k-chunks 2ks, 2ks+1
              const uint64_t adesc =
                  make_smem_desc(slab_a + (2 * ks) * plane_bytes + (mt * 128 + j) * 16, plane_bytes, 128, 0);
              umma_f16(tmem + mt * O, adesc, bdesc, idesc, (j | ks) != 0);
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == PC_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_bar);
    }
  } else {
    // ---- epilogue
    const int quarter = warp & 3;
    mbar_wait(acc_bar, 0);
    __syncwarp();
    tc_fence_after();
    for (int mt = 0; mt < p.ntiles; ++mt) {
      const int u = mt * 128 + quarter * 32 + lane;
      const int bl = u / p.P, t = u % p.P;
      const bool valid = (bl < nb) && (t < p.T);
      float* dst = p.out + (int64_t(b0 + (valid ? bl : 0)) * p.T + (valid ? t : 0)) * p.H + g * O;
      for (int c = 0; c < O / 16; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem + (uint32_t(quarter * 32) << 16) + mt * O + c * 16, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 v;
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + g * O + c * 16 + q * 4));
            v.x = gelu_erf(__uint_as_float(r[4 * q + 0]) + bb.x);
            v.y = gelu_erf(__uint_as_float(r[4 * q + 1]) + bb.y);
            v.z = gelu_erf(__uint_as_float(r[4 * q + 2]) + bb.z);
            v.w = gelu_erf(__uint_as_float(r[4 * q + 3]) + bb.w);
            *reinterpret_cast<float4*>(dst + c * 16 + q * 4) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_posconv(const void* x16, const void* w16, const float* bias, float* out, int B, int T, int H,
                            int groups, int K, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  W2V2_REQUIRE(groups > 0 && H % groups == 0, "w2v2_posconv: H=%d not divisible by groups=%d", H, groups);
  const int I = H / groups, O = H / groups;
  W2V2_REQUIRE(I % 16 == 0 && O % 16 == 0 && O <= 256, "w2v2_posconv: channels per group (%d) must be a multiple of 16", I);
  W2V2_REQUIRE(K == 2 * PC_PAD && K % PC_TAPS_PER_STAGE == 0, "w2v2_posconv: kernel size %d unsupported (expects 128)", K);
  W2V2_REQUIRE(B >= 1 && T >= 1, "w2v2_posconv: empty input");
  PosconvParams p;
  p.x = static_cast<const __half*>(x16);
  p.w = static_cast<const __half*>(w16);
  p.bias = bias;
  p.out = out;
  p.B = B; p.T = T; p.H = H; p.G = groups; p.K = K; p.I = I; p.O = O;
  p.P = T + PC_PAD;
  const int ring = PC_STAGES * PC_TAPS_PER_STAGE * I * O * 2;
  int NB = 0, ntiles = 0, R = 0, smem = 0;
  for (int nb = (B < 4 ? B : 4); nb >= 1; --nb) {
    const int U = (nb - 1) * p.P + T;
    const int nt = (U + 127) / 128;
    const int rows = ((nt * 128 + K) + 7) / 8 * 8;
    const int bytes = rows * I * 2 + ring + 256 + 128;
    if (nt * O <= 512 && bytes <= 227 * 1024) { NB = nb; ntiles = nt; R = rows; smem = bytes; break; }
  }
  W2V2_REQUIRE(NB >= 1, "w2v2_posconv: T=%d frames does not fit the single-slab kernel (needs time tiling)", T);
  p.NB = NB; p.ntiles = ntiles; p.R = R;
  int cols = 32;
  while (cols < ntiles * O) cols *= 2;
  p.tmem_cols = cols;
  static int configured = 0;
  if (smem > configured) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(posconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid(groups, (B + NB - 1) / NB);
  posconv_kernel<<<grid, PC_THREADS, smem, stream>>>(p);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}
