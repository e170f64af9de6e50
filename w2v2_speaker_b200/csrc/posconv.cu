// Positional convolution embedding on tcgen05 (HF:326-379):
//     y[b,t,o] = GELU( bias[o] + sum_{j<K, i<I} w[g(o)][j][o][i] * x[b, t + j - K/2, g(o)*I + i] )
// i.e. a grouped Conv1d(H->H, k=K=128, pad=64, groups=16) whose last output frame is dropped.
//
// cuDNN treats this as a grouped conv; here it is K/U shifted GEMMs per group accumulated in TMEM.
// One CTA owns one (group, utterance).  The utterance's [rows, I] activation slab (64 zero rows in
// front, zeros behind) sits in shared memory in the *no-swizzle K-major* UMMA layout with a 16-byte
// row pitch per 8-channel plane.  In that layout "the A tile shifted by r rows" is just the descriptor
// start address + 16*r bytes, so every tap reads the same slab -- no im2col, no re-staging.
//
// N = 48 (one group) would leave the tensor core waiting on shared memory (the 4 KB A tile is re-read
// for only 48 columns of work), so U taps share each A tile:  with j = U*j' + u
//     F_u[r] = sum_{j'} slab[r + U*j'] . W_{U*j'+u}          (one MMA per j' with N = U*48: B = [W_u]_u)
//     y[t]   = sum_u F_u[t + u]                               (row-shifted combine in the epilogue)
// which cuts the MMA count by U at full tensor rate.  The per-(group, j') weight blocks (weight norm
// folded, fp16, UMMA core-matrix order [I/8][U][O][8]) stream through a cp.async.bulk + mbarrier ring.
//   warp 0: weight producer, warp 1: MMA issuer, warps 2-5: slab fill, then epilogue
//   (TMEM -> smem (row shift) -> +bias -> GELU -> fp32 store).
#include <stdlib.h>

#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

constexpr int PC_STAGES = 4;
constexpr int PC_THREADS = 192;
constexpr int PC_PAD = 64;       // = K/2 rows of zeros in front of the utterance
constexpr int PC_CHUNK = 16;     // output columns combined per epilogue pass
constexpr int PC_ROWF = 20;      // padded floats per row in the combine buffer (16 + 4: conflict-free)

struct PosconvParams {
  const __half* x;      // [B, T, H]
  const __half* w;      // [G][K/U][I/8][U][O][8]
  const float* bias;    // [H]
  float* out;           // [B, T, H]
  int B, T, H, G, K, I, O, U;
  int ntiles, R;        // M tiles (128 rows), slab rows
  int act, in_shift;    // apply GELU in the epilogue; slab row s holds input frame s - 64 + in_shift
  int tmem_cols;
};

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__global__ void __launch_bounds__(PC_THREADS, 2) posconv_kernel(const PosconvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int I = p.I, O = p.O, R = p.R, U = p.U;
  const int planes = I / 8;
  const int plane_bytes = R * 16;
  const int NW = U * O;                                  // MMA N
  const int stage_bytes = NW * I * 2;                    // one tap group j': [I/8][U*O][8] halfs
  const int nrows = p.ntiles * 128;
  const int ring_bytes = PC_STAGES * stage_bytes;
  const int cbuf_bytes = U * nrows * PC_ROWF * 4;
  uint8_t* slab = smem;
  uint8_t* wring = slab + planes * plane_bytes;          // reused as the epilogue combine buffer
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(wring + (ring_bytes > cbuf_bytes ? ring_bytes : cbuf_bytes));
  uint64_t* empty_bar = full_bar + PC_STAGES;
  uint64_t* acc_bar = empty_bar + PC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x;
  const int b = blockIdx.y;
  const int ngroups = p.K / U;

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
  // ---- fill the slab (all threads): rows 64 + t hold x[b, t, g*I : (g+1)*I], everything else 0
  {
    const int total = R * planes;
    for (int idx = threadIdx.x; idx < total; idx += PC_THREADS) {
      const int s = idx % R, c = idx / R;
      uint4 v = make_uint4(0, 0, 0, 0);
      const int t = s - PC_PAD + p.in_shift;
      if (t >= 0 && t < p.T)
        v = *reinterpret_cast<const uint4*>(p.x + (int64_t(b) * p.T + t) * p.H + g * I + c * 8);
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
      const uint8_t* wg = reinterpret_cast<const uint8_t*>(p.w) + int64_t(g) * ngroups * stage_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < ngroups; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        bulk_copy_g2s(wring + stage * stage_bytes, wg + int64_t(it) * stage_bytes, stage_bytes, &full_bar[stage]);
        if (++stage == PC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, NW);
      const uint32_t slab_a = smem_u32(slab);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < ngroups; ++it) {       // it = j' : taps U*j' .. U*j'+U-1
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t wbase = smem_u32(wring + stage * stage_bytes);
        for (int ks = 0; ks < I / 16; ++ks) {
          // B: [U*O x 16] block: planes of U*O rows x 16 B; LBO = plane pitch, SBO = 8 rows
          const uint64_t bdesc = make_smem_desc(wbase + (2 * ks) * (NW * 16), NW * 16, 128, 0);
          for (int mt = 0; mt < p.ntiles; ++mt) {
            // A: rows (mt*128 + U*j') .. +127 of the slab, k-chunks 2ks, 2ks+1
            const uint64_t adesc =
                make_smem_desc(slab_a + (2 * ks) * plane_bytes + (mt * 128 + U * it) * 16, plane_bytes, 128, 0);
            umma_f16(tmem + mt * NW, adesc, bdesc, idesc, (it | ks) != 0);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == PC_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_bar);
    }
  } else {
    // ---- epilogue: y[t] = sum_u F_u[t + u]  (+bias, GELU).  Thread <-> accumulator row r; the row
    // shift goes through shared memory (the weight ring is idle by now), PC_CHUNK columns per pass.
    const int quarter = warp & 3;
    const int et = (warp - 2) * 32 + lane;              // 0..127
    float* cbuf = reinterpret_cast<float*>(wring);      // [U][nrows][PC_ROWF] floats
    mbar_wait(acc_bar, 0);
    __syncwarp();
    tc_fence_after();
    for (int c0 = 0; c0 < O; c0 += PC_CHUNK) {
      // 1. spill F_u[r][c0 : c0+16] of every tile to smem
      for (int mt = 0; mt < p.ntiles; ++mt) {
        const int r = mt * 128 + quarter * 32 + lane;
        for (int u = 0; u < U; ++u) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tmem + (uint32_t(quarter * 32) << 16) + mt * NW + u * O + c0, v);
          tmem_ld_wait();
          float* dst = cbuf + (int64_t(u) * nrows + r) * PC_ROWF;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(dst + 4 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
      named_bar_sync(1, 128);
      // 2. combine rows t+u, add bias, GELU, store
      for (int t = et; t < p.T; t += 128) {
        float acc[PC_CHUNK];
#pragma unroll
        for (int q = 0; q < PC_CHUNK; q += 4) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias != nullptr) bb = __ldg(reinterpret_cast<const float4*>(p.bias + g * O + c0 + q));
          acc[q] = bb.x; acc[q + 1] = bb.y; acc[q + 2] = bb.z; acc[q + 3] = bb.w;
        }
        for (int u = 0; u < U; ++u) {
          const float* src = cbuf + (int64_t(u) * nrows + t + u) * PC_ROWF;      // t + u < nrows by construction
#pragma unroll
          for (int q = 0; q < PC_CHUNK; q += 4) {
            const float4 f = *reinterpret_cast<const float4*>(src + q);
            acc[q] += f.x; acc[q + 1] += f.y; acc[q + 2] += f.z; acc[q + 3] += f.w;
          }
        }
#pragma unroll
        for (int q = 0; q < PC_CHUNK; q += 2)
          if (p.act) gelu_erf2(acc[q], acc[q + 1]);
        float* dst = p.out + (int64_t(b) * p.T + t) * p.H + g * O + c0;
#pragma unroll
        for (int q = 0; q < PC_CHUNK; q += 4)
          *reinterpret_cast<float4*>(dst + q) = make_float4(acc[q], acc[q + 1], acc[q + 2], acc[q + 3]);
      }
      named_bar_sync(1, 128);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

static int posconv_plan(int T, int H, int groups, int* ntiles_out) {
  const int O = H / groups;
  static int umax = -1;               // W2V2_POSCONV_U caps the taps per MMA (tuning knob)
  if (umax < 0) {
    const char* e = getenv("W2V2_POSCONV_U");
    umax = (e != nullptr && atoi(e) >= 1) ? atoi(e) : 4;
  }
  for (int U = 4; U >= 1; U >>= 1) {
    if (U > umax) continue;
    const int nt = (T + U - 1 + 127) / 128;       // rows t + u (u < U) must all exist
    if (nt * U * O <= 512 && U * O <= 256) {
      if (ntiles_out) *ntiles_out = nt;
      return U;
    }
  }
  return 0;
}

}  // namespace w2v2

using namespace w2v2;

// taps sharing one A tile for a given number of frames (selects the folded weight layout)
extern "C" int w2v2_posconv_taps_per_mma(int T, int H, int groups) {
  if (groups <= 0 || H % groups != 0 || T < 1) return 0;
  return posconv_plan(T, H, groups, nullptr);
}

extern "C" int w2v2_posconv_ex(const void* x16, const void* w16, const float* bias, float* out, int B, int T, int H,
                               int groups, int K, int act, int in_shift, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  W2V2_REQUIRE(groups > 0 && H % groups == 0, "w2v2_posconv: H=%d not divisible by groups=%d", H, groups);
  const int I = H / groups, O = H / groups;
  W2V2_REQUIRE(I % 16 == 0 && O % 16 == 0 && O <= 256, "w2v2_posconv: channels per group (%d) must be a multiple of 16", I);
  W2V2_REQUIRE(K == 2 * PC_PAD, "w2v2_posconv: kernel size %d unsupported (expects 128)", K);
  W2V2_REQUIRE(B >= 1 && B <= 65535 && T >= 1, "w2v2_posconv: bad batch / length (B=%d T=%d)", B, T);
  PosconvParams p;
  const int U = posconv_plan(T, H, groups, &p.ntiles);
  W2V2_REQUIRE(U >= 1, "w2v2_posconv: T=%d frames does not fit the single-slab kernel (needs time tiling)", T);
  p.x = static_cast<const __half*>(x16);
  p.w = static_cast<const __half*>(w16);
  p.bias = bias;
  p.out = out;
  p.B = B; p.T = T; p.H = H; p.G = groups; p.K = K; p.I = I; p.O = O; p.U = U;
  p.act = act; p.in_shift = in_shift;
  p.R = (p.ntiles * 128 + K + 7) / 8 * 8;
  const int ring = PC_STAGES * U * O * I * 2;
  const int cbuf = U * p.ntiles * 128 * PC_ROWF * 4;
  const int smem = p.R * I * 2 + (ring > cbuf ? ring : cbuf) + 256;
  W2V2_REQUIRE(smem <= 227 * 1024, "w2v2_posconv: T=%d needs %d bytes of shared memory (time tiling not implemented)", T, smem);
  int cols = 32;
  while (cols < p.ntiles * U * O) cols *= 2;
  p.tmem_cols = cols;
  static int configured = 0;
  if (smem > configured) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(posconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid(groups, B);
  posconv_kernel<<<grid, PC_THREADS, smem, stream>>>(p);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_posconv(const void* x16, const void* w16, const float* bias, float* out, int B, int T, int H,
                            int groups, int K, void* stream) {
  return w2v2_posconv_ex(x16, w16, bias, out, B, T, H, groups, K, 1, 0, stream);
}
