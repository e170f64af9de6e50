// Self-attention backward for T <= 160 frames (3 s utterances: T = 149) -- the short-chain variant of
// attention_bwd.cu (same math, backward of HF:438-463):
//     P = exp(S - lse), S = Q K^T;  dP = dO V^T;  dS = P * (dP - delta);  dV = P^T dO;  dQ = dS K;  dK = dS^T Q
//
// One CTA per (batch, head).  The sequence is latency-bound (a few dependent MMA -> arithmetic -> MMA
// round trips on ~20 K elements), so the kernel is organised to make the dependent chain as short as
// possible rather than to save instructions:
//   * S and dP of a query tile are issued TOGETHER into separate TMEM regions, and ONE arithmetic pass
//     reads both and writes P (dropped, if attention dropout is on) and dS -- one round trip and one
//     dropout hash per element pair instead of two;
//   * the second query tile (rows 128 .. T-1, at most 32) gets its own small P / dS buffers, so the MMAs
//     that contract over queries (dV, dK) run ONCE at the end over all rows: their 256 accumulator
//     columns reuse the S / dP regions, which is what makes room for separate S and dP in the first place;
//   * dQ of tile 0 is computed in the same MMA group as S / dP of tile 1 and drained while the tile-1
//     arithmetic runs; the O tiles ride in the same TMA transaction as Q / K / V / dO (they land in the
//     dS buffer, which is not written before delta = rowsum(dO * O) has been formed from shared memory).
// Three MMA round trips and two arithmetic passes per (batch, head) instead of six and four.
// 16 warps: warp & 3 = TMEM lane quarter (32 query rows), warp >> 2 = one of four column groups.
// TMEM (fp32 columns): S [0,160) | dP [160,320) | dQ tile 0 [320,384) | dQ tile 1 [384,448); at the end
// dV [0,128) | dK [128,256).
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 int swizzle_bytes);

constexpr int AF_D = 64;
constexpr int AF_THREADS = 512;
constexpr int AF_MAX_TK = 160;
constexpr int AF_COL_S = 0, AF_COL_DP = 160, AF_COL_DQ0 = 320, AF_COL_DQ1 = 384, AF_COL_DV = 0, AF_COL_DK = 128;
// shared memory map (bytes); P / dS as [rows x 64-key blocks] of 128-byte swizzled rows
constexpr int AF_BLK0 = 16384;                   // tile 0: 128 rows per 64-key block
constexpr int AF_BLK1 = 4096;                    // tile 1: 32 rows per 64-key block
constexpr int AF_P0 = 0;
constexpr int AF_DS0 = AF_P0 + 3 * AF_BLK0;
constexpr int AF_P1 = AF_DS0 + 3 * AF_BLK0;
constexpr int AF_DS1 = AF_P1 + 3 * AF_BLK1;
constexpr int AF_Q0 = AF_DS1 + 3 * AF_BLK1;
constexpr int AF_Q1 = AF_Q0 + 16384;
constexpr int AF_DO0 = AF_Q1 + 4096;
constexpr int AF_DO1 = AF_DO0 + 16384;
constexpr int AF_K = AF_DO1 + 4096;
constexpr int AF_V = AF_K + AF_MAX_TK * 128;
constexpr int AF_RED = AF_V + AF_MAX_TK * 128;   // float [2 tiles][4 groups][128 rows] partial deltas
constexpr int AF_BARS = AF_RED + 2 * 4 * 128 * 4;
constexpr int AF_SMEM = AF_BARS + 64;
// 128-row operand windows over the 32-row tile-1 buffers read (never use) bytes behind them: every such
// window still ends inside the allocation because the K / V tiles come last.
static_assert(AF_DO1 + 16384 <= AF_SMEM && AF_DS1 + 2 * AF_BLK1 + 16384 <= AF_SMEM, "operand windows leave the allocation");
static_assert(AF_SMEM <= 227 * 1024, "shared memory budget");

struct alignas(64) AttnBwdFusedParams {
  CUtensorMap tmQ0;    // qkv: box {64, 128, 1}
  CUtensorMap tmQ1;    // qkv: box {64, 32, 1}
  CUtensorMap tmKV;    // qkv: box {64, TK, 1}
  CUtensorMap tmDO0;   // dO : box {64, 128, 1}
  CUtensorMap tmDO1;   // dO : box {64, 32, 1}
  CUtensorMap tmO0;    // O  : box {64, 128, 1}
  CUtensorMap tmO1;    // O  : box {64, 32, 1}
  const float* lse;    // [B, heads, T]
  __half* dqkv;        // [B*T, 3H]
  int T, TK, H, heads, qtiles;
  float qscale;        // dq is multiplied by this (chain rule of the d^-0.5 folded into Wq); 1 = leave as is
  float* dbias;        // f32 [3H] (+)= column sums of (dq * qscale | dk | dv), or nullptr
  uint32_t drop_thr;
  float drop_inv_keep;
  unsigned long long drop_seed;
};

__device__ __forceinline__ void store_row_f16x16(__half* dst, const uint32_t (&r)[16]) {
  uint4 a, b;
  a.x = pack_half2(__uint_as_float(r[0]), __uint_as_float(r[1]));
  a.y = pack_half2(__uint_as_float(r[2]), __uint_as_float(r[3]));
  a.z = pack_half2(__uint_as_float(r[4]), __uint_as_float(r[5]));
  a.w = pack_half2(__uint_as_float(r[6]), __uint_as_float(r[7]));
  b.x = pack_half2(__uint_as_float(r[8]), __uint_as_float(r[9]));
  b.y = pack_half2(__uint_as_float(r[10]), __uint_as_float(r[11]));
  b.z = pack_half2(__uint_as_float(r[12]), __uint_as_float(r[13]));
  b.w = pack_half2(__uint_as_float(r[14]), __uint_as_float(r[15]));
  reinterpret_cast<uint4*>(dst)[0] = a;
  reinterpret_cast<uint4*>(dst)[1] = b;
}

__global__ void __launch_bounds__(AF_THREADS, 1) attention_bwd_fused_kernel(const __grid_constant__ AttnBwdFusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  float* red = reinterpret_cast<float*>(smem + AF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AF_BARS);
  uint64_t* bar_tma = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int TK = p.TK;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, cg = warp >> 2;
  const int row = quarter * 32 + lane;
  pdl_trigger();

  if (threadIdx.x == 0) {
    prefetch_tensormap(&p.tmQ0);
    prefetch_tensormap(&p.tmKV);
    prefetch_tensormap(&p.tmDO0);
    prefetch_tensormap(&p.tmO0);
    if (p.qtiles > 1) {
      prefetch_tensormap(&p.tmQ1);
      prefetch_tensormap(&p.tmDO1);
      prefetch_tensormap(&p.tmO1);
    }
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + (uint32_t(quarter * 32) << 16);
  pdl_wait();

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_tma, 3 * 16384 + (p.qtiles > 1 ? 3 * 4096 : 0) + 2 * TK * 128);
    tma_load_3d(smem + AF_Q0, &p.tmQ0, bar_tma, h * AF_D, 0, b);
    tma_load_3d(smem + AF_DO0, &p.tmDO0, bar_tma, h * AF_D, 0, b);
    tma_load_3d(smem + AF_K, &p.tmKV, bar_tma, p.H + h * AF_D, 0, b);
    tma_load_3d(smem + AF_V, &p.tmKV, bar_tma, 2 * p.H + h * AF_D, 0, b);
    tma_load_3d(smem + AF_DS0, &p.tmO0, bar_tma, h * AF_D, 0, b);                  // O tile 0 (borrowed buffer)
    if (p.qtiles > 1) {
      tma_load_3d(smem + AF_Q1, &p.tmQ1, bar_tma, h * AF_D, 128, b);
      tma_load_3d(smem + AF_DO1, &p.tmDO1, bar_tma, h * AF_D, 128, b);
      tma_load_3d(smem + AF_DS0 + 16384, &p.tmO1, bar_tma, h * AF_D, 128, b);     // O tile 1
    }
  }
  // lse of both tiles: fetched now, first used in the arithmetic passes
  float lse_t[2] = {0.f, 0.f};
#pragma unroll
  for (int qt = 0; qt < 2; ++qt) {
    const int t = qt * 128 + row;
    if (qt < p.qtiles && t < p.T) lse_t[qt] = __ldg(p.lse + (int64_t(b) * p.heads + h) * p.T + t);
  }

  float delta_t[2] = {0.f, 0.f};
  mbar_wait(bar_tma, 0);
  __syncwarp();
  tc_fence_after();

  const int nchunk = TK / 16;
  const int c_begin = (nchunk * cg) >> 2, c_end = (nchunk * (cg + 1)) >> 2;
  const int ktiles = (TK + 127) / 128;
  uint32_t mma_phase = 0;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t aK = sbase + AF_K, aV = sbase + AF_V;
  const DropKeys dkeys = drop_keys(p.drop_seed);
  const uint32_t drop_thr = p.drop_thr;
  const float inv_keep = p.drop_inv_keep;
  const uint32_t idesc_s = make_idesc_f16(128, TK);
  const uint32_t idesc_dq = make_idesc_f16(128, AF_D, 0, 1);        // B (= K) MN-major
  const uint32_t idesc_t = make_idesc_f16(128, AF_D, 1, 1);         // A (P^T / dS^T) and B (dO / Q) MN-major

  // S = Q_qt K^T and dP = dO_qt V^T into their own TMEM regions
  auto issue_s_dp = [&](int qt) {
    const uint32_t aQ = sbase + (qt == 0 ? AF_Q0 : AF_Q1), adO = sbase + (qt == 0 ? AF_DO0 : AF_DO1);
#pragma unroll
    for (int k = 0; k < AF_D / 16; ++k)
      umma_f16(tmem + AF_COL_S, make_desc_k_sw128(aQ + k * 32), make_desc_k_sw128(aK + k * 32), idesc_s, k != 0);
#pragma unroll
    for (int k = 0; k < AF_D / 16; ++k)
      umma_f16(tmem + AF_COL_DP, make_desc_k_sw128(adO + k * 32), make_desc_k_sw128(aV + k * 32), idesc_s, k != 0);
  };
  // dQ_qt = dS_qt K
  auto issue_dq = [&](int qt) {
    const uint32_t adS = sbase + (qt == 0 ? AF_DS0 : AF_DS1);
    const uint32_t blk = qt == 0 ? AF_BLK0 : AF_BLK1;
    for (int kk = 0; kk < nchunk; ++kk) {
      const uint64_t adesc = make_desc_k_sw128(adS + (kk >> 2) * blk + (kk & 3) * 32);
      const uint64_t bdesc = make_smem_desc(aK + kk * 2048, 16, 1024, 2);
      umma_f16(tmem + (qt == 0 ? AF_COL_DQ0 : AF_COL_DQ1), adesc, bdesc, idesc_dq, kk != 0);
    }
  };
  // dV += P_qt^T dO_qt ; dK += dS_qt^T Q_qt over the 16-query steps that hold a valid row
  auto issue_dv_dk = [&](int qt, bool first) {
    const uint32_t aP = sbase + (qt == 0 ? AF_P0 : AF_P1), adS = sbase + (qt == 0 ? AF_DS0 : AF_DS1);
    const uint32_t aQ = sbase + (qt == 0 ? AF_Q0 : AF_Q1), adO = sbase + (qt == 0 ? AF_DO0 : AF_DO1);
    const uint32_t blk = qt == 0 ? AF_BLK0 : AF_BLK1;
    const int qsteps = min(8, (p.T - qt * 128 + 15) >> 4);
    for (int kt = 0; kt < ktiles; ++kt) {
      for (int ks = 0; ks < qsteps; ++ks) {
        const uint64_t a1 = make_smem_desc(aP + kt * 2 * blk + ks * 2048, blk, 1024, 2);
        const uint64_t b1 = make_smem_desc(adO + ks * 2048, 16, 1024, 2);
        umma_f16(tmem + AF_COL_DV + kt * AF_D, a1, b1, idesc_t, !(first && ks == 0));
      }
      for (int ks = 0; ks < qsteps; ++ks) {
        const uint64_t a2 = make_smem_desc(adS + kt * 2 * blk + ks * 2048, blk, 1024, 2);
        const uint64_t b2 = make_smem_desc(aQ + ks * 2048, 16, 1024, 2);
        umma_f16(tmem + AF_COL_DK + kt * AF_D, a2, b2, idesc_t, !(first && ks == 0));
      }
    }
  };
  // the single arithmetic pass of a query tile: S, dP (TMEM) -> P (or dropped P), dS (smem)
  auto pass = [&](int qt) {
    const int t = qt * 128 + row;
    const bool valid = t < p.T;
    uint8_t* prow = smem + (qt == 0 ? AF_P0 : AF_P1) + row * 128;
    uint8_t* dsrow = smem + (qt == 0 ? AF_DS0 : AF_DS1) + row * 128;
    const int blk = qt == 0 ? AF_BLK0 : AF_BLK1;
    const float lse2 = lse_t[qt] * 1.4426950408889634f;
    const float delta = delta_t[qt];
    const uint32_t pair_row = ((uint32_t(b) * p.heads + h) * p.T + (valid ? t : 0)) * uint32_t(TK / 2);
    for (int c = c_begin; c < c_end; ++c) {
      uint32_t rs[16], rd[16];
      tmem_ld_32x32b_x16(t_row + AF_COL_S + c * 16, rs);
      tmem_ld_32x32b_x16(t_row + AF_COL_DP + c * 16, rd);
      tmem_ld_wait();
      uint32_t pk[8], dk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float e0 = fast_ex2(fmaf(__uint_as_float(rs[2 * j]), 1.4426950408889634f, -lse2));
        float e1 = fast_ex2(fmaf(__uint_as_float(rs[2 * j + 1]), 1.4426950408889634f, -lse2));
        const bool v0 = valid && (c * 16 + 2 * j < p.T), v1 = valid && (c * 16 + 2 * j + 1 < p.T);
        float d0 = __uint_as_float(rd[2 * j]), d1 = __uint_as_float(rd[2 * j + 1]);
        // the tensor core multiplies the fp16-rounded probabilities: use the same values for dS
        const __half2 eh = __floats2half2_rn(v0 ? e0 : 0.f, v1 ? e1 : 0.f);
        const float2 ef = __half22float2(eh);
        uint32_t pw = *reinterpret_cast<const uint32_t*>(&eh);
        if (drop_thr != 0) {
          // dP arrives for the dropped probabilities: the same keep / rescale applies to it
          const uint32_t hb = dropout_hash32(dkeys, pair_row + uint32_t(c * 8 + j));
          const bool k0 = (hb & 0xffffu) >= drop_thr, k1 = (hb >> 16) >= drop_thr;
          pw = pack_half2(k0 ? ef.x * inv_keep : 0.f, k1 ? ef.y * inv_keep : 0.f);
          d0 = k0 ? d0 * inv_keep : 0.f;
          d1 = k1 ? d1 * inv_keep : 0.f;
        }
        pk[j] = pw;
        // select (not multiply) on invalid rows / columns so that whatever dP holds there cannot leak
        dk[j] = pack_half2(v0 ? ef.x * (d0 - delta) : 0.f, v1 ? ef.y * (d1 - delta) : 0.f);
      }
      const int col = c * 16;
      const int c16 = (col & 63) >> 3;
      const int o0 = (col >> 6) * blk + ((c16 ^ (row & 7)) << 4), o1 = (col >> 6) * blk + (((c16 + 1) ^ (row & 7)) << 4);
      *reinterpret_cast<uint4*>(prow + o0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(prow + o1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      *reinterpret_cast<uint4*>(dsrow + o0) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
      *reinterpret_cast<uint4*>(dsrow + o1) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
    }
  };
  // 16 of the 64 dQ columns of this thread's row -> global (x qscale), and their share of the q-bias gradient
  auto drain_dq = [&](int qt) {
    const int t = qt * 128 + row;
    uint32_t r[16];
    tmem_ld_32x32b_x16(t_row + (qt == 0 ? AF_COL_DQ0 : AF_COL_DQ1) + cg * 16, r);
    tmem_ld_wait();
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[j] = __uint_as_float(r[j]) * p.qscale;
      r[j] = __float_as_uint(v[j]);
    }
    if (t < p.T) store_row_f16x16(p.dqkv + (int64_t(b) * p.T + t) * 3 * p.H + h * AF_D + cg * 16, r);
    if (p.dbias != nullptr) warp_colsum_atomic<16>(v, t < p.T, 1.0f, p.dbias + h * AF_D + cg * 16);
  };
  auto commit_and_wait = [&]() {
    if (threadIdx.x == 0) umma_commit(bar_mma);
    __syncwarp();
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    __syncwarp();
    tc_fence_after();
  };
  auto publish_smem = [&]() {       // generic-proxy writes of P / dS -> visible to the MMAs issued after the barrier
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  };

  // ---- round 1: S0, dP0 -- issued first, delta is formed while they run
  if (threadIdx.x == 0) {
    issue_s_dp(0);
    umma_commit(bar_mma);
  }
  __syncwarp();
  // ---- delta = rowsum(dO * O) of both tiles from shared memory: this thread's quarter (16 of the 64 head
  //      dims = two 16-byte chunks of the swizzled row), combined across the four column groups
#pragma unroll
  for (int qt = 0; qt < 2; ++qt) {
    float part = 0.f;
    if (qt < p.qtiles && (qt == 0 || quarter == 0)) {      // tile 1 holds 32 rows: lane quarter 0
      const uint8_t* orow = smem + AF_DS0 + qt * 16384 + row * 128;
      const uint8_t* grow = smem + (qt == 0 ? AF_DO0 : AF_DO1) + row * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int off = ((2 * cg + c) ^ (row & 7)) << 4;
        const uint4 a = *reinterpret_cast<const uint4*>(orow + off), g = *reinterpret_cast<const uint4*>(grow + off);
        const __half2* ah = reinterpret_cast<const __half2*>(&a);
        const __half2* gh = reinterpret_cast<const __half2*>(&g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = __half22float2(ah[j]), y = __half22float2(gh[j]);
          part = fmaf(x.x, y.x, part);
          part = fmaf(x.y, y.y, part);
        }
      }
    }
    red[(qt * 4 + cg) * 128 + row] = part;
  }
  __syncthreads();           // also: every read of the borrowed O tiles precedes the first dS write
#pragma unroll
  for (int qt = 0; qt < 2; ++qt)
    delta_t[qt] = (red[(qt * 4 + 0) * 128 + row] + red[(qt * 4 + 1) * 128 + row]) +
                  (red[(qt * 4 + 2) * 128 + row] + red[(qt * 4 + 3) * 128 + row]);

  mbar_wait(bar_mma, mma_phase);
  mma_phase ^= 1;
  __syncwarp();
  tc_fence_after();
  if (quarter * 32 < p.T) pass(0);                      // warp-uniform
  publish_smem();
  if (p.qtiles > 1) {
    // ---- round 2: dQ0 ; S1, dP1 (the S / dP regions were fully consumed by pass 0)
    if (threadIdx.x == 0) {
      issue_dq(0);
      issue_s_dp(1);
    }
    commit_and_wait();
    if (quarter == 0) pass(1);                          // rows 128 .. 159 live in lane quarter 0
    if (quarter * 32 < p.T) drain_dq(0);
    publish_smem();
    // ---- round 3: dQ1 ; dV, dK over all query rows (they overwrite the S / dP regions)
    if (threadIdx.x == 0) {
      issue_dq(1);
      issue_dv_dk(0, true);
      issue_dv_dk(1, false);
    }
    commit_and_wait();
    if (quarter == 0) drain_dq(1);
  } else {
    if (threadIdx.x == 0) {
      issue_dq(0);
      issue_dv_dk(0, true);
    }
    commit_and_wait();
    if (quarter * 32 < p.T) drain_dq(0);
  }

  // ---- dK / dV rows (keys) -> global: column group -> (dK | dV, 32-column half)
  for (int kt = 0; kt < ktiles; ++kt) {
    if (kt * 128 + quarter * 32 >= p.T) continue;       // warp-uniform
    const int key = kt * 128 + row;
    const int which = cg >> 1, half = cg & 1;           // 0: dK, 1: dV
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_row + (which == 0 ? AF_COL_DK : AF_COL_DV) + kt * AF_D + half * 32, r);
    tmem_ld_wait();
    if (key < p.T) {
      __half* dst = p.dqkv + (int64_t(b) * p.T + key) * 3 * p.H + h * AF_D + (which == 0 ? p.H : 2 * p.H) + half * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 q;
        q.x = pack_half2(__uint_as_float(r[8 * c]), __uint_as_float(r[8 * c + 1]));
        q.y = pack_half2(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3]));
        q.z = pack_half2(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5]));
        q.w = pack_half2(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7]));
        *reinterpret_cast<uint4*>(dst + 8 * c) = q;
      }
    }
    if (p.dbias != nullptr) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      warp_colsum_atomic<32>(v, key < p.T, 1.0f, p.dbias + (which == 0 ? p.H : 2 * p.H) + h * AF_D + half * 32);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int attention_bwd_fused_launch(const void* qkv16, const void* o16, const void* do16, const float* lse, void* dqkv16, int B,
                               int T, int H, int heads, uint32_t drop_thr, float drop_inv_keep, uint64_t drop_seed,
                               float qscale, float* dbias, cudaStream_t stream) {
  AttnBwdFusedParams p;
  const int TK = (T + 15) / 16 * 16;
  W2V2_REQUIRE(TK <= AF_MAX_TK, "attention_bwd_fused: T=%d exceeds %d", T, AF_MAX_TK);
  const uint64_t qkv_row = uint64_t(3 * H) * 2, qkv_b = uint64_t(T) * 3 * H * 2;
  int rc = make_tmap_3d(&p.tmQ0, qkv16, 2, 3 * H, T, B, qkv_row, qkv_b, AF_D, 128, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmQ1, qkv16, 2, 3 * H, T, B, qkv_row, qkv_b, AF_D, 32, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmKV, qkv16, 2, 3 * H, T, B, qkv_row, qkv_b, AF_D, TK, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmDO0, do16, 2, H, T, B, uint64_t(H) * 2, uint64_t(T) * H * 2, AF_D, 128, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmDO1, do16, 2, H, T, B, uint64_t(H) * 2, uint64_t(T) * H * 2, AF_D, 32, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmO0, o16, 2, H, T, B, uint64_t(H) * 2, uint64_t(T) * H * 2, AF_D, 128, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmO1, o16, 2, H, T, B, uint64_t(H) * 2, uint64_t(T) * H * 2, AF_D, 32, 1, 128);
  if (rc) return rc;
  p.lse = lse;
  p.dqkv = static_cast<__half*>(dqkv16);
  p.T = T; p.TK = TK; p.H = H; p.heads = heads;
  p.qtiles = (T + 127) / 128;
  p.qscale = qscale;
  p.dbias = dbias;
  p.drop_thr = drop_thr;
  p.drop_inv_keep = drop_inv_keep;
  p.drop_seed = drop_seed;
  static bool configured = false;
  if (!configured) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM));
    configured = true;
  }
  dim3 grid(heads, B);
  W2V2_CHECK_CUDA(launch_k(attention_bwd_fused_kernel, grid, dim3(AF_THREADS), size_t(AF_SMEM), stream, 1, p));
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace w2v2
