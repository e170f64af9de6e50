// Self-attention backward for T <= 160 frames (3 s utterances: T = 149) as a PERSISTENT kernel -- same math as
// attention_bwd_fused.cu (backward of HF:438-463):
//     P = exp(S - lse), S = Q K^T;  dP = dO V^T;  dS = P * (dP - delta);  dV = P^T dO;  dQ = dS K;  dK = dS^T Q
// and the same three MMA round trips per (batch, head) problem, but
//   * one CTA per SM walks its share of the B x heads problems: barrier init / TMEM allocation / tensor-map fetches
//     happen once, and the TMA loads of problem k+1 are issued the moment the last MMAs of problem k have retired,
//     so they land while dQ / dK / dV of problem k drain from TMEM to global memory;
//   * the arithmetic pass is lean: no per-element bounds selects (columns >= T of S and dP are exact zeros times
//     zero-filled K / dO rows -- whatever the pass writes there only reaches accumulator rows that are never stored),
//     dropout is a template parameter, dS uses the fp32 probabilities: ~11 issue slots per pair of logits without
//     dropout, ~30 with (was ~90);
//   * the second query tile (rows 128 .. T-1, at most 32) is loaded three times, once per TMEM lane quarter 0..2, so
//     that its arithmetic runs on twelve warps / three SM sub-partitions (one 16-key chunk each) instead of on the
//     four warps that can read lanes 0..31;
//   * the q / k / v bias gradients (column sums over 32 rows per warp) use a halving butterfly: 31 shuffles per
//     32 columns instead of 160.
// 16 warps: warp & 3 = TMEM lane quarter (32 query rows), warp >> 2 = one of four column groups.
// TMEM (fp32 columns): S [0,160) | dP [160,320) | dQ tile 0 [320,384) | dQ tile 1 [384,448); at the end
// dV [0,128) | dK [128,256).
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 int swizzle_bytes);
int device_sm_count();

constexpr int BP_D = 64;
constexpr int BP_THREADS = 512;
constexpr int BP_MAX_TK = 160;
constexpr int BP_COL_S = 0, BP_COL_DP = 160, BP_COL_DQ0 = 320, BP_COL_DQ1 = 384, BP_COL_DV = 0, BP_COL_DK = 128;
constexpr int BP_BLK0 = 16384;                   // tile 0: 128 rows per 64-key block
constexpr int BP_BLK1 = 4096;                    // tile 1: 32 rows per 64-key block
constexpr int BP_P0 = 0;
constexpr int BP_DS0 = BP_P0 + 3 * BP_BLK0;
constexpr int BP_P1 = BP_DS0 + 3 * BP_BLK0;
constexpr int BP_DS1 = BP_P1 + 3 * BP_BLK1;
constexpr int BP_Q0 = BP_DS1 + 3 * BP_BLK1;
constexpr int BP_DO0 = BP_Q0 + 16384;
constexpr int BP_Q1 = BP_DO0 + 16384;            // 3 copies of the 32-row box (lane quarters 0..2)
constexpr int BP_DO1 = BP_Q1 + 3 * 4096;         // 3 copies
constexpr int BP_K = BP_DO1 + 3 * 4096;
constexpr int BP_V = BP_K + BP_MAX_TK * 128;
constexpr int BP_RED = BP_V + BP_MAX_TK * 128;   // float [4 groups][128 rows] + [4 groups][32 rows] partial deltas
constexpr int BP_BARS = BP_RED + (4 * 128 + 4 * 32) * 4;
constexpr int BP_SMEM = BP_BARS + 64;
// 128-row operand windows over the 32-row tile-1 buffers read (never use) bytes behind them; all of them end inside
// the allocation: Q1 -> DO1, DO1 -> K, P1 / dS1 (last key block + 16 KB) -> Q0 / DO0.
static_assert(BP_DO1 + 16384 <= BP_SMEM && BP_DS1 + 2 * BP_BLK1 + 16384 <= BP_SMEM, "operand windows leave the allocation");
static_assert(BP_SMEM + 1024 <= 227 * 1024, "shared memory budget");

struct alignas(64) AttnBwdPersistParams {
  CUtensorMap tmQ0;    // qkv: box {64, 128, 1}
  CUtensorMap tmQ1;    // qkv: box {64, 32, 1}
  CUtensorMap tmKV;    // qkv: box {64, TK, 1}
  CUtensorMap tmDO0;   // dO : box {64, 128, 1}
  CUtensorMap tmDO1;   // dO : box {64, 32, 1}
  CUtensorMap tmO0;    // O  : box {64, 128, 1}
  CUtensorMap tmO1;    // O  : box {64, 32, 1}
  const float* lse;    // [B, heads, T]
  __half* dqkv;        // [B*T, 3H]
  int T, TK, H, heads, qtiles, nprob;
  float qscale;        // dq is multiplied by this (chain rule of the d^-0.5 folded into Wq); 1 = leave as is
  float* dbias;        // f32 [3H] (+)= column sums of (dq * qscale | dk | dv), or nullptr
  uint32_t drop_thr;
  float drop_inv_keep;
  unsigned long long drop_seed;
};

constexpr float BP_L2E = 1.4426950408889634f;

// One 8-key chunk of a query row: S, dP (fp32 bits from TMEM) -> P (dropped, if DROP) and dS as 8 fp16 each.
template <bool DROP>
__device__ __forceinline__ void bwd_chunk(const uint32_t (&rs)[8], const uint32_t (&rd)[8], float lse2, float delta, DropKeys dk,
                                          uint32_t pair0, uint32_t thr_hi, float inv_keep, uint4& p_out, uint4& ds_out) {
  uint32_t pk[4], dsk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float e0 = fast_ex2(fmaf(__uint_as_float(rs[2 * j]), BP_L2E, -lse2));
    const float e1 = fast_ex2(fmaf(__uint_as_float(rs[2 * j + 1]), BP_L2E, -lse2));
    float d0 = __uint_as_float(rd[2 * j]), d1 = __uint_as_float(rd[2 * j + 1]);
    if (DROP) {
      // dP arrives for the dropped probabilities: the same keep / rescale applies to it
      const uint32_t hb = dropout_hash32(dk, pair0 + uint32_t(j));
      const bool k0 = (hb << 16) >= thr_hi, k1 = hb >= thr_hi;
      pk[j] = pack_half2(k0 ? e0 * inv_keep : 0.f, k1 ? e1 * inv_keep : 0.f);
      d0 = k0 ? d0 * inv_keep : 0.f;
      d1 = k1 ? d1 * inv_keep : 0.f;
    } else {
      pk[j] = pack_half2(e0, e1);
    }
    dsk[j] = pack_half2(e0 * (d0 - delta), e1 * (d1 - delta));
  }
  p_out = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  ds_out = make_uint4(dsk[0], dsk[1], dsk[2], dsk[3]);
}

__device__ __forceinline__ void store_row_f16x16_bp(__half* dst, const float (&v)[16]) {
  uint4 a, b;
  a.x = pack_half2(v[0], v[1]);   a.y = pack_half2(v[2], v[3]);   a.z = pack_half2(v[4], v[5]);   a.w = pack_half2(v[6], v[7]);
  b.x = pack_half2(v[8], v[9]);   b.y = pack_half2(v[10], v[11]); b.z = pack_half2(v[12], v[13]); b.w = pack_half2(v[14], v[15]);
  reinterpret_cast<uint4*>(dst)[0] = a;
  reinterpret_cast<uint4*>(dst)[1] = b;
}

template <bool DROP>
__global__ void __launch_bounds__(BP_THREADS, 1) attention_bwd_persist_kernel(const __grid_constant__ AttnBwdPersistParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  float* red0 = reinterpret_cast<float*>(smem + BP_RED);        // [4][128]
  float* red1 = red0 + 4 * 128;                                 // [4][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BP_BARS);
  uint64_t* bar_tma = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int TK = p.TK, T = p.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, cg = warp >> 2;
  const int row = quarter * 32 + lane;
  const bool two = p.qtiles > 1;
  pdl_trigger();

  if (threadIdx.x == 0) {
    prefetch_tensormap(&p.tmQ0);
    prefetch_tensormap(&p.tmKV);
    prefetch_tensormap(&p.tmDO0);
    prefetch_tensormap(&p.tmO0);
    if (two) {
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

  const int nloc = (p.nprob - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  auto issue_load = [&](int k) {                                     // thread 0; every smem operand buffer is free
    const int prob = int(blockIdx.x) + k * int(gridDim.x);
    const int b = prob / p.heads, h = prob - b * p.heads;
    mbar_arrive_expect_tx(bar_tma, 3 * 16384 + (two ? 7 * 4096 : 0) + 2 * TK * 128);
    tma_load_3d(smem + BP_Q0, &p.tmQ0, bar_tma, h * BP_D, 0, b);
    tma_load_3d(smem + BP_DO0, &p.tmDO0, bar_tma, h * BP_D, 0, b);
    tma_load_3d(smem + BP_K, &p.tmKV, bar_tma, p.H + h * BP_D, 0, b);
    tma_load_3d(smem + BP_V, &p.tmKV, bar_tma, 2 * p.H + h * BP_D, 0, b);
    tma_load_3d(smem + BP_DS0, &p.tmO0, bar_tma, h * BP_D, 0, b);                  // O tile 0 (borrowed buffer)
    if (two) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        tma_load_3d(smem + BP_Q1 + r * 4096, &p.tmQ1, bar_tma, h * BP_D, 128, b);
        tma_load_3d(smem + BP_DO1 + r * 4096, &p.tmDO1, bar_tma, h * BP_D, 128, b);
      }
      tma_load_3d(smem + BP_DS0 + 16384, &p.tmO1, bar_tma, h * BP_D, 128, b);     // O tile 1
    }
  };
  if (threadIdx.x == 0 && nloc > 0) issue_load(0);

  const int nch8 = TK / 8, nchunk16 = TK / 16;
  const int c_begin = (nch8 * cg) >> 2, c_end = (nch8 * (cg + 1)) >> 2;
  const int ktiles = (TK + 127) / 128;
  uint32_t mma_phase = 0;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t aK = sbase + BP_K, aV = sbase + BP_V;
  const DropKeys dkeys = drop_keys(p.drop_seed);
  const uint32_t thr_hi = p.drop_thr << 16;
  const float inv_keep = p.drop_inv_keep;
  const uint32_t idesc_s = make_idesc_f16(128, TK);
  const uint32_t idesc_dq = make_idesc_f16(128, BP_D, 0, 1);        // B (= K) MN-major
  const uint32_t idesc_t = make_idesc_f16(128, BP_D, 1, 1);         // A (P^T / dS^T) and B (dO / Q) MN-major
  // tile 1: warp slot 0..11 (lane quarters 0..2) owns the 16-key chunk of the same number
  const int slot1 = quarter * 4 + cg;
  const bool has1 = two && quarter < 3 && slot1 < nchunk16;

  auto issue_s_dp = [&](int qt) {
    const uint32_t aQ = sbase + (qt == 0 ? BP_Q0 : BP_Q1), adO = sbase + (qt == 0 ? BP_DO0 : BP_DO1);
#pragma unroll
    for (int k = 0; k < BP_D / 16; ++k)
      umma_f16(tmem + BP_COL_S, make_desc_k_sw128(aQ + k * 32), make_desc_k_sw128(aK + k * 32), idesc_s, k != 0);
#pragma unroll
    for (int k = 0; k < BP_D / 16; ++k)
      umma_f16(tmem + BP_COL_DP, make_desc_k_sw128(adO + k * 32), make_desc_k_sw128(aV + k * 32), idesc_s, k != 0);
  };
  auto issue_dq = [&](int qt) {
    const uint32_t adS = sbase + (qt == 0 ? BP_DS0 : BP_DS1);
    const uint32_t blk = qt == 0 ? BP_BLK0 : BP_BLK1;
    for (int kk = 0; kk < nchunk16; ++kk) {
      const uint64_t adesc = make_desc_k_sw128(adS + (kk >> 2) * blk + (kk & 3) * 32);
      const uint64_t bdesc = make_smem_desc(aK + kk * 2048, 16, 1024, 2);
      umma_f16(tmem + (qt == 0 ? BP_COL_DQ0 : BP_COL_DQ1), adesc, bdesc, idesc_dq, kk != 0);
    }
  };
  auto issue_dv_dk = [&](int qt, bool first) {
    const uint32_t aP = sbase + (qt == 0 ? BP_P0 : BP_P1), adS = sbase + (qt == 0 ? BP_DS0 : BP_DS1);
    const uint32_t aQ = sbase + (qt == 0 ? BP_Q0 : BP_Q1), adO = sbase + (qt == 0 ? BP_DO0 : BP_DO1);
    const uint32_t blk = qt == 0 ? BP_BLK0 : BP_BLK1;
    const int qsteps = min(8, (T - qt * 128 + 15) >> 4);
    for (int kt = 0; kt < ktiles; ++kt) {
      for (int ks = 0; ks < qsteps; ++ks) {
        const uint64_t a1 = make_smem_desc(aP + kt * 2 * blk + ks * 2048, blk, 1024, 2);
        const uint64_t b1 = make_smem_desc(adO + ks * 2048, 16, 1024, 2);
        umma_f16(tmem + BP_COL_DV + kt * BP_D, a1, b1, idesc_t, !(first && ks == 0));
      }
      for (int ks = 0; ks < qsteps; ++ks) {
        const uint64_t a2 = make_smem_desc(adS + kt * 2 * blk + ks * 2048, blk, 1024, 2);
        const uint64_t b2 = make_smem_desc(aQ + ks * 2048, 16, 1024, 2);
        umma_f16(tmem + BP_COL_DK + kt * BP_D, a2, b2, idesc_t, !(first && ks == 0));
      }
    }
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

  for (int k = 0; k < nloc; ++k) {
    const int prob = int(blockIdx.x) + k * int(gridDim.x);
    const int b = prob / p.heads, h = prob - b * p.heads;
    const uint32_t bh = uint32_t(b) * p.heads + h;
    const int t1 = 128 + lane;
    // lse: fetched now, first used in the arithmetic passes
    float lse0 = 0.f, lse1 = 0.f;
    if (row < T) lse0 = __ldg(p.lse + int64_t(bh) * T + row);
    if (has1 && t1 < T) lse1 = __ldg(p.lse + int64_t(bh) * T + t1);

    mbar_wait(bar_tma, k & 1);
    __syncwarp();
    tc_fence_after();
    // ---- round 1: S0, dP0 -- issued first, delta is formed while they run
    if (threadIdx.x == 0) {
      issue_s_dp(0);
      umma_commit(bar_mma);
    }
    __syncwarp();
    // ---- delta = rowsum(dO * O) from shared memory: this thread's 16 of the 64 head dims (two 16-byte chunks of
    //      the swizzled row), combined across the four column groups
    {
      float part = 0.f;
      const uint8_t* orow = smem + BP_DS0 + row * 128;
      const uint8_t* grow = smem + BP_DO0 + row * 128;
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
      red0[cg * 128 + row] = part;
      if (two && quarter == 0) {              // tile 1: 32 rows
        float part1 = 0.f;
        const uint8_t* orow1 = smem + BP_DS0 + 16384 + lane * 128;
        const uint8_t* grow1 = smem + BP_DO1 + lane * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int off = ((2 * cg + c) ^ (lane & 7)) << 4;
          const uint4 a = *reinterpret_cast<const uint4*>(orow1 + off), g = *reinterpret_cast<const uint4*>(grow1 + off);
          const __half2* ah = reinterpret_cast<const __half2*>(&a);
          const __half2* gh = reinterpret_cast<const __half2*>(&g);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 x = __half22float2(ah[j]), y = __half22float2(gh[j]);
            part1 = fmaf(x.x, y.x, part1);
            part1 = fmaf(x.y, y.y, part1);
          }
        }
        red1[cg * 32 + lane] = part1;
      }
    }
    __syncthreads();           // also: every read of the borrowed O tiles precedes the first dS write
    const float delta0 = (red0[row] + red0[128 + row]) + (red0[256 + row] + red0[384 + row]);
    const float delta1 = two ? (red1[lane] + red1[32 + lane]) + (red1[64 + lane] + red1[96 + lane]) : 0.f;

    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    __syncwarp();
    tc_fence_after();
    // ---- pass 0: S0, dP0 (TMEM) -> P0, dS0 (smem)
    if (quarter * 32 < T) {
      const float lse2 = lse0 * BP_L2E;
      const uint32_t pair_row = (bh * T + (row < T ? row : 0)) * uint32_t(TK / 2);
      uint8_t* prow = smem + BP_P0 + row * 128;
      uint8_t* dsrow = smem + BP_DS0 + row * 128;
      // software pipeline over the (at most five) 8-key chunks: chunk i+1 is in flight from TMEM while chunk i is computed
      uint32_t rs[2][8], rd[2][8];
      tmem_ld_32x32b_x8(t_row + BP_COL_S + c_begin * 8, rs[0]);
      tmem_ld_32x32b_x8(t_row + BP_COL_DP + c_begin * 8, rd[0]);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int c = c_begin + i;
        if (c < c_end) {
          tmem_ld_wait();
          if (c + 1 < c_end) {
            tmem_ld_32x32b_x8(t_row + BP_COL_S + (c + 1) * 8, rs[(i + 1) & 1]);
            tmem_ld_32x32b_x8(t_row + BP_COL_DP + (c + 1) * 8, rd[(i + 1) & 1]);
          }
          uint4 po, dso;
          bwd_chunk<DROP>(rs[i & 1], rd[i & 1], lse2, delta0, dkeys, pair_row + c * 4, thr_hi, inv_keep, po, dso);
          const int off = (c >> 3) * BP_BLK0 + (((c & 7) ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(prow + off) = po;
          *reinterpret_cast<uint4*>(dsrow + off) = dso;
        }
      }
    }
    publish_smem();

    // 16 of the 64 dQ columns of this thread's row -> global (x qscale), and their share of the q-bias gradient
    auto drain_dq = [&](int qt) {
      const int t = qt * 128 + row;
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_row + (qt == 0 ? BP_COL_DQ0 : BP_COL_DQ1) + cg * 16, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * p.qscale;
      if (t < T) store_row_f16x16_bp(p.dqkv + (int64_t(b) * T + t) * 3 * p.H + h * BP_D + cg * 16, v);
      if (p.dbias != nullptr) {
        if (t >= T) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        const float s = warp_colsum_bfly<16>(v);
        if (lane < 16) atomicAdd(p.dbias + h * BP_D + cg * 16 + lane, s);
      }
    };

    if (two) {
      // ---- round 2: dQ0 ; S1, dP1 (the S / dP regions were fully consumed by pass 0)
      if (threadIdx.x == 0) {
        issue_dq(0);
        issue_s_dp(1);
      }
      commit_and_wait();
      if (has1) {                                           // one 16-key chunk of tile-1 row `lane`
        const float lse2 = lse1 * BP_L2E;
        const uint32_t pair_row = (bh * T + (t1 < T ? t1 : 0)) * uint32_t(TK / 2);
        uint8_t* prow = smem + BP_P1 + lane * 128;
        uint8_t* dsrow = smem + BP_DS1 + lane * 128;
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          const int c = slot1 * 2 + hc;
          uint32_t rs[8], rd[8];
          tmem_ld_32x32b_x8(t_row + BP_COL_S + c * 8, rs);
          tmem_ld_32x32b_x8(t_row + BP_COL_DP + c * 8, rd);
          tmem_ld_wait();
          uint4 po, dso;
          bwd_chunk<DROP>(rs, rd, lse2, delta1, dkeys, pair_row + c * 4, thr_hi, inv_keep, po, dso);
          const int off = (c >> 3) * BP_BLK1 + (((c & 7) ^ (lane & 7)) << 4);
          *reinterpret_cast<uint4*>(prow + off) = po;
          *reinterpret_cast<uint4*>(dsrow + off) = dso;
        }
      }
      if (quarter * 32 < T) drain_dq(0);
      publish_smem();
      // ---- round 3: dQ1 ; dV, dK over all query rows (they overwrite the S / dP regions)
      if (threadIdx.x == 0) {
        issue_dq(1);
        issue_dv_dk(0, true);
        issue_dv_dk(1, false);
      }
      commit_and_wait();
      if (threadIdx.x == 0 && k + 1 < nloc) issue_load(k + 1);      // lands while dQ1 / dK / dV drain
      if (quarter == 0) drain_dq(1);
    } else {
      if (threadIdx.x == 0) {
        issue_dq(0);
        issue_dv_dk(0, true);
      }
      commit_and_wait();
      if (threadIdx.x == 0 && k + 1 < nloc) issue_load(k + 1);
      if (quarter * 32 < T) drain_dq(0);
    }

    // ---- dK / dV rows (keys) -> global: column group -> (dK | dV, 32-column half)
    for (int kt = 0; kt < ktiles; ++kt) {
      if (kt * 128 + quarter * 32 >= T) continue;       // warp-uniform
      const int key = kt * 128 + row;
      const int which = cg >> 1, half = cg & 1;           // 0: dK, 1: dV
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + (which == 0 ? BP_COL_DK : BP_COL_DV) + kt * BP_D + half * 32, r);
      tmem_ld_wait();
      if (key < T) {
        __half* dst = p.dqkv + (int64_t(b) * T + key) * 3 * p.H + h * BP_D + (which == 0 ? p.H : 2 * p.H) + half * 32;
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
        for (int j = 0; j < 32; ++j) v[j] = key < T ? __uint_as_float(r[j]) : 0.f;
        const float s = warp_colsum_bfly<32>(v);
        atomicAdd(p.dbias + (which == 0 ? p.H : 2 * p.H) + h * BP_D + half * 32 + lane, s);
      }
    }
    // the next problem's S / dP overwrite the dV / dK columns: every warp has drained them
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// -> 0 launched, 1 not applicable (caller falls back to attention_bwd_fused.cu), < 0 error
int attention_bwd_persist_launch(const void* qkv16, const void* o16, const void* do16, const float* lse, void* dqkv16, int B,
                                 int T, int H, int heads, uint32_t drop_thr, float drop_inv_keep, uint64_t drop_seed,
                                 float qscale, float* dbias, cudaStream_t stream) {
  static const bool on = []() { const char* e = getenv("W2V2_ATTN_PERSIST"); return !(e != nullptr && e[0] == '0'); }();
  const int TK = (T + 15) / 16 * 16;
  if (!on || TK > BP_MAX_TK) return 1;
  AttnBwdPersistParams p;
  const uint64_t qkv_row = uint64_t(3 * H) * 2, qkv_b = uint64_t(T) * 3 * H * 2;
  int rc = make_tmap_3d(&p.tmQ0, qkv16, 2, 3 * H, T, B, qkv_row, qkv_b, BP_D, 128, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmQ1, qkv16, 2, 3 * H, T, B, qkv_row, qkv_b, BP_D, 32, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmKV, qkv16, 2, 3 * H, T, B, qkv_row, qkv_b, BP_D, TK, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmDO0, do16, 2, H, T, B, uint64_t(H) * 2, uint64_t(T) * H * 2, BP_D, 128, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmDO1, do16, 2, H, T, B, uint64_t(H) * 2, uint64_t(T) * H * 2, BP_D, 32, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmO0, o16, 2, H, T, B, uint64_t(H) * 2, uint64_t(T) * H * 2, BP_D, 128, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmO1, o16, 2, H, T, B, uint64_t(H) * 2, uint64_t(T) * H * 2, BP_D, 32, 1, 128);
  if (rc) return rc;
  p.lse = lse;
  p.dqkv = static_cast<__half*>(dqkv16);
  p.T = T; p.TK = TK; p.H = H; p.heads = heads;
  p.qtiles = (T + 127) / 128;
  p.nprob = B * heads;
  p.qscale = qscale;
  p.dbias = dbias;
  p.drop_thr = drop_thr;
  p.drop_inv_keep = drop_inv_keep;
  p.drop_seed = drop_seed;
  static bool configured = false;
  if (!configured) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BP_SMEM));
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BP_SMEM));
    configured = true;
  }
  const int sms = device_sm_count();
  const int grid = p.nprob < sms ? p.nprob : sms;
  if (drop_thr != 0)
    W2V2_CHECK_CUDA(launch_k(attention_bwd_persist_kernel<true>, dim3(grid), dim3(BP_THREADS), size_t(BP_SMEM), stream, 1, p));
  else
    W2V2_CHECK_CUDA(launch_k(attention_bwd_persist_kernel<false>, dim3(grid), dim3(BP_THREADS), size_t(BP_SMEM), stream, 1, p));
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace w2v2
