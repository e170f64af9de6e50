// Self-attention core for training crops, T <= 160 frames (3 s -> 149), as a PERSISTENT kernel (HF:438-463):
//     O[b,t,h,:] = softmax_j( q[b,t,h,:] . k[b,j,h,:] ) v[b,j,h,:]          (1/sqrt(d) folded into Wq)
//
// attention.cu runs one CTA per (batch, head, 128-query tile): 1536 short-lived CTAs whose TMA -> MMA -> softmax ->
// MMA chain is latency plus ~45 issue slots per logit (per-element bounds / dropout selects, fp16 round trips).  Here
// one CTA per SM walks its share of the B x heads (batch, head) problems:
//   * Q / K / V of problems k+1 and k+2 are in flight (two TMA stages) while problem k is computed;
//   * S = Q K^T of problem k+1 is issued right behind O = P V of problem k, so it completes under k's epilogue:
//     the only exposed tensor-core round trip per problem is the P V one;
//   * both query tiles of a problem are handled in one pass.  The second tile of a 149-frame utterance holds 21 rows;
//     a warp can only read its own 32 TMEM lanes, so a plain second tile would put all of its arithmetic on the four
//     warps (one SM sub-partition) that own lanes 0..31.  Its 32-row Q box is therefore loaded FOUR times, once per
//     lane quarter: the M = 128 MMA then leaves the same 32 logit rows in every quarter, the 8-key chunks of those
//     rows are dealt to all 16 warps, and the row max / sum meet through shared memory;
//   * each thread keeps its logits in registers between the max and the exp pass (one TMEM read per element), and
//     the inner loops are specialised: full 8-key chunks carry no per-element bounds checks, dropout is a template
//     parameter -- ~5 issue slots per logit without dropout, ~13 with.
// 16 warps: warp & 3 = TMEM lane quarter (32 query rows), warp >> 2 = one of four column groups.  The softmax
// normaliser is the fp32 sum of the un-rounded, un-dropped exponentials (HF applies dropout after the softmax).
// Dropout masks are the same function of (seed, b, h, t, key pair) as in attention.cu: the backward kernels
// regenerate them unchanged.
// TMEM (fp32 columns): S tile 0 [0,160) | S tile 1 [160,320) | O tile 0 [320,384) | O tile 1 [384,448).
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 int swizzle_bytes);
int device_sm_count();

constexpr int AP_D = 64;
constexpr int AP_THREADS = 512;
constexpr int AP_MAX_TK = 160;
constexpr int AP_COL_S0 = 0, AP_COL_S1 = 160, AP_COL_O0 = 320, AP_COL_O1 = 384;
constexpr int AP_MAXCH = 5;                      // 8-key chunks per thread in tile 0: ceil(160 / 8 / 4)
// shared memory map (bytes); kvb = round_up(TK * 128, 1024)
constexpr int AP_Q0 = 0;                         // 2 stages x 16 KB
constexpr int AP_Q1 = 32768;                     // 2 stages x 4 copies x 4 KB (the 32-row box, once per lane quarter)
constexpr int AP_K = 65536;                      // 2 stages x kvb, then V: 2 stages x kvb, then P1, P0, exchange, barriers
constexpr int AP_P1_BYTES = 3 * 4096;            // tile 1: 32 rows per 64-key block
constexpr int AP_P0_BYTES = 3 * 16384;           // tile 0: 128 rows per 64-key block
constexpr int AP_RED_BYTES = (4 * 128 + 16 * 32 + 2 * 4 * 128 + 2 * 16 * 32) * 4;
// The 128-row A-operand window of the tile-1 P V product starts at P1 and ends up to 20 KB behind it: inside P0.

struct alignas(64) AttnPersistParams {
  CUtensorMap tmQ0;   // qkv: box {64, 128, 1}
  CUtensorMap tmQ1;   // qkv: box {64, 32, 1}
  CUtensorMap tmKV;   // qkv: box {64, TK, 1}
  __half* out;
  float* lse;         // [B, heads, T] or nullptr
  const int* lens;    // [B] valid KEYS per utterance (ragged evaluation batches: frames >= lens[b] are padding) or nullptr
  int T, TK, H, heads, nprob, ntiles, kvb;
  uint32_t drop_thr;
  float drop_inv_keep;
  unsigned long long drop_seed;
};

constexpr float AP_L2E = 1.4426950408889634f;

// exp of one 8-key chunk of a row: v = logits (fp32 bits), mxl = row max * log2(e).  Returns the chunk's share of the
// softmax normaliser (fp32, before dropout and before the fp16 rounding) and the 8 fp16 probabilities in `out`.
// FULL: all 8 keys exist (no bounds checks); otherwise keys j >= nv are zero.  pair0 = dropout pair index of key 0.
template <bool FULL, bool DROP>
__device__ __forceinline__ float exp_chunk(const uint32_t (&v)[8], float mxl, int nv, DropKeys dk, uint32_t pair0,
                                           uint32_t thr_hi, float inv_keep, uint4& out) {
  float sum = 0.f;
  uint32_t pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float e0 = fast_ex2(fmaf(__uint_as_float(v[2 * j]), AP_L2E, -mxl));
    float e1 = fast_ex2(fmaf(__uint_as_float(v[2 * j + 1]), AP_L2E, -mxl));
    if (!FULL) {
      if (2 * j >= nv) e0 = 0.f;
      if (2 * j + 1 >= nv) e1 = 0.f;
    }
    sum += e0 + e1;
    if (DROP) {
      const uint32_t hb = dropout_hash32(dk, pair0 + uint32_t(j));
      // low 16 bits decide the even key, high 16 bits the odd key: keep <=> bits >= thr
      e0 = (hb << 16) >= thr_hi ? e0 * inv_keep : 0.f;
      e1 = hb >= thr_hi ? e1 * inv_keep : 0.f;
    }
    pk[j] = pack_half2(e0, e1);
  }
  out = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  return sum;
}

template <bool FULL>
__device__ __forceinline__ float max_chunk(const uint32_t (&v)[8], int nv, float mx) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (FULL || j < nv) mx = fmaxf(mx, __uint_as_float(v[j]));
  return mx;
}

template <bool DROP>
__global__ void __launch_bounds__(AP_THREADS, 1) attention_persist_kernel(const __grid_constant__ AttnPersistParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int TK = p.TK, T = p.T, kvb = p.kvb;
  const int oV = AP_K + 2 * kvb, oP1 = oV + 2 * kvb, oP0 = oP1 + AP_P1_BYTES, oRed = oP0 + AP_P0_BYTES;
  float* red_max0 = reinterpret_cast<float*>(smem + oRed);      // [4][128]
  float* red_max1 = red_max0 + 4 * 128;                         // [16][32]
  float* red_sum0 = red_max1 + 16 * 32;                         // [2][4][128]
  float* red_sum1 = red_sum0 + 2 * 4 * 128;                     // [2][16][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oRed + AP_RED_BYTES);
  uint64_t* bar_full = bars;                                    // [2]
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_o = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, cg = warp >> 2;
  const int row = quarter * 32 + lane;
  pdl_trigger();

  if (threadIdx.x == 0) {
    prefetch_tensormap(&p.tmQ0);
    prefetch_tensormap(&p.tmQ1);
    prefetch_tensormap(&p.tmKV);
    mbar_init(bar_full, 1);
    mbar_init(bar_full + 1, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
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
  const uint32_t sbase = smem_u32(smem);
  const int nch8 = TK / 8;
  const int c_begin = (nch8 * cg) >> 2, c_end = (nch8 * (cg + 1)) >> 2;
  const int nchunk16 = TK / 16;
  const bool two = p.ntiles > 1;
  const DropKeys dkeys = drop_keys(p.drop_seed);
  const uint32_t thr_hi = p.drop_thr << 16;
  const float inv_keep = p.drop_inv_keep;
  const uint32_t idesc_s = make_idesc_f16(128, TK);
  const uint32_t idesc_o = make_idesc_f16(128, AP_D, 0, 1);          // B (= V) MN-major
  // tile 1: the chunks of this warp (every quarter holds the same 32 rows)
  const int c1a = warp, c1b = warp + 16;
  const bool has1a = two && c1a < nch8, has1b = two && c1b < nch8;

  auto issue_load = [&](int k) {                                     // thread 0
    const int s = k & 1, prob = int(blockIdx.x) + k * int(gridDim.x);
    const int b = prob / p.heads, h = prob - b * p.heads;
    mbar_arrive_expect_tx(bar_full + s, 16384 + (two ? 4 * 4096 : 0) + 2 * TK * 128);
    tma_load_3d(smem + AP_Q0 + s * 16384, &p.tmQ0, bar_full + s, h * AP_D, 0, b);
    tma_load_3d(smem + AP_K + s * kvb, &p.tmKV, bar_full + s, p.H + h * AP_D, 0, b);
    tma_load_3d(smem + oV + s * kvb, &p.tmKV, bar_full + s, 2 * p.H + h * AP_D, 0, b);
    if (two) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        tma_load_3d(smem + AP_Q1 + s * 16384 + r * 4096, &p.tmQ1, bar_full + s, h * AP_D, 128, b);
    }
  };
  auto issue_s = [&](int k) {                                        // thread 0, stage k & 1 has landed
    const int s = k & 1;
    const uint32_t aK = sbase + AP_K + s * kvb;
    const uint32_t aQ0 = sbase + AP_Q0 + s * 16384, aQ1 = sbase + AP_Q1 + s * 16384;
#pragma unroll
    for (int kk = 0; kk < AP_D / 16; ++kk)
      umma_f16(tmem + AP_COL_S0, make_desc_k_sw128(aQ0 + kk * 32), make_desc_k_sw128(aK + kk * 32), idesc_s, kk != 0);
    if (two) {
#pragma unroll
      for (int kk = 0; kk < AP_D / 16; ++kk)
        umma_f16(tmem + AP_COL_S1, make_desc_k_sw128(aQ1 + kk * 32), make_desc_k_sw128(aK + kk * 32), idesc_s, kk != 0);
    }
  };

  if (threadIdx.x == 0 && nloc > 0) {
    issue_load(0);
    if (nloc > 1) issue_load(1);
    mbar_wait(bar_full, 0);
    tc_fence_after();
    issue_s(0);
    umma_commit(bar_s);
  }
  __syncwarp();

  const bool do0 = quarter * 32 < T;                         // warp-uniform (T < 128: the upper quarters idle)
  for (int k = 0; k < nloc; ++k) {
    const int s = k & 1;
    const int prob = int(blockIdx.x) + k * int(gridDim.x);
    const int b = prob / p.heads, h = prob - b * p.heads;
    const uint32_t bh = uint32_t(b) * p.heads + h;
    const int tq1 = 128 + lane;
    // keys that exist for this utterance: every query row (padding rows included -- they must stay finite) attends to them
    const int Tk = p.lens != nullptr ? __ldg(p.lens + b) : T;

    mbar_wait(bar_s, k & 1);
    __syncwarp();
    tc_fence_after();
    // ---- logits -> registers, partial row maxima
    uint32_t v0[AP_MAXCH][8], v1[2][8];
    if (do0) {
#pragma unroll
      for (int i = 0; i < AP_MAXCH; ++i)
        if (c_begin + i < c_end) tmem_ld_32x32b_x8(t_row + AP_COL_S0 + (c_begin + i) * 8, v0[i]);
    }
    if (has1a) tmem_ld_32x32b_x8(t_row + AP_COL_S1 + c1a * 8, v1[0]);
    if (has1b) tmem_ld_32x32b_x8(t_row + AP_COL_S1 + c1b * 8, v1[1]);
    tmem_ld_wait();
    float mx0 = -INFINITY, mx1 = -INFINITY;
    if (do0) {
#pragma unroll
      for (int i = 0; i < AP_MAXCH; ++i) {
        const int c = c_begin + i;
        if (c < c_end) {
          const int nv = Tk - c * 8;
          mx0 = nv >= 8 ? max_chunk<true>(v0[i], 8, mx0) : max_chunk<false>(v0[i], nv, mx0);
        }
      }
      red_max0[cg * 128 + row] = mx0;
    }
    if (two) {
      if (has1a) { const int nv = Tk - c1a * 8; mx1 = nv >= 8 ? max_chunk<true>(v1[0], 8, mx1) : max_chunk<false>(v1[0], nv, mx1); }
      if (has1b) { const int nv = Tk - c1b * 8; mx1 = nv >= 8 ? max_chunk<true>(v1[1], 8, mx1) : max_chunk<false>(v1[1], nv, mx1); }
      red_max1[warp * 32 + lane] = mx1;
    }
    __syncthreads();
    // ---- exp, (dropout,) P -> smem, partial row sums
    if (do0) {
      mx0 = fmaxf(fmaxf(red_max0[row], red_max0[128 + row]), fmaxf(red_max0[256 + row], red_max0[384 + row]));
      const float mxl = mx0 * AP_L2E;
      const uint32_t pair_row = (bh * T + (row < T ? row : 0)) * uint32_t(TK / 2);
      uint8_t* prow = smem + oP0 + row * 128;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < AP_MAXCH; ++i) {
        const int c = c_begin + i;
        if (c < c_end) {
          const int nv = Tk - c * 8;
          uint4 o;
          sum += nv >= 8 ? exp_chunk<true, DROP>(v0[i], mxl, 8, dkeys, pair_row + c * 4, thr_hi, inv_keep, o)
                         : exp_chunk<false, DROP>(v0[i], mxl, nv, dkeys, pair_row + c * 4, thr_hi, inv_keep, o);
          *reinterpret_cast<uint4*>(prow + (c >> 3) * 16384 + (((c & 7) ^ (row & 7)) << 4)) = o;
        }
      }
      red_sum0[(s * 4 + cg) * 128 + row] = sum;
    }
    if (two) {
#pragma unroll
      for (int w = 0; w < 16; ++w) mx1 = fmaxf(mx1, red_max1[w * 32 + lane]);
      const float mxl = mx1 * AP_L2E;
      const uint32_t pair_row = (bh * T + (tq1 < T ? tq1 : 0)) * uint32_t(TK / 2);
      uint8_t* prow = smem + oP1 + lane * 128;
      float sum = 0.f;
      if (has1a) {
        const int nv = Tk - c1a * 8;
        uint4 o;
        sum += nv >= 8 ? exp_chunk<true, DROP>(v1[0], mxl, 8, dkeys, pair_row + c1a * 4, thr_hi, inv_keep, o)
                       : exp_chunk<false, DROP>(v1[0], mxl, nv, dkeys, pair_row + c1a * 4, thr_hi, inv_keep, o);
        *reinterpret_cast<uint4*>(prow + (c1a >> 3) * 4096 + (((c1a & 7) ^ (lane & 7)) << 4)) = o;
      }
      if (has1b) {
        const int nv = Tk - c1b * 8;
        uint4 o;
        sum += nv >= 8 ? exp_chunk<true, DROP>(v1[1], mxl, 8, dkeys, pair_row + c1b * 4, thr_hi, inv_keep, o)
                       : exp_chunk<false, DROP>(v1[1], mxl, nv, dkeys, pair_row + c1b * 4, thr_hi, inv_keep, o);
        *reinterpret_cast<uint4*>(prow + (c1b >> 3) * 4096 + (((c1b & 7) ^ (lane & 7)) << 4)) = o;
      }
      red_sum1[(s * 16 + warp) * 32 + lane] = sum;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t aV = sbase + oV + s * kvb;
      for (int kk = 0; kk < nchunk16; ++kk) {
        const uint64_t adesc = make_desc_k_sw128(sbase + oP0 + (kk >> 2) * 16384 + (kk & 3) * 32);
        const uint64_t bdesc = make_smem_desc(aV + kk * 2048, 16, 1024, 2);       // MN-major, 128B swizzle
        umma_f16(tmem + AP_COL_O0, adesc, bdesc, idesc_o, kk != 0);
      }
      if (two) {
        for (int kk = 0; kk < nchunk16; ++kk) {
          const uint64_t adesc = make_desc_k_sw128(sbase + oP1 + (kk >> 2) * 4096 + (kk & 3) * 32);
          const uint64_t bdesc = make_smem_desc(aV + kk * 2048, 16, 1024, 2);
          umma_f16(tmem + AP_COL_O1, adesc, bdesc, idesc_o, kk != 0);
        }
      }
      umma_commit(bar_o);
      if (k + 1 < nloc) {                        // next problem's logits run under this problem's epilogue
        mbar_wait(bar_full + ((k + 1) & 1), ((k + 1) >> 1) & 1);
        tc_fence_after();
        issue_s(k + 1);
        umma_commit(bar_s);
      }
    }
    __syncwarp();
    float sum0 = 1.f, sum1 = 1.f;
    if (do0) sum0 = (red_sum0[(s * 4 + 0) * 128 + row] + red_sum0[(s * 4 + 1) * 128 + row]) +
                    (red_sum0[(s * 4 + 2) * 128 + row] + red_sum0[(s * 4 + 3) * 128 + row]);
    const bool do1 = two && quarter == 0;        // O tile 1: its rows are lanes 0..31
    if (do1) {
      sum1 = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) sum1 += red_sum1[(s * 16 + w) * 32 + lane];
    }

    mbar_wait(bar_o, k & 1);
    __syncwarp();
    tc_fence_after();
    // stage s is free (the P V MMAs were its last readers): refill it two problems ahead
    if (threadIdx.x == 0 && k + 2 < nloc) issue_load(k + 2);
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      if (tile == 0 ? !do0 : !do1) continue;
      const int t_q = tile == 0 ? row : tq1;
      const float sum = tile == 0 ? sum0 : sum1, mx = tile == 0 ? mx0 : mx1;
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_row + (tile == 0 ? AP_COL_O0 : AP_COL_O1) + cg * 16, r);
      tmem_ld_wait();
      if (t_q < T) {
        const float inv = 1.0f / sum;
        if (cg == 0 && p.lse != nullptr) p.lse[int64_t(bh) * T + t_q] = mx + __logf(sum);
        __half* dst = p.out + (int64_t(b) * T + t_q) * p.H + h * AP_D + cg * 16;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint4 o;
          o.x = pack_half2(__uint_as_float(r[8 * c]) * inv, __uint_as_float(r[8 * c + 1]) * inv);
          o.y = pack_half2(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv);
          o.z = pack_half2(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv);
          o.w = pack_half2(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 8 * c) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// -> 0 launched, 1 not applicable (caller falls back to attention.cu), < 0 error
int attention_persist_launch(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, uint32_t drop_thr,
                             float drop_inv_keep, uint64_t drop_seed, const int* lens, cudaStream_t stream) {
  static const bool on = []() { const char* e = getenv("W2V2_ATTN_PERSIST"); return !(e != nullptr && e[0] == '0'); }();
  const int TK = (T + 15) / 16 * 16;
  if (!on || TK > AP_MAX_TK) return 1;
  AttnPersistParams p;
  const uint64_t row_b = uint64_t(3 * H) * 2, utt_b = uint64_t(T) * 3 * H * 2;
  int rc = make_tmap_3d(&p.tmQ0, qkv16, 2, 3 * H, T, B, row_b, utt_b, AP_D, 128, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmQ1, qkv16, 2, 3 * H, T, B, row_b, utt_b, AP_D, 32, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmKV, qkv16, 2, 3 * H, T, B, row_b, utt_b, AP_D, TK, 1, 128);
  if (rc) return rc;
  p.out = static_cast<__half*>(out16);
  p.lse = lse;
  p.lens = lens;
  p.T = T; p.TK = TK; p.H = H; p.heads = heads;
  p.nprob = B * heads;
  p.ntiles = T > 128 ? 2 : 1;
  p.kvb = (TK * 128 + 1023) & ~1023;
  p.drop_thr = drop_thr;
  p.drop_inv_keep = drop_inv_keep;
  p.drop_seed = drop_seed;
  const int smem = AP_K + 4 * p.kvb + AP_P1_BYTES + AP_P0_BYTES + AP_RED_BYTES + 64;
  static int configured_smem[2] = {0, 0};
  const int di = drop_thr != 0 ? 1 : 0;
  if (smem > configured_smem[di]) {
    if (di) W2V2_CHECK_CUDA(cudaFuncSetAttribute(attention_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else W2V2_CHECK_CUDA(cudaFuncSetAttribute(attention_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem[di] = smem;
  }
  const int sms = device_sm_count();
  const int grid = p.nprob < sms ? p.nprob : sms;
  if (di) W2V2_CHECK_CUDA(launch_k(attention_persist_kernel<true>, dim3(grid), dim3(AP_THREADS), size_t(smem), stream, 1, p));
  else W2V2_CHECK_CUDA(launch_k(attention_persist_kernel<false>, dim3(grid), dim3(AP_THREADS), size_t(smem), stream, 1, p));
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace w2v2
