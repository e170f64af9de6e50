// Self-attention core for training crops, T <= 160 frames: persistent kernel, TWO resident CTAs per SM (HF:438-463).
//
// attention_persist.cu (one 16-warp CTA per SM, 190 KB of operands and P tiles) spends 40 % of its warp-stall samples on
// the per-problem dependency chain: two CTA-wide barriers and two MMA completions per (batch, head) problem, with every
// warp in the same phase at the same time.  A second resident CTA hides that chain -- while one CTA waits for a TMA load,
// a barrier or an MMA, the other one computes -- but needs a CTA to fit in half an SM: <= 113 KB of shared memory,
// 256 TMEM columns.  This kernel gets there by
//   * writing P of the first query tile OVER the Q / K operands, which are dead once S = Q K^T has completed (P needs
//     48 KB, Q0 + the replicated Q1 + K occupy 52 KB);
//   * running the two query tiles one after the other through the same 160 TMEM columns (tile 1 first, while K is still
//     alive), O of tile 0 into the first 64 of them once the logits have been consumed: 224 columns in all;
//   * a single operand stage (the next problem's TMA loads are issued when the last P V product has completed and land
//     under the epilogue; what is left of their latency is the sibling CTA's to hide).
// 8 warps per CTA: warp & 3 = TMEM lane quarter (32 query rows), warp >> 2 = one of two column groups (80 keys each for
// tile 0).  The 32-row Q box of tile 1 is loaded four times (once per lane quarter) so that its logit rows appear in every
// quarter and all 8 warps share its 8-key chunks.  Softmax arithmetic, dropout masks and the normaliser are those of
// attention_persist.cu (same exp_chunk / max_chunk helpers), so the backward kernels regenerate the masks unchanged.
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 int swizzle_bytes);
int device_sm_count();

constexpr int A2_D = 64;
constexpr int A2_THREADS = 256;
constexpr int A2_MAX_TK = 160;
constexpr int A2_COL_S = 0, A2_COL_O0 = 0, A2_COL_O1 = 160;      // 224 of the 256 allocated columns
constexpr int A2_MAXCH = 10;                    // 8-key chunks per thread in tile 0: 160 / 8 / 2
// shared memory map (bytes)
constexpr int A2_Q0 = 0;                        // 16 KB   } dead after S = Q K^T: P of tile 0 (3 x 16 KB) is written
constexpr int A2_Q1 = 16384;                    // 4 x 4 KB }   over this region
constexpr int A2_K = 32768;                     // 20 KB   }
constexpr int A2_P0 = 0;
constexpr int A2_P1 = 53248;                    // 3 x 4 KB (32 rows per 64-key block); its 128-row operand window ends in V
constexpr int A2_V = 65536;                     // 20 KB
constexpr int A2_RED = 86016;                   // float [2][128] max0 | [8][32] max1 | [2][128] sum0 | [8][32] sum1
constexpr int A2_RED_BYTES = (2 * 128 + 8 * 32 + 2 * 128 + 8 * 32) * 4;
constexpr int A2_BARS = A2_RED + A2_RED_BYTES;
constexpr int A2_SMEM = A2_BARS + 64;
static_assert(A2_P0 + 3 * 16384 <= A2_P1 && A2_K + A2_MAX_TK * 128 <= A2_P1, "P0 / K leave their region");
static_assert(A2_P1 + 2 * 4096 + 16384 <= A2_V + A2_MAX_TK * 128, "P1 operand window leaves the allocation");
static_assert(2 * (A2_SMEM + 1024) <= 227 * 1024, "two CTAs per SM");

struct alignas(64) AttnPersist2Params {
  CUtensorMap tmQ0;   // qkv: box {64, 128, 1}
  CUtensorMap tmQ1;   // qkv: box {64, 32, 1}
  CUtensorMap tmKV;   // qkv: box {64, TK, 1}
  __half* out;
  float* lse;         // [B, heads, T] or nullptr
  const int* lens;    // [B] valid keys per utterance (ragged evaluation batches) or nullptr
  int T, TK, H, heads, nprob, ntiles;
  uint32_t drop_thr;
  float drop_inv_keep;
  unsigned long long drop_seed;
};

constexpr float A2_L2E = 1.4426950408889634f;

// (the chunk helpers of attention_persist.cu, repeated here: both files are self-contained translation units)
template <bool FULL, bool DROP>
__device__ __forceinline__ float a2_exp_chunk(const uint32_t (&v)[8], float mxl, int nv, DropKeys dk, uint32_t pair0,
                                              uint32_t thr_hi, float inv_keep, uint4& out) {
  float sum = 0.f;
  uint32_t pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float e0 = fast_ex2(fmaf(__uint_as_float(v[2 * j]), A2_L2E, -mxl));
    float e1 = fast_ex2(fmaf(__uint_as_float(v[2 * j + 1]), A2_L2E, -mxl));
    if (!FULL) {
      if (2 * j >= nv) e0 = 0.f;
      if (2 * j + 1 >= nv) e1 = 0.f;
    }
    sum += e0 + e1;
    if (DROP) {
      const uint32_t hb = dropout_hash32(dk, pair0 + uint32_t(j));
      e0 = (hb << 16) >= thr_hi ? e0 * inv_keep : 0.f;
      e1 = hb >= thr_hi ? e1 * inv_keep : 0.f;
    }
    pk[j] = pack_half2(e0, e1);
  }
  out = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  return sum;
}
template <bool FULL>
__device__ __forceinline__ float a2_max_chunk(const uint32_t (&v)[8], int nv, float mx) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (FULL || j < nv) mx = fmaxf(mx, __uint_as_float(v[j]));
  return mx;
}

template <bool DROP>
__global__ void __launch_bounds__(A2_THREADS, 2) attention_persist2_kernel(const __grid_constant__ AttnPersist2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int TK = p.TK, T = p.T;
  float* red_max0 = reinterpret_cast<float*>(smem + A2_RED);    // [2][128]
  float* red_max1 = red_max0 + 2 * 128;                         // [8][32]
  float* red_sum0 = red_max1 + 8 * 32;                          // [2][128]
  float* red_sum1 = red_sum0 + 2 * 128;                         // [8][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A2_BARS);
  uint64_t* bar_full = bars;
  uint64_t* bar_a = bars + 1;      // S of tile 1 complete
  uint64_t* bar_b = bars + 2;      // O of tile 1 and S of tile 0 complete (Q / K dead)
  uint64_t* bar_c = bars + 3;      // O of tile 0 complete (every operand buffer dead)
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
    mbar_init(bar_a, 1);
    mbar_init(bar_b, 1);
    mbar_init(bar_c, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
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
  const int nch8 = TK / 8, nchunk16 = TK / 16;
  const int c_begin = (nch8 * cg) >> 1, c_end = (nch8 * (cg + 1)) >> 1;
  const bool two = p.ntiles > 1;
  const DropKeys dkeys = drop_keys(p.drop_seed);
  const uint32_t thr_hi = p.drop_thr << 16;
  const float inv_keep = p.drop_inv_keep;
  const uint32_t idesc_s = make_idesc_f16(128, TK);
  const uint32_t idesc_o = make_idesc_f16(128, A2_D, 0, 1);          // B (= V) MN-major
  const bool do0 = quarter * 32 < T;                                 // warp-uniform (T < 128: the upper quarters idle)

  auto issue_load = [&](int k) {                                     // thread 0; every operand buffer is dead
    const int prob = int(blockIdx.x) + k * int(gridDim.x);
    const int b = prob / p.heads, h = prob - b * p.heads;
    mbar_arrive_expect_tx(bar_full, 16384 + (two ? 4 * 4096 : 0) + 2 * TK * 128);
    tma_load_3d(smem + A2_Q0, &p.tmQ0, bar_full, h * A2_D, 0, b);
    tma_load_3d(smem + A2_K, &p.tmKV, bar_full, p.H + h * A2_D, 0, b);
    tma_load_3d(smem + A2_V, &p.tmKV, bar_full, 2 * p.H + h * A2_D, 0, b);
    if (two) {
#pragma unroll
      for (int r = 0; r < 4; ++r) tma_load_3d(smem + A2_Q1 + r * 4096, &p.tmQ1, bar_full, h * A2_D, 128, b);
    }
  };
  auto issue_s = [&](uint32_t a_q) {                                 // S = Q K^T into the logit columns
#pragma unroll
    for (int kk = 0; kk < A2_D / 16; ++kk)
      umma_f16(tmem + A2_COL_S, make_desc_k_sw128(a_q + kk * 32), make_desc_k_sw128(sbase + A2_K + kk * 32), idesc_s, kk != 0);
  };
  auto issue_o = [&](uint32_t a_p, uint32_t blk, uint32_t col) {     // O = P V
    for (int kk = 0; kk < nchunk16; ++kk) {
      const uint64_t adesc = make_desc_k_sw128(a_p + (kk >> 2) * blk + (kk & 3) * 32);
      const uint64_t bdesc = make_smem_desc(sbase + A2_V + kk * 2048, 16, 1024, 2);       // MN-major, 128B swizzle
      umma_f16(tmem + col, adesc, bdesc, idesc_o, kk != 0);
    }
  };

  if (threadIdx.x == 0 && nloc > 0) issue_load(0);

  for (int k = 0; k < nloc; ++k) {
    const uint32_t ph = k & 1;
    const int prob = int(blockIdx.x) + k * int(gridDim.x);
    const int b = prob / p.heads, h = prob - b * p.heads;
    const uint32_t bh = uint32_t(b) * p.heads + h;
    const int tq1 = 128 + lane;
    const int Tk = p.lens != nullptr ? __ldg(p.lens + b) : T;        // keys that exist for this utterance
    float mx1 = -INFINITY;

    // ---- tile 1 (rows 128 .. T-1; its logits sit in every lane quarter): S1 -> softmax -> P1
    if (two) {
      if (threadIdx.x == 0) {
        mbar_wait(bar_full, ph);
        tc_fence_after();
        issue_s(sbase + A2_Q1);
        umma_commit(bar_a);
      }
      __syncwarp();
      mbar_wait(bar_a, ph);
      __syncwarp();
      tc_fence_after();
      uint32_t v1[3][8];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (warp + 8 * i < nch8) tmem_ld_32x32b_x8(t_row + A2_COL_S + (warp + 8 * i) * 8, v1[i]);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = warp + 8 * i;
        if (c < nch8) {
          const int nv = Tk - c * 8;
          mx1 = nv >= 8 ? a2_max_chunk<true>(v1[i], 8, mx1) : a2_max_chunk<false>(v1[i], nv, mx1);
        }
      }
      red_max1[warp * 32 + lane] = mx1;
      __syncthreads();
#pragma unroll
      for (int w = 0; w < 8; ++w) mx1 = fmaxf(mx1, red_max1[w * 32 + lane]);
      const float mxl = mx1 * A2_L2E;
      const uint32_t pair_row = (bh * T + (tq1 < T ? tq1 : 0)) * uint32_t(TK / 2);
      uint8_t* prow = smem + A2_P1 + lane * 128;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = warp + 8 * i;
        if (c < nch8) {
          const int nv = Tk - c * 8;
          uint4 o;
          sum += nv >= 8 ? a2_exp_chunk<true, DROP>(v1[i], mxl, 8, dkeys, pair_row + c * 4, thr_hi, inv_keep, o)
                         : a2_exp_chunk<false, DROP>(v1[i], mxl, nv, dkeys, pair_row + c * 4, thr_hi, inv_keep, o);
          *reinterpret_cast<uint4*>(prow + (c >> 3) * 4096 + (((c & 7) ^ (lane & 7)) << 4)) = o;
        }
      }
      red_sum1[warp * 32 + lane] = sum;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();                       // P1 complete, every warp is done with the tile-1 logits
    }

    // ---- O1 = P1 V (if any) and S0 = Q0 K^T
    if (threadIdx.x == 0) {
      if (!two) mbar_wait(bar_full, ph);
      tc_fence_after();
      if (two) issue_o(sbase + A2_P1, 4096, A2_COL_O1);
      issue_s(sbase + A2_Q0);
      umma_commit(bar_b);
    }
    __syncwarp();
    mbar_wait(bar_b, ph);
    __syncwarp();
    tc_fence_after();

    // ---- tile 0: logits -> registers, max, exp, P0 over the (now dead) Q / K operands
    float mx0 = -INFINITY;
    if (do0) {
      uint32_t v0[A2_MAXCH][8];
#pragma unroll
      for (int i = 0; i < A2_MAXCH; ++i)
        if (c_begin + i < c_end) tmem_ld_32x32b_x8(t_row + A2_COL_S + (c_begin + i) * 8, v0[i]);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < A2_MAXCH; ++i) {
        const int c = c_begin + i;
        if (c < c_end) {
          const int nv = Tk - c * 8;
          mx0 = nv >= 8 ? a2_max_chunk<true>(v0[i], 8, mx0) : a2_max_chunk<false>(v0[i], nv, mx0);
        }
      }
      red_max0[cg * 128 + row] = mx0;
      named_bar_sync(1 + quarter, 64);       // the two warps (column groups) of this lane quarter
      mx0 = fmaxf(red_max0[row], red_max0[128 + row]);
      const float mxl = mx0 * A2_L2E;
      const uint32_t pair_row = (bh * T + (row < T ? row : 0)) * uint32_t(TK / 2);
      uint8_t* prow = smem + A2_P0 + row * 128;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < A2_MAXCH; ++i) {
        const int c = c_begin + i;
        if (c < c_end) {
          const int nv = Tk - c * 8;
          uint4 o;
          sum += nv >= 8 ? a2_exp_chunk<true, DROP>(v0[i], mxl, 8, dkeys, pair_row + c * 4, thr_hi, inv_keep, o)
                         : a2_exp_chunk<false, DROP>(v0[i], mxl, nv, dkeys, pair_row + c * 4, thr_hi, inv_keep, o);
          *reinterpret_cast<uint4*>(prow + (c >> 3) * 16384 + (((c & 7) ^ (row & 7)) << 4)) = o;
        }
      }
      red_sum0[cg * 128 + row] = sum;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();                         // P0 complete, every warp is done with the tile-0 logits

    // ---- O0 = P0 V into the first logit columns
    if (threadIdx.x == 0) {
      tc_fence_after();
      issue_o(sbase + A2_P0, 16384, A2_COL_O0);
      umma_commit(bar_c);
    }
    __syncwarp();
    float sum0 = 1.f, sum1 = 1.f;
    if (do0) sum0 = red_sum0[row] + red_sum0[128 + row];
    const bool do1 = two && quarter == 0;    // O of tile 1: its rows are lanes 0..31
    if (do1) {
      sum1 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum1 += red_sum1[w * 32 + lane];
    }
    mbar_wait(bar_c, ph);
    __syncwarp();
    tc_fence_after();
    // every operand buffer is dead: the next problem's loads land under this epilogue
    if (threadIdx.x == 0 && k + 1 < nloc) issue_load(k + 1);

#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      if (tile == 0 ? !do0 : !do1) continue;
      const int t_q = tile == 0 ? row : tq1;
      const float sum = tile == 0 ? sum0 : sum1, mx = tile == 0 ? mx0 : mx1;
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + (tile == 0 ? A2_COL_O0 : A2_COL_O1) + cg * 32, r);
      tmem_ld_wait();
      if (t_q < T) {
        const float inv = 1.0f / sum;
        if (cg == 0 && p.lse != nullptr) p.lse[int64_t(bh) * T + t_q] = mx + __logf(sum);
        __half* dst = p.out + (int64_t(b) * T + t_q) * p.H + h * A2_D + cg * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 o;
          o.x = pack_half2(__uint_as_float(r[8 * c]) * inv, __uint_as_float(r[8 * c + 1]) * inv);
          o.y = pack_half2(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv);
          o.z = pack_half2(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv);
          o.w = pack_half2(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 8 * c) = o;
        }
      }
    }
    // the next problem's S / O products overwrite the columns just read; its softmax reuses the exchange buffers
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// -> 0 launched, 1 not applicable / switched off (caller falls back), < 0 error
int attention_persist2_launch(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, uint32_t drop_thr,
                              float drop_inv_keep, uint64_t drop_seed, const int* lens, cudaStream_t stream) {
  // measured (profiles/r02_attention_2cta.txt): a tie with the one-CTA kernel -- off unless W2V2_ATTN_2CTA=1
  static const bool on = []() { const char* e = getenv("W2V2_ATTN_2CTA"); return e != nullptr && e[0] == '1'; }();
  const int TK = (T + 15) / 16 * 16;
  if (!on || TK > A2_MAX_TK) return 1;
  AttnPersist2Params p;
  const uint64_t row_b = uint64_t(3 * H) * 2, utt_b = uint64_t(T) * 3 * H * 2;
  int rc = make_tmap_3d(&p.tmQ0, qkv16, 2, 3 * H, T, B, row_b, utt_b, A2_D, 128, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmQ1, qkv16, 2, 3 * H, T, B, row_b, utt_b, A2_D, 32, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmKV, qkv16, 2, 3 * H, T, B, row_b, utt_b, A2_D, TK, 1, 128);
  if (rc) return rc;
  p.out = static_cast<__half*>(out16);
  p.lse = lse;
  p.lens = lens;
  p.T = T; p.TK = TK; p.H = H; p.heads = heads;
  p.nprob = B * heads;
  p.ntiles = T > 128 ? 2 : 1;
  p.drop_thr = drop_thr;
  p.drop_inv_keep = drop_inv_keep;
  p.drop_seed = drop_seed;
  static bool configured = false;
  if (!configured) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(attention_persist2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM));
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(attention_persist2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM));
    configured = true;
  }
  const int slots = 2 * device_sm_count();
  const int grid = p.nprob < slots ? p.nprob : slots;
  if (drop_thr != 0)
    W2V2_CHECK_CUDA(launch_k(attention_persist2_kernel<true>, dim3(grid), dim3(A2_THREADS), size_t(A2_SMEM), stream, 1, p));
  else
    W2V2_CHECK_CUDA(launch_k(attention_persist2_kernel<false>, dim3(grid), dim3(A2_THREADS), size_t(A2_SMEM), stream, 1, p));
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace w2v2
