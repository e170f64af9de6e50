// Self-attention forward for sequences of any length (HF:438-463; full-utterance evaluation, where T reaches
// thousands of frames: R:src/lightning_modules/speaker/speaker_recognition_module.py:462-500).
//
// One CTA per (batch, head, 128-query tile) walks the keys in blocks of 128 -- twice:
//   pass 1: S_blk = Q K_blk^T (tcgen05, N = 128) -> per-row running (max, sum of exp) in registers;
//   pass 2: S_blk again, P_blk = exp(S_blk - max) (fp16, K-major swizzled smem), O += P_blk V_blk.
// Two passes instead of an online rescale of the O accumulator: the extra Q K^T is a few percent of the work
// and O never has to leave TMEM; this is the evaluation path, not the training hot loop (T <= 256 there uses
// the single-tile kernels of attention.cu; training on longer sequences -- the paired-input model at two 3 s crops has
// 301 frames -- comes through here with attention dropout applied to P in pass 2).  K / V blocks are double-buffered: the TMA load of block j+1 is
// in flight while block j is processed.  8 warps: warp & 3 = TMEM lane quarter, warp >> 2 = column half; the
// two threads of a row combine their (max, sum) through shared memory once, after pass 1.
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 int swizzle_bytes);

constexpr int AL_D = 64;
constexpr int AL_THREADS = 256;
constexpr int AL_BK = 128;                       // keys per block
constexpr int AL_COL_S = 0, AL_COL_O = 128;      // TMEM columns (256 allocated)
constexpr int AL_Q = 0;                          // smem map (bytes)
constexpr int AL_K = AL_Q + 16384;               // 2 x 16 KB
constexpr int AL_V = AL_K + 2 * 16384;           // 2 x 16 KB
constexpr int AL_P = AL_V + 2 * 16384;           // 2 x 16 KB (two 64-key blocks)
constexpr int AL_RED = AL_P;                     // float [2 stats][2 halves][128 rows]: borrows P between the passes
constexpr int AL_BARS = AL_P + 2 * 16384;
constexpr int AL_SMEM = AL_BARS + 64;            // 112.1 KB: two CTAs per SM

struct alignas(64) AttnLongParams {
  CUtensorMap tm;      // qkv [B, T, 3H]: box {64, 128, 1}
  __half* out;         // [B*T, H]
  float* lse;          // [B, heads, T] or nullptr
  const int* lens;     // [B] valid keys per utterance (ragged evaluation batches) or nullptr
  int T, H, heads, kblocks;
  uint32_t drop_thr;   // attention dropout (training on sequences beyond 256 frames): same mask convention as attention.cu
  float drop_inv_keep;
  unsigned long long drop_seed;
};

__global__ void __launch_bounds__(AL_THREADS) attention_long_kernel(const __grid_constant__ AttnLongParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  float* red = reinterpret_cast<float*>(smem + AL_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AL_BARS);
  uint64_t* bar_q = bars;
  uint64_t* bar_kv = bars + 1;      // [2]
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int mt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, cg = warp >> 2;
  const int row = quarter * 32 + lane;
  const int t_q = mt * 128 + row;
  const bool warp_valid = mt * 128 + quarter * 32 < p.T;      // warp-uniform
  const int Tk = p.lens != nullptr ? __ldg(p.lens + b) : p.T;  // keys that exist for this utterance (>= 1)
  pdl_trigger();

  if (threadIdx.x == 0) {
    prefetch_tensormap(&p.tm);
    mbar_init(bar_q, 1);
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
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

  const uint32_t sbase = smem_u32(smem);
  const uint32_t idesc_s = make_idesc_f16(128, AL_BK);
  const uint32_t idesc_o = make_idesc_f16(128, AL_D, 0, 1);      // B (= V) MN-major
  const int nb = p.kblocks;
  // load sequence: pass 1 needs K only (nb loads), pass 2 needs K and V (nb loads); load number `i` uses buffer i & 1
  auto issue_load = [&](int i) {
    const int kb = i < nb ? i : i - nb;
    const bool with_v = i >= nb;
    uint64_t* bar = &bar_kv[i & 1];
    mbar_arrive_expect_tx(bar, with_v ? 2 * 16384 : 16384);
    tma_load_3d(smem + AL_K + (i & 1) * 16384, &p.tm, bar, p.H + h * AL_D, kb * AL_BK, b);
    if (with_v) tma_load_3d(smem + AL_V + (i & 1) * 16384, &p.tm, bar, 2 * p.H + h * AL_D, kb * AL_BK, b);
  };
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_q, 16384);
    tma_load_3d(smem + AL_Q, &p.tm, bar_q, h * AL_D, mt * 128, b);
    issue_load(0);
  }
  mbar_wait(bar_q, 0);

  uint32_t s_phase = 0, o_phase = 0;
  float m_run = -INFINITY, l_run = 0.f;          // this thread's column half: running max (raw scores) and sum
  float m_row = 0.f, inv_l = 0.f;                // combined, available after pass 1
  const int c_lo = cg * 4, c_hi = c_lo + 4;      // 16-column chunks of this thread within a 128-key block
  const DropKeys dkeys = drop_keys(p.drop_seed);
  const uint32_t drop_thr = p.drop_thr;
  const float inv_keep = p.drop_inv_keep;
  // dropout mask index of (row, key pair): the numbering of attention.cu, TK = T rounded up to 16
  const uint32_t pair_row = ((uint32_t(b) * p.heads + h) * p.T + (t_q < p.T ? t_q : 0)) * uint32_t(((p.T + 15) / 16 * 16) / 2);

  for (int i = 0; i < 2 * nb; ++i) {
    const int kb = i < nb ? i : i - nb;
    const bool pass2 = i >= nb;
    if (i == nb) {
      // ---- between the passes: combine the two column halves of every row
      red[(0 * 2 + cg) * 128 + row] = m_run;
      red[(1 * 2 + cg) * 128 + row] = l_run;
      __syncthreads();
      const float m0 = red[row], m1 = red[128 + row];
      const float l0 = red[256 + row], l1 = red[384 + row];
      m_row = fmaxf(m0, m1);
      const float l = l0 * fast_ex2((m0 - m_row) * 1.4426950408889634f) + l1 * fast_ex2((m1 - m_row) * 1.4426950408889634f);
      inv_l = 1.0f / l;
      __syncthreads();                         // the exchange area is about to become the P buffer again
      if (cg == 0 && p.lse != nullptr && t_q < p.T) p.lse[(int64_t(b) * p.heads + h) * p.T + t_q] = m_row + __logf(l);
    }
    // prefetch the next block into the other buffer (its previous user, block i-1, is fully consumed: the S MMA of
    // i-1 was waited for, and in pass 2 the PV MMA of i-1 as well -- see the o wait at the end of the iteration)
    if (threadIdx.x == 0 && i + 1 < 2 * nb) issue_load(i + 1);
    mbar_wait(&bar_kv[i & 1], (i >> 1) & 1);
    __syncwarp();
    tc_fence_after();
    // ---- S = Q K_blk^T
    if (threadIdx.x == 0) {
      const uint32_t qa = sbase + AL_Q, ka = sbase + AL_K + (i & 1) * 16384;
#pragma unroll
      for (int k = 0; k < AL_D / 16; ++k)
        umma_f16(tmem + AL_COL_S, make_desc_k_sw128(qa + k * 32), make_desc_k_sw128(ka + k * 32), idesc_s, k != 0);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, s_phase);
    s_phase ^= 1;
    __syncwarp();
    tc_fence_after();
    const int key0 = kb * AL_BK;
    if (!pass2) {
      // ---- pass 1: running max / sum over this thread's 64 columns of the block
      if (warp_valid) {
        float sv[64];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(t_row + AL_COL_S + (c_lo + c) * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const bool ok = key0 + (c_lo + c) * 16 + j < Tk;
            sv[c * 16 + j] = ok ? __uint_as_float(r[j]) : -INFINITY;
            mx = fmaxf(mx, sv[c * 16 + j]);
          }
        }
        if (mx > -INFINITY) {              // at least one valid key in this half-block
          const float m_new = fmaxf(m_run, mx);
          const float ml = m_new * 1.4426950408889634f;
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 64; ++j) s += fast_ex2(fmaf(sv[j], 1.4426950408889634f, -ml));     // exp2(-inf) = 0
          l_run = l_run * fast_ex2((m_run - m_new) * 1.4426950408889634f) + s;
          m_run = m_new;
        }
      }
      // every warp is done with S before the next iteration's MMA overwrites it
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    } else {
      // ---- pass 2: P = exp(S - max) -> smem ; O += P V_blk
      if (warp_valid) {
        uint8_t* prow = smem + AL_P + row * 128;
        const float ml = m_row * 1.4426950408889634f;
        for (int c = c_lo; c < c_hi; ++c) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(t_row + AL_COL_S + c * 16, r);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float e0 = fast_ex2(fmaf(__uint_as_float(r[2 * j]), 1.4426950408889634f, -ml));
            float e1 = fast_ex2(fmaf(__uint_as_float(r[2 * j + 1]), 1.4426950408889634f, -ml));
            if (key0 + c * 16 + 2 * j >= Tk) e0 = 0.f;
            if (key0 + c * 16 + 2 * j + 1 >= Tk) e1 = 0.f;
            if (drop_thr != 0) {               // the row sum (pass 1) is that of the undropped probabilities
              const uint32_t hb = dropout_hash32(dkeys, pair_row + uint32_t((key0 + c * 16) / 2 + j));
              e0 = (hb & 0xffffu) >= drop_thr ? e0 * inv_keep : 0.f;
              e1 = (hb >> 16) >= drop_thr ? e1 * inv_keep : 0.f;
            }
            pk[j] = pack_half2(e0, e1);
          }
          const int col = c * 16;
          uint8_t* blk = prow + (col >> 6) * 16384;
          const int c16 = (col & 63) >> 3;
          *reinterpret_cast<uint4*>(blk + ((c16 ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(blk + (((c16 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t pa = sbase + AL_P, va = sbase + AL_V + (i & 1) * 16384;
        // the keys of the last block beyond T contribute P = 0 against zero-filled V rows
        for (int kk = 0; kk < AL_BK / 16; ++kk) {
          const uint64_t adesc = make_desc_k_sw128(pa + (kk >> 2) * 16384 + (kk & 3) * 32);
          const uint64_t bdesc = make_smem_desc(va + kk * 2048, 16, 1024, 2);
          umma_f16(tmem + AL_COL_O, adesc, bdesc, idesc_o, (kb | kk) != 0);
        }
        umma_commit(bar_o);
      }
      __syncwarp();
      // the PV MMA reads P and V: it must retire before the next block overwrites P / reloads this V buffer
      mbar_wait(bar_o, o_phase);
      o_phase ^= 1;
      __syncwarp();
      tc_fence_after();
    }
  }

  // ---- epilogue: O / sum -> fp16
  if (warp_valid) {
    __half* dst = p.out + (int64_t(b) * p.T + (t_q < p.T ? t_q : 0)) * p.H + h * AL_D + cg * 32;
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_row + AL_COL_O + cg * 32, r);
    tmem_ld_wait();
    if (t_q < p.T) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 q;
        q.x = pack_half2(__uint_as_float(r[8 * c]) * inv_l, __uint_as_float(r[8 * c + 1]) * inv_l);
        q.y = pack_half2(__uint_as_float(r[8 * c + 2]) * inv_l, __uint_as_float(r[8 * c + 3]) * inv_l);
        q.z = pack_half2(__uint_as_float(r[8 * c + 4]) * inv_l, __uint_as_float(r[8 * c + 5]) * inv_l);
        q.w = pack_half2(__uint_as_float(r[8 * c + 6]) * inv_l, __uint_as_float(r[8 * c + 7]) * inv_l);
        *reinterpret_cast<uint4*>(dst + 8 * c) = q;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

int attention_long_launch(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, const int* lens,
                          uint32_t drop_thr, float drop_inv_keep, uint64_t drop_seed, cudaStream_t stream) {
  AttnLongParams p;
  int rc = make_tmap_3d(&p.tm, qkv16, 2, 3 * H, T, B, uint64_t(3 * H) * 2, uint64_t(T) * 3 * H * 2, AL_D, 128, 1, 128);
  if (rc) return rc;
  p.out = static_cast<__half*>(out16);
  p.lse = lse;
  p.lens = lens;
  p.T = T; p.H = H; p.heads = heads;
  p.kblocks = (T + AL_BK - 1) / AL_BK;
  p.drop_thr = drop_thr;
  p.drop_inv_keep = drop_inv_keep;
  p.drop_seed = drop_seed;
  static bool configured = false;
  if (!configured) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(attention_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AL_SMEM));
    configured = true;
  }
  dim3 grid((T + 127) / 128, heads, B);
  W2V2_CHECK_CUDA(launch_k(attention_long_kernel, grid, dim3(AL_THREADS), size_t(AL_SMEM), stream, 1, p));
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace w2v2
