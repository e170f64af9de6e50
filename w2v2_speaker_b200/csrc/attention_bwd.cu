// Self-attention backward on tcgen05 (backward of HF:438-463, softmax scale folded into q):
//     P  = exp(S - lse),  S = Q K^T          (recomputed; lse saved by the forward)
//     dP = dO V^T,  delta = rowsum(dO * O),  dS = P * (dP - delta)
//     dV = P^T dO,  dQ = dS K,  dK = dS^T Q
// One CTA per (batch, head) walks the (up to two) 128-query tiles; T <= 256 frames.  Every transpose in
// the formulas is free: P / dS live in shared memory as [query rows x 64-key blocks] (128-byte swizzle)
// and are consumed K-major (dQ = dS K) or MN-major (dV = P^T dO, dK = dS^T Q) through the UMMA
// descriptor's major bit; Q, K, V, dO are the [t, d] tiles TMA delivers and serve as K-major or MN-major
// operands as needed.  TMEM (fp32): S/dP [0,TK) | dQ [192,256) (aliases S/dP's first columns when
// TK > 192: it is produced after dP has been consumed) | dV [256,384) | dK [384,512).
// 8 warps: two threads per query row (one per column half) do the exp / dS arithmetic straight from
// tcgen05.ld; warps whose 32 rows lie beyond T skip it, and the MMAs that contract over queries
// (dV, dK) stop at the last 16-query step that holds a valid row, so unwritten rows are never read.
// T <= 192: both Q / dO tiles stay resident; longer sequences reload the 128-row tile per iteration
// (the P / dS buffers then need the space).
// T > 256 (the paired-input model at two 3 s crops: 301 frames): the key axis is cut into blocks of <= 256 keys and
// the kernel is launched once per block (k0 = first key of the block) over ALL query tiles -- with lse and
// delta = rowsum(dO * O) given per row, every (query tile, key block) pair is independent: dK / dV of a key block are
// complete after its launch, dQ is the sum over the blocks (launches after the first add to the fp16 rows in place).
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 int swizzle_bytes);

constexpr int AB_D = 64;
constexpr int AB_THREADS = 256;
constexpr int AB_COL_SP = 0, AB_COL_DV = 256, AB_COL_DK = 384;

struct alignas(64) AttnBwdParams {
  CUtensorMap tmQ;     // qkv: box {64, 128, 1}
  CUtensorMap tmKV;    // qkv: box {64, TK, 1}
  CUtensorMap tmDO;    // dO : box {64, 128, 1}
  const __half* o;     // [B*T, H]
  const __half* d_o;   // [B*T, H]
  const float* lse;    // [B, heads, T]
  __half* dqkv;        // [B*T, 3H]
  int T, TK, H, heads, qtiles;   // TK: keys of this launch's block, rounded up to 16
  int k0;              // first key of the block (multiple of 64)
  int tk_pairs;        // dropout mask row pitch: (T rounded up to 16) / 2
  int dq_accum;        // 1: dQ rows are added to what dqkv holds (key blocks after the first)
  int pblocks;         // 64-key blocks of P / dS
  int resident;        // 1: every Q / dO tile stays in smem; 0: one tile buffer, reloaded per query tile
  int col_dq;          // TMEM column of dQ
  float qscale;        // dq is multiplied by this (chain rule of the d^-0.5 folded into Wq); 1 = leave as is
  float* dbias;        // f32 [3H] (+)= column sums of (dq * qscale | dk | dv), or nullptr
  uint32_t drop_thr;
  float drop_inv_keep;
  unsigned long long drop_seed;
};

__global__ void __launch_bounds__(AB_THREADS, 1) attention_bwd_kernel(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int TK = p.TK;
  const int kv_bytes = ((TK * 128) + 1023) & ~1023;
  const int nq = p.resident ? p.qtiles : 1;
  uint8_t* sP = smem;                               // pblocks x 16 KB (the block after the last = sdS block 0)
  uint8_t* sdS = sP + p.pblocks * 16384;            // pblocks x 16 KB (the block after the last = sQ tile 0)
  uint8_t* sQ = sdS + p.pblocks * 16384;            // nq x 16 KB
  uint8_t* sdO = sQ + nq * 16384;                   // nq x 16 KB
  uint8_t* sK = sdO + nq * 16384;
  uint8_t* sV = sK + kv_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kv_bytes);
  uint64_t* bar_tma = bars;
  uint64_t* bar_mma = bars + 1;
  uint64_t* bar_q = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, cg = warp >> 2;
  const int row = quarter * 32 + lane;
  pdl_trigger();

  if (threadIdx.x == 0) {
    prefetch_tensormap(&p.tmQ);
    prefetch_tensormap(&p.tmKV);
    prefetch_tensormap(&p.tmDO);
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_q, 1);
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
    mbar_arrive_expect_tx(bar_tma, (p.resident ? 2 * p.qtiles * 16384 : 0) + 2 * TK * 128);
    if (p.resident) {
      for (int i = 0; i < p.qtiles; ++i) {
        tma_load_3d(sQ + i * 16384, &p.tmQ, bar_tma, h * AB_D, i * 128, b);
        tma_load_3d(sdO + i * 16384, &p.tmDO, bar_tma, h * AB_D, i * 128, b);
      }
    }
    tma_load_3d(sK, &p.tmKV, bar_tma, p.H + h * AB_D, p.k0, b);
    tma_load_3d(sV, &p.tmKV, bar_tma, 2 * p.H + h * AB_D, p.k0, b);
  }
  mbar_wait(bar_tma, 0);
  __syncwarp();
  tc_fence_after();

  const int nchunk = TK / 16;
  const int c_begin = (nchunk * cg) >> 1, c_end = (nchunk * (cg + 1)) >> 1;
  const int ktiles = (TK + 127) / 128;
  uint32_t mma_phase = 0, q_phase = 0;
  const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS), aK = smem_u32(sK), aV = smem_u32(sV);
  const DropKeys dkeys = drop_keys(p.drop_seed);
  const uint32_t drop_thr = p.drop_thr;
  const float inv_keep = p.drop_inv_keep;

  for (int qt = 0; qt < p.qtiles; ++qt) {
    const int t = qt * 128 + row;
    const bool valid = t < p.T;
    const bool warp_valid = qt * 128 + quarter * 32 < p.T;          // warp-uniform
    const int qb = p.resident ? qt : 0;
    const uint32_t aQ = smem_u32(sQ) + qb * 16384, adO = smem_u32(sdO) + qb * 16384;
    // 16-query steps of this tile that hold at least one valid row (contraction length of dV / dK)
    const int qsteps = min(8, (p.T - qt * 128 + 15) >> 4);
    if (!p.resident) {
      // every MMA that read the tile buffers has completed (bar_mma wait at the end of the previous iteration)
      if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar_q, 2 * 16384);
        tma_load_3d(sQ, &p.tmQ, bar_q, h * AB_D, qt * 128, b);
        tma_load_3d(sdO, &p.tmDO, bar_q, h * AB_D, qt * 128, b);
      }
      mbar_wait(bar_q, q_phase);
      q_phase ^= 1;
      __syncwarp();
      tc_fence_after();
    }
    // ---- (a) S = Q_qt K^T
    if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc_f16(128, TK);
#pragma unroll
      for (int k = 0; k < AB_D / 16; ++k)
        umma_f16(tmem + AB_COL_SP, make_desc_k_sw128(aQ + k * 32), make_desc_k_sw128(aK + k * 32), idesc, k != 0);
      umma_commit(bar_mma);
    }
    __syncwarp();
    // per-row scalars while the MMA runs (both threads of a row compute them)
    float lse = 0.f, delta = 0.f;
    if (valid) {
      lse = p.lse[(int64_t(b) * p.heads + h) * p.T + t];
      const uint4* po = reinterpret_cast<const uint4*>(p.o + (int64_t(b) * p.T + t) * p.H + h * AB_D);
      const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + (int64_t(b) * p.T + t) * p.H + h * AB_D);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 a = po[c], g = pd[c];
        const __half2* ah = reinterpret_cast<const __half2*>(&a);
        const __half2* gh = reinterpret_cast<const __half2*>(&g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = __half22float2(ah[j]), y = __half22float2(gh[j]);
          delta = fmaf(x.x, y.x, delta);
          delta = fmaf(x.y, y.y, delta);
        }
      }
    }
    const uint32_t pair_row = ((uint32_t(b) * p.heads + h) * p.T + (valid ? t : 0)) * uint32_t(p.tk_pairs) + uint32_t(p.k0 >> 1);
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    __syncwarp();
    tc_fence_after();
    // ---- (b) P = exp(S - lse) -> smem (fp16, K-major swizzled blocks); invalid rows / columns -> 0
    uint8_t* prow = sP + row * 128;
    const float lse2 = lse * 1.4426950408889634f;
    if (warp_valid) {
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_row + AB_COL_SP + c * 16, r);
        tmem_ld_wait();
        uint32_t pk[8], pd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float e0 = fast_ex2(fmaf(__uint_as_float(r[2 * j]), 1.4426950408889634f, -lse2));
          float e1 = fast_ex2(fmaf(__uint_as_float(r[2 * j + 1]), 1.4426950408889634f, -lse2));
          if (!valid || p.k0 + c * 16 + 2 * j >= p.T) e0 = 0.f;
          if (!valid || p.k0 + c * 16 + 2 * j + 1 >= p.T) e1 = 0.f;
          pk[j] = pack_half2(e0, e1);
          if (drop_thr != 0) {
            const uint32_t hb = dropout_hash32(dkeys, pair_row + uint32_t(c * 8 + j));
            pd[j] = pack_half2((hb & 0xffffu) >= drop_thr ? e0 * inv_keep : 0.f,
                               (hb >> 16) >= drop_thr ? e1 * inv_keep : 0.f);
          }
        }
        const int col = c * 16;
        uint8_t* blk = prow + (col >> 6) * 16384;
        const int c16 = (col & 63) >> 3;
        *reinterpret_cast<uint4*>(blk + ((c16 ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(blk + (((c16 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (drop_thr != 0) {
          // the dropped probabilities (operand of dV = Pd^T dO) borrow the dS buffer until step (d) fills it
          uint8_t* blk2 = sdS + row * 128 + (col >> 6) * 16384;
          *reinterpret_cast<uint4*>(blk2 + ((c16 ^ (row & 7)) << 4)) = make_uint4(pd[0], pd[1], pd[2], pd[3]);
          *reinterpret_cast<uint4*>(blk2 + (((c16 + 1) ^ (row & 7)) << 4)) = make_uint4(pd[4], pd[5], pd[6], pd[7]);
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- (c) dP = dO_qt V^T (overwrites S) ;  dV += P^T dO_qt
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc_dp = make_idesc_f16(128, TK);
#pragma unroll
      for (int k = 0; k < AB_D / 16; ++k)
        umma_f16(tmem + AB_COL_SP, make_desc_k_sw128(adO + k * 32), make_desc_k_sw128(aV + k * 32), idesc_dp, k != 0);
      const uint32_t idesc_t = make_idesc_f16(128, AB_D, 1, 1);            // A (P^T) and B (dO) MN-major
      for (int kt = 0; kt < ktiles; ++kt) {
        for (int ks = 0; ks < qsteps; ++ks) {                             // 16 queries per step
          const uint64_t adesc = make_smem_desc((drop_thr != 0 ? adS : aP) + kt * 2 * 16384 + ks * 2048, 16384, 1024, 2);
          const uint64_t bdesc = make_smem_desc(adO + ks * 2048, 16, 1024, 2);
          umma_f16(tmem + AB_COL_DV + kt * AB_D, adesc, bdesc, idesc_t, (qt | ks) != 0);
        }
      }
      umma_commit(bar_mma);
    }
    __syncwarp();
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    __syncwarp();
    tc_fence_after();
    // ---- (d) dS = P * (dP - delta) -> smem
    uint8_t* dsrow = sdS + row * 128;
    if (warp_valid) {
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_row + AB_COL_SP + c * 16, r);
        tmem_ld_wait();
        const int col = c * 16;
        const int c16 = (col & 63) >> 3;
        const uint4 p0 = *reinterpret_cast<const uint4*>(prow + (col >> 6) * 16384 + ((c16 ^ (row & 7)) << 4));
        const uint4 p1 = *reinterpret_cast<const uint4*>(prow + (col >> 6) * 16384 + (((c16 + 1) ^ (row & 7)) << 4));
        const __half2* ph0 = reinterpret_cast<const __half2*>(&p0);
        const __half2* ph1 = reinterpret_cast<const __half2*>(&p1);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 pp = __half22float2(j < 4 ? ph0[j] : ph1[j - 4]);
          float d0 = __uint_as_float(r[2 * j]), d1 = __uint_as_float(r[2 * j + 1]);
          if (drop_thr != 0) {                       // dP arrives for the dropped probabilities: mask / keep
            const uint32_t hb = dropout_hash32(dkeys, pair_row + uint32_t(c * 8 + j));
            d0 = (hb & 0xffffu) >= drop_thr ? d0 * inv_keep : 0.f;
            d1 = (hb >> 16) >= drop_thr ? d1 * inv_keep : 0.f;
          }
          // P is exactly 0 on invalid rows / columns: select, so that whatever dP holds there cannot leak
          pk[j] = pack_half2(pp.x != 0.f ? pp.x * (d0 - delta) : 0.f, pp.y != 0.f ? pp.y * (d1 - delta) : 0.f);
        }
        uint8_t* blk = dsrow + (col >> 6) * 16384;
        *reinterpret_cast<uint4*>(blk + ((c16 ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(blk + (((c16 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- (e) dQ_qt = dS K ;  dK += dS^T Q_qt
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc_dq = make_idesc_f16(128, AB_D, 0, 1);           // B (= K) MN-major
      for (int kk = 0; kk < nchunk; ++kk) {
        const uint64_t adesc = make_desc_k_sw128(adS + (kk >> 2) * 16384 + (kk & 3) * 32);
        const uint64_t bdesc = make_smem_desc(aK + kk * 2048, 16, 1024, 2);
        umma_f16(tmem + p.col_dq, adesc, bdesc, idesc_dq, kk != 0);
      }
      const uint32_t idesc_t = make_idesc_f16(128, AB_D, 1, 1);
      for (int kt = 0; kt < ktiles; ++kt) {
        for (int ks = 0; ks < qsteps; ++ks) {
          const uint64_t adesc = make_smem_desc(adS + kt * 2 * 16384 + ks * 2048, 16384, 1024, 2);
          const uint64_t bdesc = make_smem_desc(aQ + ks * 2048, 16, 1024, 2);
          umma_f16(tmem + AB_COL_DK + kt * AB_D, adesc, bdesc, idesc_t, (qt | ks) != 0);
        }
      }
      umma_commit(bar_mma);
    }
    __syncwarp();
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    __syncwarp();
    tc_fence_after();
    // ---- (f) dQ rows -> global (each thread: its half of the 64 columns)
    if (warp_valid) {
      __half* dst = p.dqkv + (int64_t(b) * p.T + (valid ? t : 0)) * 3 * p.H + h * AB_D + cg * 32;
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + p.col_dq + cg * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = __uint_as_float(r[j]) * p.qscale;
        r[j] = __float_as_uint(v[j]);
      }
      if (valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = v[8 * c + j];
          if (p.dq_accum) {                 // the earlier key blocks' share of this row (written by the previous launch)
            const uint4 old = *reinterpret_cast<const uint4*>(dst + 8 * c);
            const __half2* oh = reinterpret_cast<const __half2*>(&old);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __half22float2(oh[j]);
              a[2 * j] += f.x;
              a[2 * j + 1] += f.y;
            }
          }
          uint4 q;
          q.x = pack_half2(a[0], a[1]);
          q.y = pack_half2(a[2], a[3]);
          q.z = pack_half2(a[4], a[5]);
          q.w = pack_half2(a[6], a[7]);
          *reinterpret_cast<uint4*>(dst + 8 * c) = q;
        }
      }
      // the bias gradient is linear in the blocks: each launch adds the column sums of its own share
      if (p.dbias != nullptr) warp_colsum_atomic<32>(v, valid, 1.0f, p.dbias + h * AB_D + cg * 32);
    }
    // all threads must be done with S/dP and dQ columns before the next tile's MMAs overwrite them
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  // ---- dK (column half 0 warps), dV (column half 1 warps) rows (keys) -> global
  for (int kt = 0; kt < ktiles; ++kt) {
    const int key = p.k0 + kt * 128 + row;
    const bool kvalid = key < p.T && kt * 128 + row < TK;
    if (p.k0 + kt * 128 + quarter * 32 >= p.T || kt * 128 + quarter * 32 >= TK) continue;      // warp-uniform
    const uint32_t col = (cg == 0 ? AB_COL_DK : AB_COL_DV) + kt * AB_D;
    __half* dst = p.dqkv + (int64_t(b) * p.T + (kvalid ? key : 0)) * 3 * p.H + h * AB_D + (cg == 0 ? p.H : 2 * p.H);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + col + hh * 32, r);
      tmem_ld_wait();
      if (kvalid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 q;
          q.x = pack_half2(__uint_as_float(r[8 * c]), __uint_as_float(r[8 * c + 1]));
          q.y = pack_half2(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3]));
          q.z = pack_half2(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5]));
          q.w = pack_half2(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7]));
          *reinterpret_cast<uint4*>(dst + hh * 32 + 8 * c) = q;
        }
      }
      if (p.dbias != nullptr) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        warp_colsum_atomic<32>(v, kvalid, 1.0f, p.dbias + (cg == 0 ? p.H : 2 * p.H) + h * AB_D + hh * 32);
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

int attention_bwd_fused_launch(const void* qkv16, const void* o16, const void* do16, const float* lse, void* dqkv16, int B,
                               int T, int H, int heads, uint32_t drop_thr, float drop_inv_keep, uint64_t drop_seed,
                               float qscale, float* dbias, cudaStream_t stream);
int attention_bwd_persist_launch(const void* qkv16, const void* o16, const void* do16, const float* lse, void* dqkv16, int B,
                                 int T, int H, int heads, uint32_t drop_thr, float drop_inv_keep, uint64_t drop_seed,
                                 float qscale, float* dbias, cudaStream_t stream);

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_attention_bwd_ex(const void* qkv16, const void* o16, const void* do16, const float* lse, void* dqkv16,
                                     int B, int T, int H, int heads, float drop_p, uint64_t drop_seed, void* stream_) {
  return w2v2_attention_bwd_ex2(qkv16, o16, do16, lse, dqkv16, B, T, H, heads, drop_p, drop_seed, 1.0f, nullptr, stream_);
}

extern "C" int w2v2_attention_bwd_ex2(const void* qkv16, const void* o16, const void* do16, const float* lse, void* dqkv16,
                                      int B, int T, int H, int heads, float drop_p, uint64_t drop_seed, float qscale,
                                      float* dbias, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  W2V2_REQUIRE(heads > 0 && H == heads * AB_D, "w2v2_attention_bwd: head dim must be 64 (H=%d heads=%d)", H, heads);
  W2V2_REQUIRE(T >= 1, "w2v2_attention_bwd: empty sequence");
  W2V2_REQUIRE(B >= 1 && B <= 65535, "w2v2_attention_bwd: bad batch %d", B);
  AttnBwdParams p;
  const int TK = (T + 15) / 16 * 16;
  W2V2_REQUIRE(uint64_t(B) * heads * T * (TK / 2) < (1ull << 32), "w2v2_attention_bwd: dropout mask index exceeds 32 bits");
  W2V2_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "w2v2_attention_bwd: dropout p=%f out of [0,1)", drop_p);
  if (TK <= 160) {     // 3 s utterances: the short-chain kernel (attention_bwd_fused.cu)
    const uint32_t thr = uint32_t(drop_p * 65536.0f + 0.5f);
    // the persistent form (attention_bwd_persist.cu); 1 = switched off (W2V2_ATTN_PERSIST=0)
    const int prc = attention_bwd_persist_launch(qkv16, o16, do16, lse, dqkv16, B, T, H, heads, thr,
                                                 1.0f / (1.0f - float(thr) / 65536.0f), drop_seed, qscale, dbias, stream);
    if (prc <= 0) return prc;
    return attention_bwd_fused_launch(qkv16, o16, do16, lse, dqkv16, B, T, H, heads, thr,
                                      1.0f / (1.0f - float(thr) / 65536.0f), drop_seed, qscale, dbias, stream);
  }
  int rc = make_tmap_3d(&p.tmQ, qkv16, 2, 3 * H, T, B, uint64_t(3 * H) * 2, uint64_t(T) * 3 * H * 2, AB_D, 128, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmDO, do16, 2, H, T, B, uint64_t(H) * 2, uint64_t(T) * H * 2, AB_D, 128, 1, 128);
  if (rc) return rc;
  p.o = static_cast<const __half*>(o16);
  p.d_o = static_cast<const __half*>(do16);
  p.lse = lse;
  p.dqkv = static_cast<__half*>(dqkv16);
  p.T = T; p.H = H; p.heads = heads;
  p.qtiles = (T + 127) / 128;
  p.tk_pairs = TK / 2;
  p.qscale = qscale;
  p.dbias = dbias;
  p.drop_thr = uint32_t(drop_p * 65536.0f + 0.5f);
  p.drop_inv_keep = 1.0f / (1.0f - float(p.drop_thr) / 65536.0f);
  p.drop_seed = drop_seed;
  // key blocks: one launch when the whole key axis fits (T <= 256), else equal blocks rounded up to 64 keys
  const int nblk = (TK + 255) / 256;
  const int kblk = nblk == 1 ? TK : ((TK + nblk - 1) / nblk + 63) / 64 * 64;
  for (int k0 = 0; k0 < TK; k0 += kblk) {
    const int tkb = TK - k0 < kblk ? TK - k0 : kblk;
    rc = make_tmap_3d(&p.tmKV, qkv16, 2, 3 * H, T, B, uint64_t(3 * H) * 2, uint64_t(T) * 3 * H * 2, AB_D, tkb, 1, 128);
    if (rc) return rc;
    p.TK = tkb;
    p.k0 = k0;
    p.dq_accum = k0 != 0;
    p.pblocks = (tkb + 63) / 64;
    p.resident = (nblk == 1 && tkb <= 192) ? 1 : 0;
    p.col_dq = tkb <= 192 ? 192 : 0;
    const int kvb = (tkb * 128 + 1023) & ~1023;
    const int nq = p.resident ? p.qtiles : 1;
    // the MMAs over 128-row operand windows may read (never use) up to one 16 KB block past sP / sdS: keep
    // at least that much behind them inside the allocation
    const int smem = 2 * p.pblocks * 16384 + 2 * nq * 16384 + 2 * kvb + 64;
    W2V2_REQUIRE(smem <= 227 * 1024, "w2v2_attention_bwd: shared memory budget exceeded (%d bytes)", smem);
    static int configured = 0;
    if (smem > configured) {
      W2V2_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured = smem;
    }
    dim3 grid(heads, B);
    W2V2_CHECK_CUDA(launch_k(attention_bwd_kernel, grid, dim3(AB_THREADS), smem, stream, 1, p));
    count_launches(1);
    W2V2_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int w2v2_attention_bwd(const void* qkv16, const void* o16, const void* do16, const float* lse, void* dqkv16,
                                  int B, int T, int H, int heads, void* stream) {
  return w2v2_attention_bwd_ex(qkv16, o16, do16, lse, dqkv16, B, T, H, heads, 0.f, 0, stream);
}
