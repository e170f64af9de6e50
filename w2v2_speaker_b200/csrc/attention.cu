// Self-attention core on tcgen05 (HF:438-463, eager form; no mask -- the reference never passes one):
//     O[b,t,h,:] = softmax_j( q[b,t,h,:] . k[b,j,h,:] ) v[b,j,h,:]          (1/sqrt(d) folded into Wq)
//
// T <= 256 frames (3 s -> 149, 5 s -> 249): the whole key axis fits one UMMA N, so there is no
// online-softmax loop.  One CTA per (batch, head, 128-query tile):
//   TMA   : Q tile [128 x 64], K [TK x 64], V [TK x 64] straight out of the fused qkv activation
//           (3-D tensor map (column, t, b): rows t >= T are zero-filled, TK = round_up(T, 16))
//   MMA 1 : S = Q K^T   (M=128, N=TK, K=64; both operands K-major, 128B swizzle) -> TMEM fp32
//   softmax: four threads per query row (one per column group) read the row from TMEM (tcgen05.ld 32x32b),
//           max / exp / sum in registers, halves combined through smem; P (fp16, unnormalised) goes to
//           smem in the K-major swizzled layout
//   MMA 2 : O = P V     (M=128, N=64, K=TK; V is consumed as an MN-major B operand, i.e. exactly the
//           [t, d] tile TMA delivered -- no transpose)
//   epilogue: O / rowsum -> fp16 -> out[b*T + t, h*64 : h*64+64]
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 int swizzle_bytes);

struct alignas(64) AttnParams {
  CUtensorMap tmQ;    // box {64, 128, 1}
  CUtensorMap tmKV;   // box {64, TK, 1}
  __half* out;
  float* lse;         // [B, heads, T] or nullptr
  const int* lens;    // [B] valid keys per utterance (ragged evaluation batches) or nullptr
  int T, TK, H, heads;
  int tmem_cols, o_col;
  uint32_t drop_thr;  // attention dropout (HF:456): keep <=> 16 random bits >= thr; 0 = off
  float drop_inv_keep;
  unsigned long long drop_seed;
};

constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 512;      // 16 warps: warp & 3 = TMEM lane quarter (32 query rows), warp >> 2 = column group
constexpr int ATT_NCG = ATT_THREADS / 128;

// Four threads per query row (one per column group): the per-row max / sum meet through shared memory.
// Warps whose 32 query rows all lie beyond T (most of the second tile of a 149-frame utterance) skip the
// arithmetic; their P rows stay unwritten and only feed output rows that are never stored.
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int TK = p.TK;
  const int kv_bytes = TK * 128;
  const int pblocks = (TK + 63) / 64;
  uint8_t* sQ = smem;                          // 16 KB
  uint8_t* sK = sQ + 16384;
  uint8_t* sV = sK + ((kv_bytes + 1023) & ~1023);
  uint8_t* sP = sV + ((kv_bytes + 1023) & ~1023);   // pblocks x 16 KB
  float* red_max = reinterpret_cast<float*>(sP + pblocks * 16384);   // [ATT_NCG][128]
  float* red_sum = red_max + ATT_NCG * 128;                          // [ATT_NCG][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red_sum + ATT_NCG * 128);
  uint64_t* bar_tma = bars;
  uint64_t* bar_s = bars + 1;
  uint64_t* bar_o = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int mt = blockIdx.x;                   // query tile
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, cg = warp >> 2;
  pdl_trigger();
  const int Tk = p.lens != nullptr ? __ldg(p.lens + b) : p.T;       // keys that exist for this utterance

  if (threadIdx.x == 0) {
    prefetch_tensormap(&p.tmQ);
    prefetch_tensormap(&p.tmKV);
    mbar_init(bar_tma, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_tma, 16384 + 2 * kv_bytes);
    tma_load_3d(sQ, &p.tmQ, bar_tma, h * ATT_D, mt * 128, b);
    tma_load_3d(sK, &p.tmKV, bar_tma, p.H + h * ATT_D, 0, b);
    tma_load_3d(sV, &p.tmKV, bar_tma, 2 * p.H + h * ATT_D, 0, b);
    mbar_wait(bar_tma, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, TK);
    const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
    for (int k = 0; k < ATT_D / 16; ++k)
      umma_f16(tmem, make_desc_k_sw128(qa + k * 32), make_desc_k_sw128(ka + k * 32), idesc, k != 0);
    umma_commit(bar_s);
  }
  __syncwarp();

  // ---- softmax: two threads <-> one query row, each owning half of the 16-column chunks
  const int row = quarter * 32 + lane;
  const int t_q = mt * 128 + row;
  const bool warp_valid = mt * 128 + quarter * 32 < p.T;         // warp-uniform
  const uint32_t t_row = tmem + (uint32_t(quarter * 32) << 16);
  const int nchunk = TK / 16;
  const int c_begin = (nchunk * cg) / ATT_NCG, c_end = (nchunk * (cg + 1)) / ATT_NCG;
  const DropKeys dkeys = drop_keys(p.drop_seed);
  const uint32_t drop_thr = p.drop_thr;
  const float inv_keep = p.drop_inv_keep;
  // flat pair index of this row's first pair in the [B, heads, T, TK] mask (fits 32 bits: checked on the host)
  const uint32_t pair_row = ((uint32_t(b) * p.heads + h) * p.T + (t_q < p.T ? t_q : 0)) * uint32_t(TK / 2);

  mbar_wait(bar_s, 0);
  __syncwarp();
  tc_fence_after();
  float mx = -INFINITY;
  if (warp_valid) {
    for (int c = c_begin; c < c_end; ++c) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_row + c * 16, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c * 16 + j < Tk) mx = fmaxf(mx, __uint_as_float(r[j]));
    }
  }
  red_max[cg * 128 + row] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red_max[row], red_max[128 + row]), fmaxf(red_max[256 + row], red_max[384 + row]));
  const float mxl = mx * 1.4426950408889634f;
  float sum = 0.f;
  if (warp_valid) {
    uint8_t* prow = sP + row * 128;
    for (int c = c_begin; c < c_end; ++c) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_row + c * 16, r);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float e0 = fast_ex2(fmaf(__uint_as_float(r[2 * j]), 1.4426950408889634f, -mxl));
        float e1 = fast_ex2(fmaf(__uint_as_float(r[2 * j + 1]), 1.4426950408889634f, -mxl));
        if (c * 16 + 2 * j >= Tk) e0 = 0.f;
        if (c * 16 + 2 * j + 1 >= Tk) e1 = 0.f;
        // the value the tensor core will see is the fp16-rounded one: sum those for a consistent normaliser
        __half2 hh = __floats2half2_rn(e0, e1);
        const float2 f = __half22float2(hh);
        sum += f.x + f.y;                           // the normaliser is that of the un-dropped softmax
        if (drop_thr != 0) {
          const uint32_t hb = dropout_hash32(dkeys, pair_row + uint32_t(c * 8 + j));
          hh = __floats2half2_rn((hb & 0xffffu) >= drop_thr ? f.x * inv_keep : 0.f,
                                 (hb >> 16) >= drop_thr ? f.y * inv_keep : 0.f);
        }
        pk[j] = *reinterpret_cast<const uint32_t*>(&hh);
      }
      const int col = c * 16;                     // 16 halfs = two 16-byte chunks
      uint8_t* blk = prow + (col >> 6) * 16384;
      const int c16 = (col & 63) >> 3;
      *reinterpret_cast<uint4*>(blk + ((c16 ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(blk + (((c16 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
  }
  red_sum[cg * 128 + row] = sum;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, ATT_D, 0, 1);     // B (= V) is MN-major
    const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
    for (int kk = 0; kk < nchunk; ++kk) {
      const uint64_t adesc = make_desc_k_sw128(pa + (kk >> 2) * 16384 + (kk & 3) * 32);
      // MN-major, 128B swizzle: 8 k-rows of 128 B per atom (SBO = 1024), 16 k-rows per MMA
      const uint64_t bdesc = make_smem_desc(va + kk * 2048, 16, 1024, 2);
      umma_f16(tmem + p.o_col, adesc, bdesc, idesc, kk != 0);
    }
    umma_commit(bar_o);
  }
  __syncwarp();
  sum = (red_sum[row] + red_sum[128 + row]) + (red_sum[256 + row] + red_sum[384 + row]);

  mbar_wait(bar_o, 0);
  __syncwarp();
  tc_fence_after();
  if (warp_valid) {
    const float inv = 1.0f / sum;
    if (cg == 0 && p.lse != nullptr && t_q < p.T) p.lse[(int64_t(b) * p.heads + h) * p.T + t_q] = mx + __logf(sum);
    __half* dst = p.out + (int64_t(b) * p.T + (t_q < p.T ? t_q : 0)) * p.H + h * ATT_D + cg * 16;
    uint32_t r[16];
    tmem_ld_32x32b_x16(t_row + p.o_col + cg * 16, r);
    tmem_ld_wait();
    if (t_q < p.T) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint4 q;
        q.x = pack_half2(__uint_as_float(r[8 * c]) * inv, __uint_as_float(r[8 * c + 1]) * inv);
        q.y = pack_half2(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv);
        q.z = pack_half2(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv);
        q.w = pack_half2(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + 8 * c) = q;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

int attention_long_launch(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, const int* lens,
                          uint32_t drop_thr, float drop_inv_keep, uint64_t drop_seed, cudaStream_t stream);
int attention_persist_launch(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, uint32_t drop_thr,
                             float drop_inv_keep, uint64_t drop_seed, const int* lens, cudaStream_t stream);
int attention_persist2_launch(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, uint32_t drop_thr,
                              float drop_inv_keep, uint64_t drop_seed, const int* lens, cudaStream_t stream);

}  // namespace w2v2

using namespace w2v2;

static int attention_dispatch(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, float drop_p,
                              uint64_t drop_seed, const int* lens, void* stream_);

extern "C" int w2v2_attention_ex(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, float drop_p,
                                 uint64_t drop_seed, void* stream_) {
  return attention_dispatch(qkv16, out16, lse, B, T, H, heads, drop_p, drop_seed, nullptr, stream_);
}

// Ragged evaluation batches (utterances zero-padded to a common length): lens[b] = frames of utterance b.  Keys
// >= lens[b] are excluded from the softmax; every query row is still computed (padding rows stay finite).
extern "C" int w2v2_attention_lens(const void* qkv16, void* out16, int B, int T, int H, int heads, const int* lens,
                                   void* stream_) {
  W2V2_REQUIRE(lens != nullptr, "w2v2_attention_lens: lens is required (use w2v2_attention for full batches)");
  return attention_dispatch(qkv16, out16, nullptr, B, T, H, heads, 0.f, 0, lens, stream_);
}

static int attention_dispatch(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, float drop_p,
                              uint64_t drop_seed, const int* lens, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  W2V2_REQUIRE(heads > 0 && H == heads * ATT_D, "w2v2_attention: head dim must be 64 (H=%d heads=%d)", H, heads);
  W2V2_REQUIRE(T >= 1, "w2v2_attention: empty sequence");
  W2V2_REQUIRE(B >= 1 && B <= 65535, "w2v2_attention: bad batch %d", B);
  W2V2_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "w2v2_attention: dropout p=%f out of [0,1)", drop_p);
  if (T > 256) {       // full-utterance evaluation, training on paired crops: key-tiled two-pass kernel (attention_long.cu)
    const uint32_t thr = uint32_t(drop_p * 65536.0f + 0.5f);
    W2V2_REQUIRE(thr == 0 || uint64_t(B) * heads * T * ((T + 15) / 16 * 8) < (1ull << 32),
                 "w2v2_attention: dropout mask index exceeds 32 bits");
    return attention_long_launch(qkv16, out16, lse, B, T, H, heads, lens, thr, 1.0f / (1.0f - float(thr) / 65536.0f),
                                 drop_seed, stream);
  }
  AttnParams p;
  const int TK = (T + 15) / 16 * 16;
  int rc = make_tmap_3d(&p.tmQ, qkv16, 2, 3 * H, T, B, uint64_t(3 * H) * 2, uint64_t(T) * 3 * H * 2, ATT_D, 128, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmKV, qkv16, 2, 3 * H, T, B, uint64_t(3 * H) * 2, uint64_t(T) * 3 * H * 2, ATT_D, TK, 1, 128);
  if (rc) return rc;
  p.out = static_cast<__half*>(out16);
  p.lse = lse;
  p.lens = lens;
  W2V2_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "w2v2_attention: dropout p=%f out of [0,1)", drop_p);
  p.drop_thr = uint32_t(drop_p * 65536.0f + 0.5f);
  p.drop_inv_keep = 1.0f / (1.0f - float(p.drop_thr) / 65536.0f);
  p.drop_seed = drop_seed;
  p.T = T; p.TK = TK; p.H = H; p.heads = heads;
  W2V2_REQUIRE(uint64_t(B) * heads * T * (TK / 2) < (1ull << 32), "w2v2_attention: dropout mask index exceeds 32 bits");
  {     // T <= 160 (training crops): the persistent kernels -- two CTAs per SM (attention_persist2.cu), else one
        // (attention_persist.cu); 1 = not applicable / switched off
    const int prc2 = attention_persist2_launch(qkv16, out16, lse, B, T, H, heads, p.drop_thr, p.drop_inv_keep, drop_seed, lens, stream);
    if (prc2 <= 0) return prc2;
    const int prc = attention_persist_launch(qkv16, out16, lse, B, T, H, heads, p.drop_thr, p.drop_inv_keep, drop_seed, lens, stream);
    if (prc <= 0) return prc;
  }
  if (TK <= 192) { p.tmem_cols = 256; p.o_col = 192; } else { p.tmem_cols = 512; p.o_col = 256; }
  if (TK <= 64) { p.tmem_cols = 128; p.o_col = 64; }
  const int kvb = (TK * 128 + 1023) & ~1023;
  const int smem = 16384 + 2 * kvb + ((TK + 63) / 64) * 16384 + 4096 /*row max / sum exchange*/ + 64 + 1024;
  W2V2_REQUIRE(uint64_t(B) * heads * T * (TK / 2) < (1ull << 32), "w2v2_attention: dropout mask index exceeds 32 bits");
  static int configured_smem = 0;
  if (smem > configured_smem) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  dim3 grid((T + 127) / 128, heads, B);
  W2V2_CHECK_CUDA(launch_k(attention_kernel, grid, dim3(ATT_THREADS), smem, stream, 1, p));
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_attention(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, void* stream) {
  return w2v2_attention_ex(qkv16, out16, lse, B, T, H, heads, 0.f, 0, stream);
}
