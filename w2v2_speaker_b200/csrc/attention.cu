// Self-attention core on tcgen05 (HF:438-463, eager form; no mask -- the reference never passes one):
//     O[b,t,h,:] = softmax_j( q[b,t,h,:] . k[b,j,h,:] ) v[b,j,h,:]          (1/sqrt(d) folded into Wq)
//
// T <= 256 frames (3 s -> 149, 5 s -> 249): the whole key axis fits one UMMA N, so there is no
// online-softmax loop.  One CTA per (batch, head, 128-query tile):
//   TMA   : Q tile [128 x 64], K [TK x 64], V [TK x 64] straight out of the fused qkv activation
//           (3-D tensor map (column, t, b): rows t >= T are zero-filled, TK = round_up(T, 16))
//   MMA 1 : S = Q K^T   (M=128, N=TK, K=64; both operands K-major, 128B swizzle) -> TMEM fp32
//   softmax: one thread per query row reads its row from TMEM (tcgen05.ld 32x32b), max / exp / sum in
//           registers -- no shuffles; P (fp16, unnormalised) goes to smem in the K-major swizzled layout
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
  int T, TK, H, heads;
  int tmem_cols, o_col;
  uint32_t drop_thr;  // attention dropout (HF:456): keep <=> 16 random bits >= thr; 0 = off
  float drop_inv_keep;
  unsigned long long drop_seed;
};

constexpr int ATT_D = 64;

__global__ void __launch_bounds__(128) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int TK = p.TK;
  const int kv_bytes = TK * 128;
  const int pblocks = (TK + 63) / 64;
  uint8_t* sQ = smem;                          // 16 KB
  uint8_t* sK = sQ + 16384;
  uint8_t* sV = sK + ((kv_bytes + 1023) & ~1023);
  uint8_t* sP = sV + ((kv_bytes + 1023) & ~1023);   // pblocks x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + pblocks * 16384);
  uint64_t* bar_tma = bars;
  uint64_t* bar_s = bars + 1;
  uint64_t* bar_o = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int mt = blockIdx.x;                   // query tile
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

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

  // ---- softmax: thread <-> query row
  const int row = warp * 32 + lane;
  const uint32_t t_row = tmem + (uint32_t(warp * 32) << 16);
  mbar_wait(bar_s, 0);
  __syncwarp();
  tc_fence_after();
  const int nchunk = TK / 16;
  float mx = -INFINITY;
  for (int c = 0; c < nchunk; ++c) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(t_row + c * 16, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c * 16 + j < p.T) mx = fmaxf(mx, __uint_as_float(r[j]));
  }
  const float mxl = mx * 1.4426950408889634f;
  float sum = 0.f;
  uint8_t* prow = sP + row * 128;
  for (int c = 0; c < nchunk; ++c) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(t_row + c * 16, r);
    tmem_ld_wait();
    float e[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = exp2f(fmaf(__uint_as_float(r[j]), 1.4426950408889634f, -mxl));
      e[j] = (c * 16 + j < p.T) ? v : 0.f;
    }
    // the value the tensor core will see is the fp16-rounded one: sum those for a consistent normaliser
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __half2 hh = __floats2half2_rn(e[2 * j], e[2 * j + 1]);
      const float2 f = __half22float2(hh);
      sum += f.x + f.y;                           // the normaliser is that of the un-dropped softmax
      if (p.drop_thr != 0) {
        const int t_q = mt * 128 + row;
        const uint64_t pair = ((uint64_t(b) * p.heads + h) * p.T + t_q) * uint64_t(TK / 2) + (c * 8 + j);
        const uint32_t hb = dropout_hash(p.drop_seed, pair);
        hh = __floats2half2_rn((hb & 0xffffu) >= p.drop_thr ? f.x * p.drop_inv_keep : 0.f,
                               (hb >> 16) >= p.drop_thr ? f.y * p.drop_inv_keep : 0.f);
      }
      pk[j] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    const int col = c * 16;                     // 16 halfs = two 16-byte chunks
    uint8_t* blk = prow + (col >> 6) * 16384;
    const int c16 = (col & 63) >> 3;
    *reinterpret_cast<uint4*>(blk + ((c16 ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(blk + (((c16 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
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

  mbar_wait(bar_o, 0);
  __syncwarp();
  tc_fence_after();
  const int t = mt * 128 + row;
  const float inv = 1.0f / sum;
  if (p.lse != nullptr && t < p.T) p.lse[(int64_t(b) * p.heads + h) * p.T + t] = mx + __logf(sum);
  __half* dst = p.out + (int64_t(b) * p.T + (t < p.T ? t : 0)) * p.H + h * ATT_D;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_row + p.o_col + hh * 32, r);
    tmem_ld_wait();
    if (t < p.T) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 q;
        q.x = pack_half2(__uint_as_float(r[8 * c]) * inv, __uint_as_float(r[8 * c + 1]) * inv);
        q.y = pack_half2(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv);
        q.z = pack_half2(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv);
        q.w = pack_half2(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + hh * 32 + 8 * c) = q;
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

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_attention_ex(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, float drop_p,
                                 uint64_t drop_seed, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  W2V2_REQUIRE(heads > 0 && H == heads * ATT_D, "w2v2_attention: head dim must be 64 (H=%d heads=%d)", H, heads);
  W2V2_REQUIRE(T >= 1 && T <= 256,
               "w2v2_attention: T=%d frames not supported (single-tile kernel handles T <= 256; longer "
               "utterances need the tiled variant)", T);
  W2V2_REQUIRE(B >= 1 && B <= 65535, "w2v2_attention: bad batch %d", B);
  AttnParams p;
  const int TK = (T + 15) / 16 * 16;
  int rc = make_tmap_3d(&p.tmQ, qkv16, 2, 3 * H, T, B, uint64_t(3 * H) * 2, uint64_t(T) * 3 * H * 2, ATT_D, 128, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmKV, qkv16, 2, 3 * H, T, B, uint64_t(3 * H) * 2, uint64_t(T) * 3 * H * 2, ATT_D, TK, 1, 128);
  if (rc) return rc;
  p.out = static_cast<__half*>(out16);
  p.lse = lse;
  W2V2_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "w2v2_attention: dropout p=%f out of [0,1)", drop_p);
  p.drop_thr = uint32_t(drop_p * 65536.0f + 0.5f);
  p.drop_inv_keep = 1.0f / (1.0f - float(p.drop_thr) / 65536.0f);
  p.drop_seed = drop_seed;
  p.T = T; p.TK = TK; p.H = H; p.heads = heads;
  if (TK <= 192) { p.tmem_cols = 256; p.o_col = 192; } else { p.tmem_cols = 512; p.o_col = 256; }
  if (TK <= 64) { p.tmem_cols = 128; p.o_col = 64; }
  const int kvb = (TK * 128 + 1023) & ~1023;
  const int smem = 16384 + 2 * kvb + ((TK + 63) / 64) * 16384 + 64 + 1024;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  dim3 grid((T + 127) / 128, heads, B);
  attention_kernel<<<grid, 128, smem, stream>>>(p);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_attention(const void* qkv16, void* out16, float* lse, int B, int T, int H, int heads, void* stream) {
  return w2v2_attention_ex(qkv16, out16, lse, B, T, H, heads, 0.f, 0, stream);
}
