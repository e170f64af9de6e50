// Tap-GEMM on tcgen05 tensor cores (sm_100a):
//
//     out[b, r, n] = act( sum_{tap, c} A[b, r*row_stride + tap*tap_stride + c] * W[n, tap*Cin + c] + bias[n] )
//
// One kernel serves every dense contraction of the hot path:
//   * strided Conv1d layers 1..6 of the wav2vec2 feature extractor (HF:254-272) as an im2col-free
//     implicit GEMM over a channels-last [B, L, C] activation: tap j of output row t reads input row
//     (stride*t + j), which is just a TMA tensor map with row pitch stride*C and base offset j*C;
//   * every Linear of the transformer (QKV, out_proj, FFN1, FFN2; HF:500-573), the feature
//     projection (HF:429-434), the ASP 1x1 convs and the CE / AAM classifier GEMMs (ntaps = 1).
//
// Operands are fp16 (RNE-rounded by the producing kernel), accumulation is fp32 in TMEM.
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0    : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
//   warp 1    : MMA issuer    (one elected thread: tcgen05.mma 128xBNx16, commit -> mbarriers)
//   warps 2-9 : epilogue      (two warps per TMEM lane quarter, each owning half of the tile's columns:
//                              double-buffered tcgen05.ld 32x32b.x32 -> bias (from smem) / GELU ->
//                              fp16|fp32 -> warp-private swizzled staging -> warp-issued TMA store,
//                              which also clips the M / N tails; no CTA-wide barrier in steady state)
// The TMEM accumulator is double-buffered (2 x BN columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

constexpr int BM = 128;
constexpr int BK = 64;            // fp16 elements = 128 bytes = one swizzle row
constexpr int A_BYTES = BM * BK * 2;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;
constexpr int WSTAGE_BYTES = 32 * 128;   // per-warp staging: 32 rows x 128 bytes

struct alignas(64) GemmParams {
  CUtensorMap tmA[3];
  CUtensorMap tmB;
  CUtensorMap tmOut;             // box {64 (f16) | 32 (f32) columns, 32 rows, 1}
  const float* bias;             // EPI 1: [N];  EPI 2: per-(batch, column) scale [batch, N]
  const float* shift;            // EPI 2: per-(batch, column) shift [batch, N]
  int ntaps, kblocks_per_tap;
  int m_tiles, n_tiles, batch;
  int N;
  long long rows;                // valid rows per batch element
};

template <int BN>
struct GemmCfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BIAS_BYTES = 2 * BN * 4;        // double-buffered bias tile
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * WSTAGE_BYTES + BIAS_BYTES + 256 /*barriers*/;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit of sm_100");
  static_assert((2 * STAGES + 4) * 8 + 4 <= 256, "barrier area too small");
};

// EPI: 0 = none, 1 = + bias[n], 2 = * scale[b, n] + shift[b, n]  (GroupNorm affine of conv layer 0)
template <int BN, bool OUT_F32, int ACT, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  // the 128B-swizzled tiles need a 1024-byte aligned base; the kernel has no static shared memory, so
  // the dynamic window starts at the CTA's (1024-aligned) shared base -- verified, not assumed
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* wstage = smem + STAGES * Cfg::STAGE_BYTES;
  float* sbias = reinterpret_cast<float*>(wstage + EPI_WARPS * WSTAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sbias) + Cfg::BIAS_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles * p.batch;
  const int k_iters = p.ntaps * p.kblocks_per_tap;
  // Tile schedule.  Default: grid-stride, n fastest (CTAs running concurrently share the A tile in L2).
  // EPI 2 (per-(batch, column) affine): contiguous chunk per CTA, m fastest, so the (batch, n-tile)
  // dependent scale/shift vectors change only once or twice per CTA.
  constexpr bool CHUNKED = (EPI == 2);
  const int tiles_per_cta = (num_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_begin = CHUNKED ? blockIdx.x * tiles_per_cta : blockIdx.x;
  const int tile_end = CHUNKED ? min(num_tiles, tile_begin + tiles_per_cta) : num_tiles;
  const int tile_step = CHUNKED ? 1 : gridDim.x;
  auto decode = [&](int tile, int& nt, int& mt, int& b) {
    if constexpr (CHUNKED) {
      mt = tile % p.m_tiles;
      const int rest = tile / p.m_tiles;
      nt = rest % p.n_tiles;
      b = rest / p.n_tiles;
    } else {
      nt = tile % p.n_tiles;
      const int rest = tile / p.n_tiles;
      mt = rest % p.m_tiles;
      b = rest / p.m_tiles;
    }
  };

  if (warp == 0 && lane == 0) {
    for (int t = 0; t < p.ntaps; ++t) prefetch_tensormap(&p.tmA[t]);
    prefetch_tensormap(&p.tmB);
    prefetch_tensormap(&p.tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        int nt, mt, b;
        decode(tile, nt, mt, b);
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int kb = 0; kb < p.kblocks_per_tap; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + A_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_3d(sa, &p.tmA[tap], &full_bar[stage], kb * BK, mt * BM, b);
            tma_load_3d(sb, &p.tmB, &full_bar[stage], (tap * p.kblocks_per_tap + kb) * BK, nt * BN, 0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_f16(d_tmem, make_desc_k_sw128(a_addr + k * 32), make_desc_k_sw128(b_addr + k * 32), idesc,
                     (it | k) != 0);
          }
          umma_commit(&empty_bar[stage]);       // smem slot free once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);           // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int ew = warp - 2;                    // 0..7
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    const int half = ew >> 2;                   // which half of the tile's columns this warp owns
    constexpr int HALF_COLS = BN / 2;
    constexpr int NCHUNK = HALF_COLS / 32;      // 32-column chunks per warp per tile (4 or 2)
    constexpr int CHUNKS_PER_STORE = OUT_F32 ? 1 : 2;      // 128 bytes of output per staged row
    uint8_t* mystage = wstage + ew * WSTAGE_BYTES;
    uint8_t* crow = mystage + lane * 128;
    const int epi_tid = threadIdx.x - 64;       // 0..255
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t tcount = 0;

    int prev_b = -1, prev_nt = -1;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++tcount) {
      int nt, mt, b;
      decode(tile, nt, mt, b);
      const float* bias_tile = sbias + (EPI == 1 ? (tcount & 1) * BN : 0);
      if constexpr (EPI == 1) {
        // stage this tile's bias slice once (the previous user of this buffer was two tiles ago and
        // every epilogue warp has passed the barrier of the tile in between)
        if (epi_tid < BN) {
          const int n = nt * BN + epi_tid;
          sbias[(tcount & 1) * BN + epi_tid] = (n < p.N) ? __ldg(p.bias + n) : 0.f;
        }
        named_bar_sync(1, EPI_WARPS * 32);
      } else if constexpr (EPI == 2) {
        // scale in sbias[0:BN], shift in sbias[BN:2BN]; single-buffered -> fence both sides.  With the
        // chunked schedule (b, nt) changes at most a couple of times per CTA.
        if (b != prev_b || nt != prev_nt) {
          named_bar_sync(1, EPI_WARPS * 32);
          if (epi_tid < BN) {
            const int n = nt * BN + epi_tid;
            sbias[epi_tid] = (n < p.N) ? __ldg(p.bias + (long long)b * p.N + n) : 0.f;
            sbias[BN + epi_tid] = (n < p.N) ? __ldg(p.shift + (long long)b * p.N + n) : 0.f;
          }
          named_bar_sync(1, EPI_WARPS * 32);
          prev_b = b; prev_nt = nt;
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      __syncwarp();
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN + half * HALF_COLS;
      const int col0 = nt * BN + half * HALF_COLS;      // first output column of this warp
      const int row0 = mt * BM + quarter * 32;          // first output row of this warp

      uint32_t ra[32], rb[32];
      tmem_ld_32x32b_x32(t_addr, ra);

      auto process = [&](uint32_t (&r)[32], int c) {
        // c: chunk index within this warp's half; 32 consecutive columns of one row per thread
        const float* bsrc = bias_tile + half * HALF_COLS + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if constexpr (EPI == 2) {
            const float4 sc = *reinterpret_cast<const float4*>(bsrc + j);
            const float4 sh = *reinterpret_cast<const float4*>(bsrc + BN + j);
            v[j] = fmaf(__uint_as_float(r[j]), sc.x, sh.x);
            v[j + 1] = fmaf(__uint_as_float(r[j + 1]), sc.y, sh.y);
            v[j + 2] = fmaf(__uint_as_float(r[j + 2]), sc.z, sh.z);
            v[j + 3] = fmaf(__uint_as_float(r[j + 3]), sc.w, sh.w);
          } else {
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (EPI == 1) bb = *reinterpret_cast<const float4*>(bsrc + j);
            v[j] = __uint_as_float(r[j]) + bb.x;
            v[j + 1] = __uint_as_float(r[j + 1]) + bb.y;
            v[j + 2] = __uint_as_float(r[j + 2]) + bb.z;
            v[j + 3] = __uint_as_float(r[j + 3]) + bb.w;
          }
        }
        if constexpr (ACT == 1) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_erf2(v[j], v[j + 1]);
        }
        const int sub = c % CHUNKS_PER_STORE;            // position inside the staged 128-byte row
        if (sub == 0) {
          // staging buffer reuse: the previous TMA store of this warp must have finished reading it
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
        if constexpr (OUT_F32) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {                  // 8 chunks of 16 B (4 floats)
            const int pc = q ^ (lane & 7);
            *reinterpret_cast<float4*>(crow + pc * 16) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {                  // 4 chunks of 16 B (8 halfs) per 32 columns
            const int pc = (sub * 4 + q) ^ (lane & 7);
            uint4 w;
            w.x = pack_half2(v[8 * q], v[8 * q + 1]);
            w.y = pack_half2(v[8 * q + 2], v[8 * q + 3]);
            w.z = pack_half2(v[8 * q + 4], v[8 * q + 5]);
            w.w = pack_half2(v[8 * q + 6], v[8 * q + 7]);
            *reinterpret_cast<uint4*>(crow + pc * 16) = w;
          }
        }
        if (sub == CHUNKS_PER_STORE - 1) {
          fence_proxy_async_smem();
          __syncwarp();
          const int scol = col0 + (c - sub) * 32;
          if (lane == 0 && row0 < p.rows && scol < p.N) {   // skip boxes that lie entirely in the M / N tail
            tma_store_3d(&p.tmOut, mystage, scol, row0, b);
            tma_store_commit();
          }
        }
      };

#pragma unroll 1
      for (int c = 0; c < NCHUNK; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32b_x32(t_addr + (c + 1) * 32, rb);
        process(ra, c);
        tmem_ld_wait();
        if (c + 2 < NCHUNK) {
          tmem_ld_32x32b_x32(t_addr + (c + 2) * 32, ra);
        } else {
          // all TMEM reads of this accumulator by this warp are done -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        process(rb, c + 1);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  W2V2_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not found (driver too old?)");
  W2V2_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16B aligned", base);
  W2V2_REQUIRE(stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, "TMA strides (%llu, %llu) not multiples of 16 B",
               (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes);
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  W2V2_REQUIRE(r == CUDA_SUCCESS,
               "cuTensorMapEncodeTiled failed (%d): base=%p dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u,%u)", (int)r,
               base, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
               (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, b0, b1, b2);
  return 0;
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

template <int BN, bool OUT_F32, int ACT, int EPI>
static int launch_gemm(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, OUT_F32, ACT, EPI>;
  if (!configured) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.batch;
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int BN, bool OUT_F32>
static int dispatch_epilogue(const GemmParams& p, int act, cudaStream_t stream) {
  const int epi = p.shift != nullptr ? 2 : (p.bias != nullptr ? 1 : 0);
  if (epi == 2) {
    if constexpr (!OUT_F32 && BN == 256) return launch_gemm<256, false, 1, 2>(p, stream);   // conv0: GN affine + GELU -> f16
    set_last_error("w2v2 gemm: the per-batch affine epilogue is built for f16 output, N > 128");
    return -1;
  }
  if (act == 1) return epi ? launch_gemm<BN, OUT_F32, 1, 1>(p, stream) : launch_gemm<BN, OUT_F32, 1, 0>(p, stream);
  return epi ? launch_gemm<BN, OUT_F32, 0, 1>(p, stream) : launch_gemm<BN, OUT_F32, 0, 0>(p, stream);
}

int gemm_f16_impl(const void* A, int64_t a_rows, int64_t a_row_stride, int64_t a_batch_stride, int batch, int ntaps,
                  int64_t a_tap_stride, int cin, const void* W, int64_t ldw, int N, const float* bias,
                  const float* shift, int act, void* out, int out_dtype, int64_t ldo, int64_t out_batch_stride,
                  cudaStream_t stream) {
  W2V2_REQUIRE(ntaps >= 1 && ntaps <= 3, "w2v2_gemm_f16: ntaps=%d not in [1,3]", ntaps);
  W2V2_REQUIRE(cin % BK == 0, "w2v2_gemm_f16: cin=%d must be a multiple of %d", cin, BK);
  W2V2_REQUIRE(out_dtype == 0 || out_dtype == 1, "w2v2_gemm_f16: out_dtype must be 0 (f16) or 1 (f32)");
  W2V2_REQUIRE(act == 0 || act == 1, "w2v2_gemm_f16: act must be 0 (none) or 1 (gelu)");
  W2V2_REQUIRE(a_rows > 0 && batch > 0 && N > 0, "w2v2_gemm_f16: empty problem");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int BN = (N <= 128) ? 128 : 256;
  const int osz = out_dtype == 1 ? 4 : 2;
  const uint64_t a_bstride = batch > 1 ? uint64_t(a_batch_stride) * 2 : uint64_t(a_rows) * uint64_t(a_row_stride) * 2;
  for (int t = 0; t < ntaps; ++t) {
    const __half* base = static_cast<const __half*>(A) + t * a_tap_stride;
    int rc = make_tmap_3d(&p.tmA[t], base, 2, cin, a_rows, batch, uint64_t(a_row_stride) * 2, a_bstride, BK, BM, 1, 128);
    if (rc) return rc;
  }
  int rc = make_tmap_3d(&p.tmB, W, 2, uint64_t(ntaps) * cin, N, 1, uint64_t(ldw) * 2, uint64_t(N) * ldw * 2, BK, BN, 1, 128);
  if (rc) return rc;
  const uint64_t o_bstride = batch > 1 ? uint64_t(out_batch_stride) * osz : uint64_t(a_rows) * uint64_t(ldo) * osz;
  rc = make_tmap_3d(&p.tmOut, out, osz, N, a_rows, batch, uint64_t(ldo) * osz, o_bstride, out_dtype == 1 ? 32 : 64, 32, 1, 128);
  if (rc) return rc;
  p.bias = bias;
  p.shift = shift;
  p.ntaps = ntaps;
  p.kblocks_per_tap = cin / BK;
  p.m_tiles = int((a_rows + BM - 1) / BM);
  p.n_tiles = (N + BN - 1) / BN;
  p.batch = batch;
  p.N = N;
  p.rows = a_rows;
  if (BN == 256) return out_dtype == 1 ? dispatch_epilogue<256, true>(p, act, stream) : dispatch_epilogue<256, false>(p, act, stream);
  return out_dtype == 1 ? dispatch_epilogue<128, true>(p, act, stream) : dispatch_epilogue<128, false>(p, act, stream);
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_gemm_f16(const void* A, int64_t a_rows, int64_t a_row_stride, int64_t a_batch_stride, int batch,
                             int ntaps, int64_t a_tap_stride, int cin, const void* W, int64_t ldw, int N,
                             const float* bias, int act, void* out, int out_dtype, int64_t ldo,
                             int64_t out_batch_stride, void* stream_) {
  return gemm_f16_impl(A, a_rows, a_row_stride, a_batch_stride, batch, ntaps, a_tap_stride, cin, W, ldw, N, bias, nullptr,
                       act, out, out_dtype, ldo, out_batch_stride, static_cast<cudaStream_t>(stream_));
}
