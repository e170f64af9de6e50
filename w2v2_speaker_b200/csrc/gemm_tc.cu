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
  CUtensorMap tmOut2;            // DUAL: second f16 output, the pre-activation (bias added, before the GELU)
  const float* bias;             // EPI 1: [N];  EPI 2: per-(batch, column) scale [batch, N]
  const float* shift;            // EPI 2: per-(batch, column) shift [batch, N]
  int ntaps, kblocks_per_tap;
  int tap_row[3];                // row-coordinate offset of each tap (may be negative: rows before the first are zero-filled)
  int m_tiles, n_tiles, batch;
  int N;
  long long rows;                // valid rows per batch element
  unsigned int* sk_flags;        // stream-K: per-tile hand-shake counters (nullptr = whole tiles per CTA)
  const __half* aux;             // ACT 2: z [rows, ld_aux] f16, the pre-activation whose gelu' multiplies the output
  long long ld_aux;
  float* colsum;                 // ACT 2: colsum[n] += sum over rows of the (fp16-rounded) output
  int accumulate;                // fp32 output without activation: out += A W^T (every store is a TMA reduce-add)
  int no_tail_skip;              // W2V2_GEMM_TAIL_SKIP=0 (A/B): epilogue warps whose rows are all out of range still run
};

// Stream-K (fp32 output, no activation): the (tile, k-block) iteration space is cut into gridDim.x equal
// contiguous ranges, so a problem of 1.5 waves of tiles costs 1.5 instead of 2 tile times.  A tile cut
// by a range boundary is produced by two CTAs: the one holding its leading k-blocks stores bias + partial
// sums plainly, the other one adds its partial sums with TMA reduce-add stores once the first is done.
// Every CTA walks its range from the top, so the plain store is the first thing the lower-numbered CTA
// does and the reduce-add the last thing the higher-numbered one does: the wait is (almost) never
// taken, and it only ever points at a CTA that was scheduled earlier.
constexpr int SK_RING = 8, SK_MAX_TILES = 4096;
__device__ unsigned int g_sk_flags[SK_RING * SK_MAX_TILES];

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// The per-CTA work list, identical for the three warp roles: (tile, first k-iteration, end k-iteration).
struct WorkIter {
  int tile, tile_end, tile_step;       // whole-tile schedules
  long long cur, lo;                   // stream-K: remaining range [lo, cur) of global k-iterations
  int k_iters;
  bool sk;
  __device__ __forceinline__ bool next(int& t, int& kb, int& ke) {
    if (!sk) {
      if (tile >= tile_end) return false;
      t = tile; kb = 0; ke = k_iters;
      tile += tile_step;
      return true;
    }
    if (cur <= lo) return false;
    t = int((cur - 1) / k_iters);
    const long long t0 = (long long)t * k_iters;
    const long long seg_lo = t0 > lo ? t0 : lo;
    kb = int(seg_lo - t0);
    ke = int(cur - t0);
    cur = seg_lo;
    return true;
  }
};

// CL = CTAs per MMA: 1 (tcgen05.mma.cta_group::1, tile 128 x BN) or 2 (a CTA pair, cta_group::2, tile 256 x BN:
// each CTA stages its 128 rows of A and BN/2 columns of B)
// MINB = CTAs per SM the kernel is built for: 2 only for the epilogue-bound conv0 GEMM (K = 64: one k-block per
// tile, so two pipeline stages are plenty and two CTAs -- 16 epilogue warps -- share an SM)
// DUAL: 1 = the epilogue stores the pre-activation next to the activated output (training keeps both), 2 = it stores
// gelu'(pre-activation) instead (all the backward needs of it; experimental, see schedule.cu): twice the
// staging, one pipeline stage less
constexpr int EPI2_MAX_N = 512;
template <int BN, int CL = 1, int MINB = 1, int DUAL = 0, bool AFFINE = false>
struct GemmCfg {
  static constexpr int B_BYTES = (BN / CL) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (MINB == 2 ? 2 : ((BN == 256 && CL == 1) ? (DUAL ? 3 : 4) : (DUAL ? 5 : 6))) -
                                ((AFFINE && BN == 256) ? 1 : 0);     // (the wide affine staging costs the A/B variants a stage)
  static constexpr int WSTAGE = WSTAGE_BYTES * (DUAL ? 2 : 1);
  static constexpr int TMEM_COLS = 2 * BN;
  // double-buffered bias tile; the per-(batch, column) affine of conv layer 0 (EPI 2) keeps scale and shift of ALL
  // its columns (<= EPI2_MAX_N each) so that the n-tiles of one m-tile can run back to back
  static constexpr int BIAS_BYTES = AFFINE ? 2 * EPI2_MAX_N * 4 : 2 * BN * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * WSTAGE + BIAS_BYTES + 256 /*barriers*/;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit of sm_100");
  static_assert((2 * STAGES + 4) * 8 + 4 <= 256, "barrier area too small");
};

// EPI: 0 = none, 1 = + bias[n], 2 = * scale[b, n] + shift[b, n]  (GroupNorm affine of conv layer 0)
// ACT: 0 = none, 1 = GELU, 2 = multiply by gelu'(aux[r, n]) and accumulate the column sums of the result: the data
//      gradient of FFN2 fused with the GELU backward and the FFN1 bias gradient (f16 output, EPI 0, one batch)
template <int BN, bool OUT_F32, int ACT, int EPI, int CL, int MINB = 1, int DUAL = 0>
__global__ void __launch_bounds__(NUM_THREADS, MINB) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, CL, MINB, DUAL, EPI == 2>;
  static_assert(!DUAL || (!OUT_F32 && ACT == 1), "the dual-output epilogue is the fp16 pre-activation + GELU pair");
  static_assert(ACT < 2 || (!OUT_F32 && EPI == 0 && !DUAL), "the gelu'-multiply epilogue writes f16 and takes no bias");
  constexpr bool GRAD = (ACT == 2 || ACT == 3);      // 3: aux already holds gelu'(z), the epilogue only multiplies
  constexpr int STAGES = Cfg::STAGES;
  // the 128B-swizzled tiles need a 1024-byte aligned base; the kernel has no static shared memory, so
  // the dynamic window starts at the CTA's (1024-aligned) shared base -- verified, not assumed
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_trigger();
  uint8_t* wstage = smem + STAGES * Cfg::STAGE_BYTES;
  float* sbias = reinterpret_cast<float*>(wstage + EPI_WARPS * Cfg::WSTAGE);
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
  constexpr bool SK_OK = OUT_F32 && ACT == 0 && EPI != 2;       // stream-K is only ever requested for these
  // the scheduling unit is the CTA (CL = 1) or the CTA pair (CL = 2): both CTAs of a pair walk the same list
  const int cta_rank = CL == 2 ? int(cluster_ctarank()) : 0;
  const int sched_id = blockIdx.x / CL, num_sched = gridDim.x / CL;
  const int tiles_per_cta = (num_tiles + num_sched - 1) / num_sched;
  const int tile_begin = CHUNKED ? sched_id * tiles_per_cta : sched_id;
  const int tile_end = CHUNKED ? min(num_tiles, tile_begin + tiles_per_cta) : num_tiles;
  const int tile_step = CHUNKED ? 1 : num_sched;
  WorkIter work0;
  work0.tile = tile_begin; work0.tile_end = tile_end; work0.tile_step = tile_step;
  work0.k_iters = k_iters;
  work0.sk = SK_OK && p.sk_flags != nullptr;
  work0.cur = work0.lo = 0;
  if (work0.sk) {
    const long long total = (long long)num_tiles * k_iters;
    const long long per = total / num_sched, rem = total % num_sched;
    work0.lo = per * sched_id + (sched_id < rem ? sched_id : rem);
    work0.cur = work0.lo + per + (sched_id < rem ? 1 : 0);
  }
  auto decode = [&](int tile, int& nt, int& mt, int& b) {
    if constexpr (CHUNKED) {
      // contiguous chunk per CTA, n fastest inside it: the n-tiles of one m-tile run back to back on the same CTA, so the
      // A tile (conv layer 0: the 16 KB window operand) is read from DRAM once instead of once per n-tile
      nt = tile % p.n_tiles;
      const int rest = tile / p.n_tiles;
      mt = rest % p.m_tiles;
      b = rest / p.m_tiles;
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
      mbar_init(&tmem_empty[a], EPI_WARPS * CL);     // pair: the peer's epilogue warps arrive remotely
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CL == 2) {
      tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CL == 2) cluster_sync_all();         // the peer's barriers exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // everything above overlapped the previous kernel's tail; its results are needed from here on

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      WorkIter work = work0;
      int tile, k_begin, k_end;
      while (work.next(tile, k_begin, k_end)) {
        int nt, mt, b;
        decode(tile, nt, mt, b);
        int tap = k_begin / p.kblocks_per_tap, kb = k_begin % p.kblocks_per_tap;
        for (int it = k_begin; it < k_end; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if constexpr (CL == 2) {
            // both CTAs load their halves; the bytes of both are counted on the leader's barrier
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            tma_load_3d_2sm(sa, &p.tmA[tap], lead_bar, kb * BK, (mt * 2 + cta_rank) * BM + p.tap_row[tap], b);
            tma_load_3d_2sm(sb, &p.tmB, lead_bar, (tap * p.kblocks_per_tap + kb) * BK, nt * BN + cta_rank * (BN / 2), 0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_3d(sa, &p.tmA[tap], &full_bar[stage], kb * BK, mt * BM + p.tap_row[tap], b);
            tma_load_3d(sb, &p.tmB, &full_bar[stage], (tap * p.kblocks_per_tap + kb) * BK, nt * BN, 0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++kb == p.kblocks_per_tap) { kb = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && cta_rank == 0) {               // pair: the leader CTA issues for both
      constexpr uint32_t idesc = make_idesc_f16(BM * CL, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      WorkIter work = work0;
      int tile, k_begin, k_end;
      while (work.next(tile, k_begin, k_end)) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < k_end - k_begin; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if constexpr (CL == 2)
              umma_f16_2sm(d_tmem, make_desc_k_sw128(a_addr + k * 32), make_desc_k_sw128(b_addr + k * 32), idesc,
                           (it | k) != 0);
            else
              umma_f16(d_tmem, make_desc_k_sw128(a_addr + k * 32), make_desc_k_sw128(b_addr + k * 32), idesc,
                       (it | k) != 0);
          }
          // smem slot free once these MMAs retire (pair: in both CTAs)
          if constexpr (CL == 2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete
        if constexpr (CL == 2) umma_commit_2sm(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
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
    uint8_t* mystage = wstage + ew * Cfg::WSTAGE;
    uint8_t* crow = mystage + lane * 128;
    const int epi_tid = threadIdx.x - 64;       // 0..255
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t tcount = 0;

    int prev_b = -1, prev_nt = -1;
    if constexpr (GRAD) {       // sbias holds the column-sum accumulators of two tiles in flight
      for (int i = epi_tid; i < 2 * BN; i += EPI_WARPS * 32) sbias[i] = 0.f;
      named_bar_sync(1, EPI_WARPS * 32);
    }
    WorkIter work = work0;
    int tile, k_begin, k_end;
    for (; work.next(tile, k_begin, k_end); ++tcount) {
      int nt, mt, b;
      decode(tile, nt, mt, b);
      const int col0 = nt * BN + half * HALF_COLS;      // first output column of this warp
      const int row0 = (mt * CL + cta_rank) * BM + quarter * 32;      // first output row of this warp
      // GRAD: this thread's 32-column slices of the aux row, fetched one chunk pair ahead of their use
      uint4 za[4], zb[4];
      const bool zrow_ok = GRAD && (row0 + lane) < p.rows;
      const __half* zrow = GRAD ? p.aux + (long long)(row0 + lane) * p.ld_aux + col0 : nullptr;
      auto load_z = [&](uint4 (&zz)[4], int c) {
        if (zrow_ok && col0 + c * 32 < p.N) {
#pragma unroll
          for (int q = 0; q < 4; ++q) zz[q] = __ldg(reinterpret_cast<const uint4*>(zrow + c * 32 + q * 8));
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) zz[q] = make_uint4(0u, 0u, 0u, 0u);
        }
      };
      if constexpr (GRAD) {
        load_z(za, 0);
        load_z(zb, 1);
      }
      // stream-K roles of this segment: `lead` holds the tile's first k-blocks (plain store, adds the bias,
      // then signals), `trail` the rest (waits for the signal, reduce-adds, no bias)
      // accumulate mode (out += A W^T): every segment reduce-adds, sums commute, so no hand-shake is needed at all
      const bool acc_out = SK_OK && p.accumulate != 0;
      const bool sk_lead = SK_OK && !acc_out && k_begin == 0 && k_end < k_iters;
      const bool sk_trail = SK_OK && !acc_out && k_begin != 0;
      const bool use_red = SK_OK && (k_begin != 0 || acc_out);
      const bool add_bias = k_begin == 0;
      const float* bias_tile = sbias + (EPI == 1 ? (tcount & 1) * BN : (EPI == 2 ? nt * BN : 0));
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
        if (b != prev_b) {
          named_bar_sync(1, EPI_WARPS * 32);
          for (int n = epi_tid; n < EPI2_MAX_N; n += EPI_WARPS * 32) {
            sbias[n] = (n < p.N) ? __ldg(p.bias + (long long)b * p.N + n) : 0.f;
            sbias[EPI2_MAX_N + n] = (n < p.N) ? __ldg(p.shift + (long long)b * p.N + n) : 0.f;
          }
          named_bar_sync(1, EPI_WARPS * 32);
          prev_b = b; prev_nt = nt;
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      __syncwarp();
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN + half * HALF_COLS;

      if (SK_OK && sk_trail) {
        // the lead CTA's plain stores of this tile must have landed before anything is added to them
        if (lane == 0) {
          volatile unsigned int* f = p.sk_flags + tile;
          unsigned int spins = 0;
          while (*f < EPI_WARPS * CL) {
            if (++spins > (1u << 28)) __trap();       // a lost hand-shake must fail loudly, not hang
          }
          __threadfence();
        }
        __syncwarp();
      }
      // A warp whose 32 rows lie entirely behind the last valid row (the m tail: 9536 rows leave 64 of the last 256-row
      // pair tile, so six of its eight row groups are empty) has nothing to compute or store: it hands the accumulator
      // back at once.  For the epilogue-bound variants that makes the tail tiles of the last scheduling round cheap.
      const bool rows_live = row0 < p.rows || p.no_tail_skip;      // warp-uniform
      uint32_t ra[32], rb[32];
      if (rows_live) tmem_ld_32x32b_x32(t_addr, ra);

      auto process = [&](uint32_t (&r)[32], int c, const uint4 (&zz)[4]) {
        // c: chunk index within this warp's half; 32 consecutive columns of one row per thread
        const float* bsrc = bias_tile + half * HALF_COLS + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if constexpr (EPI == 2) {
            const float4 sc = *reinterpret_cast<const float4*>(bsrc + j);
            const float4 sh = *reinterpret_cast<const float4*>(bsrc + EPI2_MAX_N + j);
            v[j] = fmaf(__uint_as_float(r[j]), sc.x, sh.x);
            v[j + 1] = fmaf(__uint_as_float(r[j + 1]), sc.y, sh.y);
            v[j + 2] = fmaf(__uint_as_float(r[j + 2]), sc.z, sh.z);
            v[j + 3] = fmaf(__uint_as_float(r[j + 3]), sc.w, sh.w);
          } else {
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (EPI == 1) {
              if (add_bias) bb = *reinterpret_cast<const float4*>(bsrc + j);
            }
            v[j] = __uint_as_float(r[j]) + bb.x;
            v[j + 1] = __uint_as_float(r[j + 1]) + bb.y;
            v[j + 2] = __uint_as_float(r[j + 2]) + bb.z;
            v[j + 3] = __uint_as_float(r[j + 3]) + bb.w;
          }
        }
        const int sub = c % CHUNKS_PER_STORE;            // position inside the staged 128-byte row
        if (sub == 0) {
          // staging buffer reuse: the previous TMA store of this warp must have finished reading it
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
        if constexpr (DUAL == 2) {                       // gelu'(pre-activation) goes to the second staging buffer
          float dv[32];
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_and_grad2(v[j], v[j + 1], dv[j], dv[j + 1]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int pc = (sub * 4 + q) ^ (lane & 7);
            uint4 w;
            w.x = pack_half2(dv[8 * q], dv[8 * q + 1]);
            w.y = pack_half2(dv[8 * q + 2], dv[8 * q + 3]);
            w.z = pack_half2(dv[8 * q + 4], dv[8 * q + 5]);
            w.w = pack_half2(dv[8 * q + 6], dv[8 * q + 7]);
            *reinterpret_cast<uint4*>(crow + WSTAGE_BYTES + pc * 16) = w;
          }
        } else if constexpr (DUAL == 1) {                // the pre-activation goes to the second staging buffer
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int pc = (sub * 4 + q) ^ (lane & 7);
            uint4 w;
            w.x = pack_half2(v[8 * q], v[8 * q + 1]);
            w.y = pack_half2(v[8 * q + 2], v[8 * q + 3]);
            w.z = pack_half2(v[8 * q + 4], v[8 * q + 5]);
            w.w = pack_half2(v[8 * q + 6], v[8 * q + 7]);
            *reinterpret_cast<uint4*>(crow + WSTAGE_BYTES + pc * 16) = w;
          }
        }
        if constexpr (ACT == 1 && DUAL != 2) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_erf2(v[j], v[j + 1]);
        }
        if constexpr (GRAD) {
          const __half2* zh = reinterpret_cast<const __half2*>(&zz[0]);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 zf = __half22float2(zh[j >> 1]);
            if constexpr (ACT == 2) {
              float d0, d1;
              gelu_grad2(zf.x, zf.y, d0, d1);
              v[j] *= d0;
              v[j + 1] *= d1;
            } else {
              v[j] *= zf.x;
              v[j + 1] *= zf.y;
            }
          }
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
            if (SK_OK && use_red) tma_reduce_add_3d(&p.tmOut, mystage, scol, row0, b);
            else tma_store_3d(&p.tmOut, mystage, scol, row0, b);
            if constexpr (DUAL) tma_store_3d(&p.tmOut2, mystage + WSTAGE_BYTES, scol, row0, b);
            tma_store_commit();
          }
          if constexpr (GRAD) {
            // column sums of the 32 x 64 block just staged (what the weight-gradient GEMM will read: the fp16
            // values): lane owns two adjacent columns, walks the rows through the swizzle, conflict-free
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
              const int pc = (lane >> 2) ^ (rr & 7);
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(mystage + rr * 128 + pc * 16 + (lane & 3) * 4));
              s0 += f.x;
              s1 += f.y;
            }
            float* dst = sbias + (tcount & 1) * BN + half * HALF_COLS + (c - sub) * 32 + lane * 2;
            atomicAdd(dst, s0);
            atomicAdd(dst + 1, s1);
          }
        }
      };

      if (!rows_live) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CL == 2 && cta_rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
          else mbar_arrive(&tmem_empty[acc]);
        }
      }
#pragma unroll 1
      for (int c = 0; rows_live && c < NCHUNK; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32b_x32(t_addr + (c + 1) * 32, rb);
        process(ra, c, za);
        if constexpr (GRAD) {
          if (c + 2 < NCHUNK) load_z(za, c + 2);
        }
        tmem_ld_wait();
        if (c + 2 < NCHUNK) {
          tmem_ld_32x32b_x32(t_addr + (c + 2) * 32, ra);
        } else {
          // all TMEM reads of this accumulator by this warp are done -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CL == 2 && cta_rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
            else mbar_arrive(&tmem_empty[acc]);
          }
        }
        process(rb, c + 1, zb);
        if constexpr (GRAD) {
          if (c + 3 < NCHUNK) load_z(zb, c + 3);
        }
      }
      if constexpr (GRAD) {
        // all eight warps have added this tile's partial sums: one global atomic per column, accumulator re-armed
        // (the other buffer serves the next tile, so nobody adds to this one before the next barrier)
        named_bar_sync(1, EPI_WARPS * 32);
        if (epi_tid < BN) {
          float* sacc = sbias + (tcount & 1) * BN + epi_tid;
          const float t = *sacc;
          *sacc = 0.f;
          const int n = nt * BN + epi_tid;
          if (n < p.N) atomicAdd(p.colsum + n, t);
        }
      }
      if (SK_OK && (sk_lead || sk_trail) && lane == 0) {
        if (sk_lead) {
          tma_store_wait<0>();                        // the stores have been performed, not merely read out
          __threadfence();
          atomicAdd(p.sk_flags + tile, 1u);
        } else if (atomicAdd(p.sk_flags + tile, 1u) == 2 * EPI_WARPS * CL - 1) {
          atomicExch(p.sk_flags + tile, 0u);          // last trailing warp: re-arm the counter for the next launch
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  if constexpr (CL == 2) cluster_sync_all();         // nothing of the peer may still point at this CTA
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CL == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("W2V2_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// W2V2_STREAMK=0 disables the stream-K schedule (A/B measurements, tools/time_ops.py)
static bool sk_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("W2V2_STREAMK");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// W2V2_GEMM_PAIR=0 keeps every GEMM on the single-CTA MMA (A/B measurements)
static bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("W2V2_GEMM_PAIR");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// W2V2_GEMM_TAIL_SKIP=0 (A/B measurements): epilogue warps whose rows all lie behind the last valid row still run
static int no_tail_skip() {
  static const int v = []() { const char* e = getenv("W2V2_GEMM_TAIL_SKIP"); return (e != nullptr && e[0] == '0') ? 1 : 0; }();
  return v;
}

template <int BN, bool OUT_F32, int ACT, int EPI, int CL, int MINB = 1, int DUAL = 0>
static int launch_gemm(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CL, MINB, DUAL, EPI == 2>;
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, OUT_F32, ACT, EPI, CL, MINB, DUAL>;
  if (!configured) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.batch;          // scheduling units: CTAs (CL = 1) or CTA pairs
  const int units = device_sm_count() / CL * MINB;
  const int grid = tiles < units ? tiles : units;
  GemmParams q = p;
  if constexpr (OUT_F32 && ACT == 0 && EPI != 2) {
    // stream-K when whole tiles would leave a large part of the last wave idle (e.g. 225 tiles on 148 SMs)
    const int waves = (tiles + grid - 1) / grid;
    if (tiles > grid && tiles <= SK_MAX_TILES && double(tiles) / (double(waves) * grid) < 0.9 && sk_enabled()) {
      static unsigned int* flags = nullptr;
      static unsigned int slot = 0;
      if (flags == nullptr) W2V2_CHECK_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&flags), g_sk_flags));
      q.sk_flags = flags + (slot++ % SK_RING) * SK_MAX_TILES;
    }
  }
  W2V2_CHECK_CUDA(launch_k(kern, dim3(grid * CL), dim3(NUM_THREADS), Cfg::SMEM_BYTES, stream, CL, q));
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int BN, bool OUT_F32, int CL>
static int dispatch_epilogue(const GemmParams& p, int act, cudaStream_t stream) {
  const int epi = p.shift != nullptr ? 2 : (p.bias != nullptr ? 1 : 0);
  if (epi == 2) {
    if constexpr (!OUT_F32 && BN == 256) {      // conv0: GN affine (+ GELU) -> f16
      return act == 1 ? launch_gemm<256, false, 1, 2, CL>(p, stream) : launch_gemm<256, false, 0, 2, CL>(p, stream);
    }
    if constexpr (!OUT_F32 && BN == 128 && CL == 1) {   // same, 128-wide tiles, two CTAs per SM (W2V2_CONV0_2CTA)
      return act == 1 ? launch_gemm<128, false, 1, 2, 1, 2>(p, stream) : launch_gemm<128, false, 0, 2, 1, 2>(p, stream);
    }
    set_last_error("w2v2 gemm: the per-batch affine epilogue is built for f16 output");
    return -1;
  }
  if (act == 1) return epi ? launch_gemm<BN, OUT_F32, 1, 1, CL>(p, stream) : launch_gemm<BN, OUT_F32, 1, 0, CL>(p, stream);
  return epi ? launch_gemm<BN, OUT_F32, 0, 1, CL>(p, stream) : launch_gemm<BN, OUT_F32, 0, 0, CL>(p, stream);
}

// ---- optional per-launch timing of the tensor-core GEMMs (bench.py's roofline figure) ---------------------
// Between w2v2_gemm_profile_start() and w2v2_gemm_profile_stop() every tap-GEMM / wgrad launch is bracketed
// by CUDA events on its own stream; stop() synchronises and returns the summed time and FLOPs.
struct ProfRec {
  cudaEvent_t s, e;
  double flops;
};
static bool g_prof_on = false;
static ProfRec g_prof[4096];
static int g_prof_n = 0;

int gemm_prof_begin(double flops, cudaStream_t stream) {
  if (!g_prof_on || g_prof_n >= 4096) return -1;
  ProfRec& r = g_prof[g_prof_n];
  if (cudaEventCreate(&r.s) != cudaSuccess || cudaEventCreate(&r.e) != cudaSuccess) return -1;
  r.flops = flops;
  cudaEventRecord(r.s, stream);
  return g_prof_n++;
}
void gemm_prof_end(int slot, cudaStream_t stream) {
  if (slot >= 0) cudaEventRecord(g_prof[slot].e, stream);
}

int gemm_f16_impl2(const void* A, int64_t a_rows, int64_t a_extent, const int* tap_row, int64_t a_row_stride,
                   int64_t a_batch_stride, int batch, int ntaps, int64_t a_tap_stride, int cin, const void* W, int64_t ldw,
                   int N, const float* bias, const float* shift, int act, void* out, int out_dtype, int64_t ldo,
                   int64_t out_batch_stride, cudaStream_t stream, int accumulate = 0);

int gemm_f16_impl(const void* A, int64_t a_rows, int64_t a_row_stride, int64_t a_batch_stride, int batch, int ntaps,
                  int64_t a_tap_stride, int cin, const void* W, int64_t ldw, int N, const float* bias,
                  const float* shift, int act, void* out, int out_dtype, int64_t ldo, int64_t out_batch_stride,
                  cudaStream_t stream) {
  return gemm_f16_impl2(A, a_rows, a_rows, nullptr, a_row_stride, a_batch_stride, batch, ntaps, a_tap_stride, cin, W, ldw, N,
                        bias, shift, act, out, out_dtype, ldo, out_batch_stride, stream);
}

// a_rows: output rows per batch element; a_extent: rows of A that exist per batch element (rows outside
// [0, a_extent) read as zero); tap_row[t]: row offset of tap t relative to the output row (nullptr = 0).
int gemm_f16_impl2(const void* A, int64_t a_rows, int64_t a_extent, const int* tap_row, int64_t a_row_stride,
                   int64_t a_batch_stride, int batch, int ntaps, int64_t a_tap_stride, int cin, const void* W, int64_t ldw,
                   int N, const float* bias, const float* shift, int act, void* out, int out_dtype, int64_t ldo,
                   int64_t out_batch_stride, cudaStream_t stream, int accumulate) {
  W2V2_REQUIRE(!accumulate || (out_dtype == 1 && act == 0 && shift == nullptr),
               "w2v2_gemm_f16: accumulate needs fp32 output without activation / affine epilogue");
  W2V2_REQUIRE(ntaps >= 1 && ntaps <= 3, "w2v2_gemm_f16: ntaps=%d not in [1,3]", ntaps);
  W2V2_REQUIRE(shift == nullptr || N <= EPI2_MAX_N, "w2v2_gemm_f16: the per-batch affine epilogue holds at most %d columns", EPI2_MAX_N);
  W2V2_REQUIRE(cin % BK == 0, "w2v2_gemm_f16: cin=%d must be a multiple of %d", cin, BK);
  W2V2_REQUIRE(out_dtype == 0 || out_dtype == 1, "w2v2_gemm_f16: out_dtype must be 0 (f16) or 1 (f32)");
  W2V2_REQUIRE(act == 0 || act == 1, "w2v2_gemm_f16: act must be 0 (none) or 1 (gelu)");
  W2V2_REQUIRE(a_rows > 0 && batch > 0 && N > 0, "w2v2_gemm_f16: empty problem");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  static const bool conv0_2cta = []() { const char* e = getenv("W2V2_CONV0_2CTA"); return !(e != nullptr && e[0] == '0'); }();
  // the GroupNorm-affine + GELU epilogue of conv layer 0 is issue-bound, not tensor-bound: 128-wide tiles with two
  // CTAs (16 epilogue warps) per SM instead of one pair-CTA with 8
  const bool small_tiles = shift != nullptr && out_dtype == 0 && cin == BK && ntaps == 1 && conv0_2cta;
  const int BN = (N <= 128 || small_tiles) ? 128 : 256;
  const int CL = (BN == 256 && pair_enabled()) ? 2 : 1;         // CTA pairs for every wide GEMM
  const int osz = out_dtype == 1 ? 4 : 2;
  const uint64_t a_bstride = batch > 1 ? uint64_t(a_batch_stride) * 2 : uint64_t(a_extent) * uint64_t(a_row_stride) * 2;
  for (int t = 0; t < ntaps; ++t) {
    const __half* base = static_cast<const __half*>(A) + t * a_tap_stride;
    int rc = make_tmap_3d(&p.tmA[t], base, 2, cin, a_extent, batch, uint64_t(a_row_stride) * 2, a_bstride, BK, BM, 1, 128);
    if (rc) return rc;
    p.tap_row[t] = tap_row != nullptr ? tap_row[t] : 0;
  }
  int rc = make_tmap_3d(&p.tmB, W, 2, uint64_t(ntaps) * cin, N, 1, uint64_t(ldw) * 2, uint64_t(N) * ldw * 2, BK, BN / CL, 1, 128);
  if (rc) return rc;
  const uint64_t o_bstride = batch > 1 ? uint64_t(out_batch_stride) * osz : uint64_t(a_rows) * uint64_t(ldo) * osz;
  rc = make_tmap_3d(&p.tmOut, out, osz, N, a_rows, batch, uint64_t(ldo) * osz, o_bstride, out_dtype == 1 ? 32 : 64, 32, 1, 128);
  if (rc) return rc;
  p.bias = bias;
  p.shift = shift;
  p.ntaps = ntaps;
  p.kblocks_per_tap = cin / BK;
  p.m_tiles = int((a_rows + BM * CL - 1) / (BM * CL));
  p.n_tiles = (N + BN - 1) / BN;
  p.batch = batch;
  p.N = N;
  p.rows = a_rows;
  p.accumulate = accumulate;
  p.no_tail_skip = no_tail_skip();
  const int slot = gemm_prof_begin(2.0 * double(a_rows) * batch * ntaps * cin * N, stream);
  if (CL == 2) rc = out_dtype == 1 ? dispatch_epilogue<256, true, 2>(p, act, stream) : dispatch_epilogue<256, false, 2>(p, act, stream);
  else if (BN == 256) rc = out_dtype == 1 ? dispatch_epilogue<256, true, 1>(p, act, stream) : dispatch_epilogue<256, false, 1>(p, act, stream);
  else rc = out_dtype == 1 ? dispatch_epilogue<128, true, 1>(p, act, stream) : dispatch_epilogue<128, false, 1>(p, act, stream);
  gemm_prof_end(slot, stream);
  return rc;
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_gemm_profile_start(void) {
  for (int i = 0; i < g_prof_n; ++i) {
    cudaEventDestroy(g_prof[i].s);
    cudaEventDestroy(g_prof[i].e);
  }
  g_prof_n = 0;
  g_prof_on = true;
  return 0;
}

extern "C" int w2v2_gemm_profile_stop(double* total_ms, double* total_flops, int* launches) {
  g_prof_on = false;
  W2V2_CHECK_CUDA(cudaDeviceSynchronize());
  double ms = 0.0, fl = 0.0;
  for (int i = 0; i < g_prof_n; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof[i].s, g_prof[i].e) == cudaSuccess) ms += t;
    fl += g_prof[i].flops;
    cudaEventDestroy(g_prof[i].s);
    cudaEventDestroy(g_prof[i].e);
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = g_prof_n;
  g_prof_n = 0;
  return 0;
}

// out_act = gelu(A W^T + bias) and out_pre = A W^T + bias, both f16 [M, N]: the FFN1 GEMM of the training forward,
// which has to keep the pre-activation for the backward (saves the separate GELU pass over [M, FF]).
static int gemm_dual_gelu(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N, const float* bias,
                          void* out_act16, void* out_pre16, int64_t ldo, bool store_grad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  W2V2_REQUIRE(K % BK == 0 && N > 128 && bias != nullptr && M > 0, "w2v2_gemm_f16_dual_gelu: needs K %% 64 == 0, N > 128, a bias");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.no_tail_skip = no_tail_skip();
  const int CL = pair_enabled() ? 2 : 1;
  int rc = make_tmap_3d(&p.tmA[0], A, 2, K, M, 1, uint64_t(lda) * 2, uint64_t(M) * lda * 2, BK, BM, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmB, W, 2, K, N, 1, uint64_t(ldw) * 2, uint64_t(N) * ldw * 2, BK, 256 / CL, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmOut, out_act16, 2, N, M, 1, uint64_t(ldo) * 2, uint64_t(M) * ldo * 2, 64, 32, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmOut2, out_pre16, 2, N, M, 1, uint64_t(ldo) * 2, uint64_t(M) * ldo * 2, 64, 32, 1, 128);
  if (rc) return rc;
  p.bias = bias;
  p.ntaps = 1;
  p.kblocks_per_tap = K / BK;
  p.m_tiles = int((M + BM * CL - 1) / (BM * CL));
  p.n_tiles = (N + 255) / 256;
  p.batch = 1;
  p.N = N;
  p.rows = M;
  const int slot = gemm_prof_begin(2.0 * double(M) * K * N, stream);
  if (store_grad)
    rc = CL == 2 ? launch_gemm<256, false, 1, 1, 2, 1, 2>(p, stream) : launch_gemm<256, false, 1, 1, 1, 1, 2>(p, stream);
  else
    rc = CL == 2 ? launch_gemm<256, false, 1, 1, 2, 1, 1>(p, stream) : launch_gemm<256, false, 1, 1, 1, 1, 1>(p, stream);
  gemm_prof_end(slot, stream);
  return rc;
}

extern "C" int w2v2_gemm_f16_dual_gelu(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N,
                                       const float* bias, void* out_act16, void* out_pre16, int64_t ldo, void* stream) {
  return gemm_dual_gelu(A, M, lda, K, W, ldw, N, bias, out_act16, out_pre16, ldo, false, stream);
}

// Same GEMM, second output = gelu'(A W^T + bias) instead of the pre-activation itself (the only thing the backward
// computes from it): pairs with w2v2_gemm_f16_mul_colsum, whose epilogue then only multiplies.
extern "C" int w2v2_gemm_f16_dual_gelu_grad(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N,
                                            const float* bias, void* out_act16, void* out_grad16, int64_t ldo, void* stream) {
  return gemm_dual_gelu(A, M, lda, K, W, ldw, N, bias, out_act16, out_grad16, ldo, true, stream);
}

// dz = (A W^T) * gelu'(z) (f16 [M, N]) and dbias[n] += sum_r dz[r, n]: the FFN2 data-gradient GEMM of the training
// backward with the GELU backward and the FFN1 bias gradient done in its epilogue (saves writing and re-reading the
// [M, FF] gradient of the activation).
static int gemm_mul_epilogue(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N, const void* z16,
                             int64_t ldz, void* dz16, int64_t ldo, float* dbias, bool z_is_grad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  W2V2_REQUIRE(K % BK == 0 && N > 128 && N % 32 == 0 && M > 0, "w2v2_gemm_f16_gelu_bwd: needs K %% 64 == 0, N > 128, N %% 32 == 0");
  W2V2_REQUIRE(z16 != nullptr && dbias != nullptr && ldz % 8 == 0 && (reinterpret_cast<uintptr_t>(z16) & 15) == 0,
               "w2v2_gemm_f16_gelu_bwd: z must be 16-byte aligned with ldz %% 8 == 0, dbias is required");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.no_tail_skip = no_tail_skip();
  const int CL = pair_enabled() ? 2 : 1;
  int rc = make_tmap_3d(&p.tmA[0], A, 2, K, M, 1, uint64_t(lda) * 2, uint64_t(M) * lda * 2, BK, BM, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmB, W, 2, K, N, 1, uint64_t(ldw) * 2, uint64_t(N) * ldw * 2, BK, 256 / CL, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmOut, dz16, 2, N, M, 1, uint64_t(ldo) * 2, uint64_t(M) * ldo * 2, 64, 32, 1, 128);
  if (rc) return rc;
  p.aux = static_cast<const __half*>(z16);
  p.ld_aux = ldz;
  p.colsum = dbias;
  p.ntaps = 1;
  p.kblocks_per_tap = K / BK;
  p.m_tiles = int((M + BM * CL - 1) / (BM * CL));
  p.n_tiles = (N + 255) / 256;
  p.batch = 1;
  p.N = N;
  p.rows = M;
  const int slot = gemm_prof_begin(2.0 * double(M) * K * N, stream);
  if (z_is_grad) rc = CL == 2 ? launch_gemm<256, false, 3, 0, 2>(p, stream) : launch_gemm<256, false, 3, 0, 1>(p, stream);
  else rc = CL == 2 ? launch_gemm<256, false, 2, 0, 2>(p, stream) : launch_gemm<256, false, 2, 0, 1>(p, stream);
  gemm_prof_end(slot, stream);
  return rc;
}

extern "C" int w2v2_gemm_f16_gelu_bwd(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N,
                                      const void* z16, int64_t ldz, void* dz16, int64_t ldo, float* dbias, void* stream) {
  return gemm_mul_epilogue(A, M, lda, K, W, ldw, N, z16, ldz, dz16, ldo, dbias, false, stream);
}

// out = (A W^T) * mul (f16 [M, N]) and colsum[n] += sum_r out[r, n]: the gelu'-multiply epilogue when the forward
// kept gelu'(z) itself (w2v2_gemm_f16_dual_gelu_grad).
extern "C" int w2v2_gemm_f16_mul_colsum(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N,
                                        const void* mul16, int64_t ldm, void* out16, int64_t ldo, float* colsum, void* stream) {
  return gemm_mul_epilogue(A, M, lda, K, W, ldw, N, mul16, ldm, out16, ldo, colsum, true, stream);
}

extern "C" int w2v2_gemm_f16_taps(const void* A, int64_t out_rows, int64_t a_extent, const int* tap_row, int64_t a_row_stride,
                                  int64_t a_batch_stride, int batch, int ntaps, int cin, const void* W, int64_t ldw, int N,
                                  void* out, int out_dtype, int64_t ldo, int64_t out_batch_stride, void* stream_) {
  W2V2_REQUIRE(tap_row != nullptr, "w2v2_gemm_f16_taps: tap_row is required");
  return gemm_f16_impl2(A, out_rows, a_extent, tap_row, a_row_stride, a_batch_stride, batch, ntaps, 0, cin, W, ldw, N, nullptr,
                        nullptr, 0, out, out_dtype, ldo, out_batch_stride, static_cast<cudaStream_t>(stream_));
}

extern "C" int w2v2_gemm_f16_accum(const void* A, int64_t M, int64_t lda, int K, const void* W, int64_t ldw, int N,
                                   float* out32, int64_t ldo, void* stream_) {
  return gemm_f16_impl2(A, M, M, nullptr, lda, 0, 1, 1, 0, K, W, ldw, N, nullptr, nullptr, 0, out32, 1, ldo, 0,
                        static_cast<cudaStream_t>(stream_), 1);
}

extern "C" int w2v2_gemm_f16(const void* A, int64_t a_rows, int64_t a_row_stride, int64_t a_batch_stride, int batch,
                             int ntaps, int64_t a_tap_stride, int cin, const void* W, int64_t ldw, int N,
                             const float* bias, int act, void* out, int out_dtype, int64_t ldo,
                             int64_t out_batch_stride, void* stream_) {
  return gemm_f16_impl(A, a_rows, a_row_stride, a_batch_stride, batch, ntaps, a_tap_stride, cin, W, ldw, N, bias, nullptr,
                       act, out, out_dtype, ldo, out_batch_stride, static_cast<cudaStream_t>(stream_));
}
