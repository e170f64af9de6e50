// Weight-gradient GEMM on tcgen05:   dW[n, k] += sum_m dY[m, n] * X[m, k]
//
// The contraction runs over the ROW index of both row-major operands (dY [M, N], X [M, K]), i.e.
// both UMMA operands are MN-major: the [64 rows(m) x 64 columns] boxes TMA delivers (128-byte swizzle)
// are consumed as they are -- no transposed copies of activations or gradients ever exist.
//   A tile (128 n x 64 m) = 2 boxes, B tile (256 k x 64 m) = 4 boxes; MN-major SW128 descriptors:
//   SBO = 1024 B (8 m-rows), LBO = 8192 B (next 64-wide block), 16 m-rows (2048 B) per MMA K-step.
// The M (row) range is split over CTAs (split-K) so that small weight matrices still fill the GPU; the
// fp32 partial tiles are accumulated into dW by TMA reduce-add stores (cp.reduce.async.bulk.tensor),
// which is also what makes dW accumulate across micro-batches like torch's .grad does.
// Warp roles as in gemm_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue.
#include <string.h>

#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 int swizzle_bytes);
int device_sm_count();

constexpr int WG_BM = 128;        // output rows per tile  (columns of dY)
constexpr int WG_BN = 256;        // output cols per tile  (columns of X)
constexpr int WG_BK = 64;         // contraction rows per stage
constexpr int WG_STAGES = 4;
constexpr int WG_A_BYTES = WG_BM * WG_BK * 2;     // 16 KB
constexpr int WG_B_BYTES = WG_BN * WG_BK * 2;     // 32 KB
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_EPI_WARPS = 8;
constexpr int WG_THREADS = 64 + WG_EPI_WARPS * 32;
constexpr int WG_WSTAGE = 32 * 128;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + WG_EPI_WARPS * WG_WSTAGE + 256;
static_assert(WG_SMEM <= 232448, "shared memory budget");

struct alignas(64) WgradParams {
  CUtensorMap tmA;      // dY [M, N]: dims {N, M, 1}, box {64, 64, 1}
  CUtensorMap tmB;      // X  [M, K]: dims {K, M, 1}, box {64, 64, 1}
  CUtensorMap tmOut;    // dW [N, K] f32: dims {K, N, 1}, box {32, 32, 1}
  int n_tiles, k_tiles, splits;
  int kblocks, kblocks_per_split;
  int kb_per_batch;     // contraction blocks per batch element (rows of one element never share a block with the next)
  int N, K;
};

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1) gemm_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_trigger();
  uint8_t* wstage = smem + WG_STAGES * WG_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(wstage + WG_EPI_WARPS * WG_WSTAGE);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* tmem_full = empty_bar + WG_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.n_tiles * p.k_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tmA);
    prefetch_tensormap(&p.tmB);
    prefetch_tensormap(&p.tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], WG_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * WG_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // item -> (split, output row tile, output column tile); split fastest so the CTAs that run together
  // reduce into different addresses as little as possible... they do not: they share the tile, which is
  // what we want for L2 locality of the reduce-adds.
  auto decode = [&](int item, int& kt, int& nt, int& sp) {
    sp = item % p.splits;
    const int rest = item / p.splits;
    kt = rest % p.k_tiles;
    nt = rest / p.k_tiles;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int kt, nt, sp;
        decode(item, kt, nt, sp);
        const int kb0 = sp * p.kblocks_per_split;
        const int kb1 = min(p.kblocks, kb0 + p.kblocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * WG_STAGE_BYTES;
          uint8_t* sb = sa + WG_A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], WG_STAGE_BYTES);
#pragma unroll
          const int bi = kb / p.kb_per_batch, r0 = (kb - bi * p.kb_per_batch) * WG_BK;
#pragma unroll
          for (int i = 0; i < WG_BM / 64; ++i) tma_load_3d(sa + i * 8192, &p.tmA, &full_bar[stage], nt * WG_BM + i * 64, r0, bi);
#pragma unroll
          for (int i = 0; i < WG_BN / 64; ++i) tma_load_3d(sb + i * 8192, &p.tmB, &full_bar[stage], kt * WG_BN + i * 64, r0, bi);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(WG_BM, WG_BN, 1, 1);       // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int kt, nt, sp;
        decode(item, kt, nt, sp);
        const int kb0 = sp * p.kblocks_per_split;
        const int kb1 = min(p.kblocks, kb0 + p.kblocks_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * WG_BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * WG_STAGE_BYTES);
          const uint32_t b_addr = a_addr + WG_A_BYTES;
#pragma unroll
          for (int k = 0; k < WG_BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc(a_addr + k * 2048, 8192, 1024, 2);
            const uint64_t bdesc = make_smem_desc(b_addr + k * 2048, 8192, 1024, 2);
            umma_f16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    constexpr int HALF_COLS = WG_BN / 2;
    constexpr int NCHUNK = HALF_COLS / 32;
    uint8_t* mystage = wstage + ew * WG_WSTAGE;
    uint8_t* crow = mystage + lane * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int kt, nt, sp;
      decode(item, kt, nt, sp);
      const bool empty_split = sp * p.kblocks_per_split >= p.kblocks;   // nothing accumulated (cannot happen by construction)
      mbar_wait(&tmem_full[acc], acc_phase);
      __syncwarp();
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * WG_BN + half * HALF_COLS;
      const int col0 = kt * WG_BN + half * HALF_COLS;
      const int row0 = nt * WG_BM + quarter * 32;
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_addr + c * 32, r);
        tmem_ld_wait();
        if (c == NCHUNK - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int pc = q ^ (lane & 7);
          *reinterpret_cast<uint4*>(crow + pc * 16) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        const int scol = col0 + c * 32;
        if (lane == 0 && !empty_split && row0 < p.N && scol < p.K) {
          tma_reduce_add_3d(&p.tmOut, mystage, scol, row0, 0);
          tma_store_commit();
        }
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
    tmem_dealloc(tmem_base, 2 * WG_BN);
  }
}

int gemm_prof_begin(double flops, cudaStream_t stream);
void gemm_prof_end(int slot, cudaStream_t stream);

}  // namespace w2v2

using namespace w2v2;

static int wgrad_impl(const void* dY, int64_t ldy, int64_t dy_batch_stride, const void* X, int64_t ldx, int64_t x_batch_stride,
                      int64_t rows, int batch, int N, int K, float* dW, int64_t ldw, cudaStream_t stream);

extern "C" int w2v2_gemm_wgrad_f16(const void* dY, int64_t ldy, const void* X, int64_t ldx, int64_t M, int N, int K,
                                   float* dW, int64_t ldw, void* stream_) {
  return wgrad_impl(dY, ldy, M * ldy, X, ldx, M * ldx, M, 1, N, K, dW, ldw, static_cast<cudaStream_t>(stream_));
}

// Batched contraction: dW[n, k] += sum_b sum_{r < rows} dY[b][r, n] * X[b][r, k], each operand with its own row pitch and
// batch stride (strided Conv1d weight gradient: X rows are every `stride`-th input frame of one tap).
extern "C" int w2v2_gemm_wgrad_f16_batched(const void* dY, int64_t ldy, int64_t dy_batch_stride, const void* X, int64_t ldx,
                                           int64_t x_batch_stride, int64_t rows, int batch, int N, int K, float* dW,
                                           int64_t ldw, void* stream_) {
  return wgrad_impl(dY, ldy, dy_batch_stride, X, ldx, x_batch_stride, rows, batch, N, K, dW, ldw,
                    static_cast<cudaStream_t>(stream_));
}

static int wgrad_impl(const void* dY, int64_t ldy, int64_t dy_batch_stride, const void* X, int64_t ldx, int64_t x_batch_stride,
                      int64_t rows, int batch, int N, int K, float* dW, int64_t ldw, cudaStream_t stream) {
  const int64_t M = rows * batch;
  W2V2_REQUIRE(rows > 0 && batch > 0 && N > 0 && K > 0, "w2v2_gemm_wgrad_f16: empty problem");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap_3d(&p.tmA, dY, 2, N, rows, batch, uint64_t(ldy) * 2, uint64_t(dy_batch_stride) * 2, 64, WG_BK, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmB, X, 2, K, rows, batch, uint64_t(ldx) * 2, uint64_t(x_batch_stride) * 2, 64, WG_BK, 1, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&p.tmOut, dW, 4, K, N, 1, uint64_t(ldw) * 4, uint64_t(N) * ldw * 4, 32, 32, 1, 128);
  if (rc) return rc;
  p.n_tiles = (N + WG_BM - 1) / WG_BM;
  p.k_tiles = (K + WG_BN - 1) / WG_BN;
  p.kb_per_batch = int((rows + WG_BK - 1) / WG_BK);
  p.kblocks = p.kb_per_batch * batch;
  const int out_tiles = p.n_tiles * p.k_tiles;
  const int sms = device_sm_count();
  // Split the contraction: pick the split count that minimises (waves of work items) x (k-blocks per item
  // + a fixed per-item cost for pipeline fill and the reduce-add epilogue).  "About one wave" is not good
  // enough: 72 output tiles x 3 splits = 216 items are two waves of 50 k-blocks, 72 x 2 = 144 items one
  // wave of 75.
  const int max_splits = (p.kblocks + 7) / 8;
  int splits = 1;
  long long best = -1;
  for (int s = 1; s <= max_splits; ++s) {
    const int per = (p.kblocks + s - 1) / s;
    const int real = (p.kblocks + per - 1) / per;                  // no empty split
    const long long waves = (static_cast<long long>(out_tiles) * real + sms - 1) / sms;
    const long long cost = waves * (per + 6);
    if (best < 0 || cost < best) {
      best = cost;
      splits = real;
    }
  }
  p.kblocks_per_split = (p.kblocks + splits - 1) / splits;
  p.splits = (p.kblocks + p.kblocks_per_split - 1) / p.kblocks_per_split;      // no empty split
  p.N = N;
  p.K = K;
  static bool configured = false;
  if (!configured) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    configured = true;
  }
  const int items = out_tiles * p.splits;
  const int grid = items < sms ? items : sms;
  const int slot = gemm_prof_begin(2.0 * double(M) * N * K, stream);
  W2V2_CHECK_CUDA(launch_k(gemm_wgrad_kernel, dim3(grid), dim3(WG_THREADS), WG_SMEM, stream, 1, p));
  gemm_prof_end(slot, stream);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}
