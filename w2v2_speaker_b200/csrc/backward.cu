// Backward-pass kernels that are not GEMMs: LayerNorm backward (+ parameter-gradient reductions),
// GELU backward, column sums (bias gradients), softmax-CE backward, pooling backward, transposed
// fp16 weight copies for the dgrad GEMMs, fused Adam.  All activation gradients are carried in a
// loss-scaled fp16 copy (GEMM operand) next to the fp32 residual-stream gradient.
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

int device_sm_count();

static inline int grid_cap(int64_t blocks, int per_sm) {
  const int64_t cap = int64_t(device_sm_count()) * per_sm;
  return int(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// =================================================================================================
// w [R, C] f32  ->  wT [C, ldt] f16 (transposed, columns >= R zero-filled up to ldt), optional row scale
__global__ void cast_f16_transpose_kernel(const float* __restrict__ w, __half* __restrict__ wt, int R, int C, int ldt,
                                          const float* __restrict__ row_scale) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) v = w[int64_t(r) * C + c] * (row_scale != nullptr ? row_scale[r] : 1.0f);
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < ldt) wt[int64_t(c) * ldt + r] = __float2half_rn(tile[threadIdx.x][i]);
  }
}

// =================================================================================================
// LayerNorm backward.  x = drop(xa (+ bias)) (+ residual) is recomputed (as are mean / rstd and the dropout
// mask), so the forward saves nothing extra.  dy = dy_a (+ dy_b).  Outputs dx (f32, the gradient of the
// residual input) and/or dx16 (f16, the gradient of the branch input: dx times the dropout mask);
// dgamma / dbeta are accumulated with one atomicAdd per column per block.
// EXACT: H == 128 * MAXV, the per-slot bounds checks (and the branches that would serialise the loads) vanish.
// The per-warp dgamma / dbeta partial sums live in shared memory (registers are what limits the number of
// rows in flight per SM), one private strip per warp, reduced over the block's warps at the end.
// FROM_Y: the normalised activation is recovered from the LayerNorm OUTPUT, xhat = (y - beta) / gamma, with the
// row's rstd saved by the forward (w2v2_layernorm_ex2): one [rows, H] fp32 stream (y) instead of two (xa and the
// residual) and no mean / variance reductions.  `xa_` then carries y, `bias` carries beta, `residual` carries the
// saved rstd [rows].
constexpr int LNB_WARPS = 4;
__device__ __forceinline__ float ln_safe_inv(float g) { return fabsf(g) > 1e-30f ? 1.0f / g : 0.f; }
template <bool XA_F32, int MAXV, bool EXACT, bool FROM_Y = false>
__global__ void __launch_bounds__(LNB_WARPS * 32, MAXV <= 6 ? 4 : 3) layernorm_bwd_kernel(const float* __restrict__ dy_a, const float* __restrict__ dy_b,
                                                            const void* __restrict__ xa_, const float* __restrict__ bias,
                                                            const float* __restrict__ residual,
                                                            const float* __restrict__ gamma, float eps,
                                                            float* __restrict__ dx32, __half* __restrict__ dx16,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            float* __restrict__ dbias,
                                                            int64_t rows, int H, uint32_t thr, float inv_keep,
                                                            uint64_t seed) {
  pdl_trigger();
  pdl_wait();
  // MAXV = ceil(H / 128) float4 slots per lane (4: H <= 512, 6: H <= 768, 8: H <= 1024)
  // strips: 0 = dgamma, 1 = dbeta, 2 = dbias (column sums of the branch gradient, i.e. the bias gradient of
  // the Linear that produced xa)
  __shared__ float4 sacc[LNB_WARPS][3][MAXV * 32];
  // FROM_Y: 1 / gamma of the block's columns, formed once instead of per row (H <= 768: with H = 1024 the accumulator
  // strips already fill the 48 KB of static shared memory, and the division stays in the row loop)
  constexpr bool STAGE_INV = FROM_Y && MAXV <= 6;
  __shared__ float4 sinvg[STAGE_INV ? MAXV * 32 : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if constexpr (STAGE_INV) {
    for (int v = threadIdx.x; v < MAXV * 32; v += LNB_WARPS * 32) {
      float4 ig = make_float4(0.f, 0.f, 0.f, 0.f);
      if (EXACT || v * 4 < H) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + v * 4));
        ig = make_float4(ln_safe_inv(gm.x), ln_safe_inv(gm.y), ln_safe_inv(gm.z), ln_safe_inv(gm.w));
      }
      sinvg[v] = ig;
    }
    __syncthreads();
  }
  float4* sg = sacc[warp][0];
  float4* sb = sacc[warp][1];
  float4* sd = sacc[warp][2];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    sg[i * 32 + lane] = make_float4(0, 0, 0, 0);
    sb[i * 32 + lane] = make_float4(0, 0, 0, 0);
    sd[i * 32 + lane] = make_float4(0, 0, 0, 0);
  }
  const int64_t wstride = int64_t(gridDim.x) * LNB_WARPS;
  const DropKeys dkeys = drop_keys(seed);
  for (int64_t row = int64_t(blockIdx.x) * LNB_WARPS + warp; row < rows; row += wstride) {
    float4 x[MAXV], d[MAXV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (EXACT || c < H) {
        float4 a;
        if constexpr (XA_F32) {
          a = *reinterpret_cast<const float4*>(static_cast<const float*>(xa_) + row * H + c);
        } else {
          const uint2 q = *reinterpret_cast<const uint2*>(static_cast<const __half*>(xa_) + row * H + c);
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
          const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
          a = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        float4 g = *reinterpret_cast<const float4*>(dy_a + row * H + c);
        if (dy_b != nullptr) {
          const float4 g2 = *reinterpret_cast<const float4*>(dy_b + row * H + c);
          g.x += g2.x; g.y += g2.y; g.z += g2.z; g.w += g2.w;
        }
        d[i] = g;
        if constexpr (FROM_Y) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));          // beta
          float4 ig;
          if constexpr (STAGE_INV) {
            ig = sinvg[i * 32 + lane];
          } else {
            const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
            ig = make_float4(ln_safe_inv(gm.x), ln_safe_inv(gm.y), ln_safe_inv(gm.z), ln_safe_inv(gm.w));
          }
          a.x = (a.x - bb.x) * ig.x; a.y = (a.y - bb.y) * ig.y;
          a.z = (a.z - bb.z) * ig.z; a.w = (a.w - bb.w) * ig.w;
        } else {
          if (bias != nullptr) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
            a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
          }
          if (thr != 0) dropout4(a, dkeys, row * H + c, thr, inv_keep);
          if (residual != nullptr) {
            const float4 r = *reinterpret_cast<const float4*>(residual + row * H + c);
            a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
          }
        }
        x[i] = a;
        sum += (a.x + a.y) + (a.z + a.w);
      }
    }
    float rstd;
    if constexpr (FROM_Y) {
      rstd = __ldg(residual + row);                   // saved by the forward
    } else {
      const float mean = warp_sum(sum) / float(H);
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (EXACT || c < H) {
          x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
          sq += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
        }
      }
      rstd = rsqrtf(warp_sum(sq) / float(H) + eps);
    }
    // xhat = x * rstd ; g = dy * gamma ; dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (EXACT || c < H) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
        if constexpr (!FROM_Y) { x[i].x *= rstd; x[i].y *= rstd; x[i].z *= rstd; x[i].w *= rstd; }
        float4 ag = sg[i * 32 + lane], ab = sb[i * 32 + lane];
        ag.x += d[i].x * x[i].x; ag.y += d[i].y * x[i].y; ag.z += d[i].z * x[i].z; ag.w += d[i].w * x[i].w;
        ab.x += d[i].x; ab.y += d[i].y; ab.z += d[i].z; ab.w += d[i].w;
        sg[i * 32 + lane] = ag;
        sb[i * 32 + lane] = ab;
        d[i].x *= gm.x; d[i].y *= gm.y; d[i].z *= gm.z; d[i].w *= gm.w;
        s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
        s2 += (d[i].x * x[i].x + d[i].y * x[i].y) + (d[i].z * x[i].z + d[i].w * x[i].w);
      }
    }
    s1 = warp_sum(s1) / float(H);
    s2 = warp_sum(s2) / float(H);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (EXACT || c < H) {
        float4 o;
        o.x = rstd * (d[i].x - s1 - x[i].x * s2);
        o.y = rstd * (d[i].y - s1 - x[i].y * s2);
        o.z = rstd * (d[i].z - s1 - x[i].z * s2);
        o.w = rstd * (d[i].w - s1 - x[i].w * s2);
        if (dx32 != nullptr) *reinterpret_cast<float4*>(dx32 + row * H + c) = o;
        if (dx16 != nullptr || dbias != nullptr) {
          if (thr != 0) dropout4(o, dkeys, row * H + c, thr, inv_keep);
          if (dx16 != nullptr) {
            uint2 q;
            q.x = pack_half2(o.x, o.y);
            q.y = pack_half2(o.z, o.w);
            *reinterpret_cast<uint2*>(dx16 + row * H + c) = q;
          }
          if (dbias != nullptr) {
            float4 a = sd[i * 32 + lane];
            a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
            sd[i * 32 + lane] = a;
          }
        }
      }
    }
  }
  if (dgamma == nullptr && dbeta == nullptr && dbias == nullptr) return;
  // block reduction of the per-warp column sums, one (vector) atomic per 4 columns per block
  __syncthreads();
  for (int v = threadIdx.x; v < H / 4; v += LNB_WARPS * 32) {
    float4 tg = sacc[0][0][v], tb = sacc[0][1][v], td = sacc[0][2][v];
#pragma unroll
    for (int w = 1; w < LNB_WARPS; ++w) {
      const float4 a = sacc[w][0][v], b = sacc[w][1][v], d = sacc[w][2][v];
      tg.x += a.x; tg.y += a.y; tg.z += a.z; tg.w += a.w;
      tb.x += b.x; tb.y += b.y; tb.z += b.z; tb.w += b.w;
      td.x += d.x; td.y += d.y; td.z += d.z; td.w += d.w;
    }
    if (dbias != nullptr)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbias + 4 * v), "f"(td.x), "f"(td.y), "f"(td.z), "f"(td.w) : "memory");
    if (dgamma != nullptr)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dgamma + 4 * v), "f"(tg.x), "f"(tg.y), "f"(tg.z), "f"(tg.w) : "memory");
    if (dbeta != nullptr)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbeta + 4 * v), "f"(tb.x), "f"(tb.y), "f"(tb.z), "f"(tb.w) : "memory");
  }
}

// =================================================================================================
// LayerNorm backward from the LayerNorm output with the rows STREAMED through shared memory (the form the encoder layers
// use; default, W2V2_LNB_RING=0 restores the register-resident kernel above).  The ncu capture of that kernel
// (profiles/r02_d_hot_kernels_ncu.txt, r02_h_layernorm_bwd.txt) shows what holds it at 63 % of the copy bandwidth: a warp
// loads its row (y, dy: 6 KB at H = 768) into 48 registers, waits, computes, and only then issues the next row's loads --
// 45 % of the warp stalls are the L1TEX scoreboard, DRAM is 33 % busy, and with 128 registers per thread only 13 warps
// per SM take turns.  (Two warps per row -- half the registers, twice the warps -- was tried first: 27.7 us against
// 24.6 us; the bytes in flight per SM stay the same and a barrier per row is added.)
// Here every warp owns a two-stage ring of row buffers in shared memory, filled by 1-D bulk copies (cp.async.bulk +
// mbarrier transaction count) that its lane 0 issues one row AHEAD of the arithmetic: loads are in flight all the time,
// the arithmetic reads the row from shared memory, and the registers go to the dgamma / dbeta / dbias column sums instead
// (no shared-memory read-modify-write per row).  Same arithmetic per element in the same order.
constexpr int LNR_WARPS = 4, LNR_ST = 2;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int MAXV>
__global__ void __launch_bounds__(LNR_WARPS * 32, 3) layernorm_bwd_ring_kernel(
    const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ beta,
    const float* __restrict__ rstd_saved, const float* __restrict__ gamma, float* __restrict__ dx32, __half* __restrict__ dx16,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, int64_t rows, uint32_t thr,
    float inv_keep, uint64_t seed) {
  constexpr int H = MAXV * 128;
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(128) unsigned char lnr_smem[];
  float* ring = reinterpret_cast<float*>(lnr_smem);                            // [warps][stages][y | dy][H]
  float4* sconst = reinterpret_cast<float4*>(ring + LNR_WARPS * LNR_ST * 2 * H);       // [3][H / 4]: beta, 1 / gamma, gamma
  uint64_t* bars = reinterpret_cast<uint64_t*>(sconst + 3 * (H / 4));          // [warps][stages]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int v = threadIdx.x; v < H / 4; v += LNR_WARPS * 32) {
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + v * 4));
    sconst[v] = __ldg(reinterpret_cast<const float4*>(beta + v * 4));
    sconst[H / 4 + v] = make_float4(ln_safe_inv(gm.x), ln_safe_inv(gm.y), ln_safe_inv(gm.z), ln_safe_inv(gm.w));
    sconst[2 * (H / 4) + v] = gm;
  }
  if (threadIdx.x < LNR_WARPS * LNR_ST) mbar_init(bars + threadIdx.x, 1);
  fence_barrier_init();
  __syncthreads();
  float* my_ring = ring + warp * LNR_ST * 2 * H;
  uint64_t* my_bars = bars + warp * LNR_ST;
  const int64_t wstride = int64_t(gridDim.x) * LNR_WARPS;
  const int64_t row0 = int64_t(blockIdx.x) * LNR_WARPS + warp;
  auto issue = [&](int64_t row, int st) {                       // lane 0: both halves of the stage are free
    mbar_arrive_expect_tx(my_bars + st, 2 * H * 4);
    bulk_load_1d(my_ring + st * 2 * H, y + row * H, H * 4, my_bars + st);
    bulk_load_1d(my_ring + st * 2 * H + H, dy + row * H, H * 4, my_bars + st);
  };
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < LNR_ST; ++st)
      if (row0 + st * wstride < rows) issue(row0 + st * wstride, st);
  }
  float4 ag[MAXV], ab[MAXV], ad[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) ag[i] = ab[i] = ad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const DropKeys dkeys = drop_keys(seed);
  int it = 0;
  for (int64_t row = row0; row < rows; row += wstride, ++it) {
    const int st = it % LNR_ST;
    const float rstd = __ldg(rstd_saved + row);
    mbar_wait(my_bars + st, (it / LNR_ST) & 1);
    const float4* ybuf = reinterpret_cast<const float4*>(my_ring + st * 2 * H);
    const float4* dbuf = ybuf + H / 4;
    float4 x[MAXV], g[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = i * 32 + lane;
      float4 a = ybuf[v];
      const float4 d = dbuf[v];
      const float4 bb = sconst[v], ig = sconst[H / 4 + v], gm = sconst[2 * (H / 4) + v];
      a.x = (a.x - bb.x) * ig.x; a.y = (a.y - bb.y) * ig.y; a.z = (a.z - bb.z) * ig.z; a.w = (a.w - bb.w) * ig.w;
      ag[i].x += d.x * a.x; ag[i].y += d.y * a.y; ag[i].z += d.z * a.z; ag[i].w += d.w * a.w;
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
      float4 q = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += (q.x + q.y) + (q.z + q.w);
      s2 += (q.x * a.x + q.y * a.y) + (q.z * a.z + q.w * a.w);
      x[i] = a;
      g[i] = q;
    }
    // this warp is done with the stage: refill it with the row LNR_ST iterations ahead
    __syncwarp();
    if (lane == 0 && row + LNR_ST * wstride < rows) issue(row + LNR_ST * wstride, st);
    s1 = warp_sum(s1) / float(H);
    s2 = warp_sum(s2) / float(H);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 o;
      o.x = rstd * (g[i].x - s1 - x[i].x * s2);
      o.y = rstd * (g[i].y - s1 - x[i].y * s2);
      o.z = rstd * (g[i].z - s1 - x[i].z * s2);
      o.w = rstd * (g[i].w - s1 - x[i].w * s2);
      if (dx32 != nullptr) *reinterpret_cast<float4*>(dx32 + row * H + c) = o;
      if (dx16 != nullptr || dbias != nullptr) {
        if (thr != 0) dropout4(o, dkeys, row * H + c, thr, inv_keep);
        if (dx16 != nullptr) {
          uint2 q;
          q.x = pack_half2(o.x, o.y);
          q.y = pack_half2(o.z, o.w);
          *reinterpret_cast<uint2*>(dx16 + row * H + c) = q;
        }
        ad[i].x += o.x; ad[i].y += o.y; ad[i].z += o.z; ad[i].w += o.w;
      }
    }
  }
  if (dgamma == nullptr && dbeta == nullptr && dbias == nullptr) return;
  // block reduction of the per-warp column sums through the (now idle) ring, one vector atomic per 4 columns per block
  __syncthreads();                              // every warp has consumed its last stage: the ring is free
  float4* red = reinterpret_cast<float4*>(lnr_smem);                           // [warps][3][H / 4]
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = i * 32 + lane;
    red[(warp * 3 + 0) * (H / 4) + v] = ag[i];
    red[(warp * 3 + 1) * (H / 4) + v] = ab[i];
    red[(warp * 3 + 2) * (H / 4) + v] = ad[i];
  }
  __syncthreads();
  for (int v = threadIdx.x; v < H / 4; v += LNR_WARPS * 32) {
    float4 tg = red[v], tb = red[(H / 4) + v], td = red[2 * (H / 4) + v];
#pragma unroll
    for (int w = 1; w < LNR_WARPS; ++w) {
      const float4 a = red[(w * 3 + 0) * (H / 4) + v], b = red[(w * 3 + 1) * (H / 4) + v], d = red[(w * 3 + 2) * (H / 4) + v];
      tg.x += a.x; tg.y += a.y; tg.z += a.z; tg.w += a.w;
      tb.x += b.x; tb.y += b.y; tb.z += b.z; tb.w += b.w;
      td.x += d.x; td.y += d.y; td.z += d.z; td.w += d.w;
    }
    if (dbias != nullptr)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbias + 4 * v), "f"(td.x), "f"(td.y), "f"(td.z), "f"(td.w) : "memory");
    if (dgamma != nullptr)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dgamma + 4 * v), "f"(tg.x), "f"(tg.y), "f"(tg.z), "f"(tg.w) : "memory");
    if (dbeta != nullptr)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbeta + 4 * v), "f"(tb.x), "f"(tb.y), "f"(tb.z), "f"(tb.w) : "memory");
  }
}

template <int MAXV>
static int launch_lnb_ring(const float* dy, const float* y32, const float* rstd, const float* gamma, const float* beta,
                           float* dx32, void* dx16, float* dgamma, float* dbeta, float* dbias, int64_t rows, uint32_t thr,
                           float inv_keep, uint64_t drop_seed, cudaStream_t st) {
  constexpr int H = MAXV * 128;
  constexpr int smem = LNR_WARPS * LNR_ST * 2 * H * 4 + 3 * H * 4 + LNR_WARPS * LNR_ST * 8;
  static_assert(LNR_WARPS * 3 * H * 4 <= LNR_WARPS * LNR_ST * 2 * H * 4, "the final reduction borrows the ring");
  static bool configured = false;
  if (!configured) {
    W2V2_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_ring_kernel<MAXV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  // one wave of resident blocks, every warp the same number of rows (+-1)
  constexpr int per_sm = 3 * (smem + 1024) <= 227 * 1024 ? 3 : 2;           // H = 1024: two 78 KB blocks per SM
  const int64_t want = (rows + LNR_WARPS - 1) / LNR_WARPS, slots = int64_t(device_sm_count()) * per_sm;
  const int64_t per_warp = (want + slots - 1) / slots;
  const int grid = int((want + per_warp - 1) / per_warp);
  W2V2_CHECK_CUDA(launch_k(layernorm_bwd_ring_kernel<MAXV>, dim3(grid), dim3(LNR_WARPS * 32), size_t(smem), st, 1, dy, y32, beta,
                           rstd, gamma, dx32, (__half*)dx16, dgamma, dbeta, dbias, rows, thr, inv_keep, drop_seed));
  return 0;
}

// =================================================================================================
// dz = dg * gelu'(z)  (f16 in / f16 out), gelu'(z) = Phi(z) + z * phi(z)
__device__ __forceinline__ float gelu_grad(float z) {
  const float a = fabsf(z);
  const float u = fmaf(fminf(a, W2V2_GELU_A), 2.0f / W2V2_GELU_A, -1.0f);
  float q = fmaf(W2V2_GELU_C8, u, W2V2_GELU_C7);
  q = fmaf(q, u, W2V2_GELU_C6);
  q = fmaf(q, u, W2V2_GELU_C5);
  q = fmaf(q, u, W2V2_GELU_C4);
  q = fmaf(q, u, W2V2_GELU_C3);
  q = fmaf(q, u, W2V2_GELU_C2);
  q = fmaf(q, u, W2V2_GELU_C1);
  q = fmaf(q, u, W2V2_GELU_C0);
  const float tail = fast_ex2(q);                                   // Phi(-|z|)
  const float cdf = z >= 0.f ? 1.0f - tail : tail;
  const float pdf = 0.3989422804014327f * fast_ex2(z * z * -0.72134752044448170f);
  return fmaf(z, pdf, cdf);
}

__global__ void gelu_bwd_kernel(const __half* __restrict__ dg, const __half* __restrict__ z, __half* __restrict__ dz,
                                int64_t n) {
  int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    const uint4 a = *reinterpret_cast<const uint4*>(dg + i);
    const uint4 b = *reinterpret_cast<const uint4*>(z + i);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
    uint4 o;
    uint32_t* oo = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 g = __half22float2(ah[j]);
      const float2 zz = __half22float2(bh[j]);
      oo[j] = pack_half2(g.x * gelu_grad(zz.x), g.y * gelu_grad(zz.y));
    }
    *reinterpret_cast<uint4*>(dz + i) = o;
  }
}

// dz = dg * gelu'(z) over [rows, cols] f16, plus dbias[c] += sum_r dz[r, c] (the bias gradient of the Linear
// that produced z) in the same pass: block = 32 column lanes (8 columns each) x 8 row lanes.
__global__ void __launch_bounds__(256) gelu_bwd_colsum_kernel(const __half* __restrict__ dg, const __half* __restrict__ z,
                                                              __half* __restrict__ dz, int64_t rows, int cols,
                                                              float* __restrict__ dbias) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sm[8][256 + 8];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + cl * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c < cols) {
    for (int64_t r = int64_t(blockIdx.y) * 8 + rl; r < rows; r += int64_t(gridDim.y) * 8) {
      const uint4 a = *reinterpret_cast<const uint4*>(dg + r * cols + c);
      const uint4 b = *reinterpret_cast<const uint4*>(z + r * cols + c);
      const __half2* ah = reinterpret_cast<const __half2*>(&a);
      const __half2* bh = reinterpret_cast<const __half2*>(&b);
      uint4 o;
      uint32_t* oo = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 g = __half22float2(ah[j]);
        const float2 zz = __half22float2(bh[j]);
        const __half2 h = __floats2half2_rn(g.x * gelu_grad(zz.x), g.y * gelu_grad(zz.y));
        oo[j] = *reinterpret_cast<const uint32_t*>(&h);
        const float2 f = __half22float2(h);              // sum what the weight-gradient GEMM will see
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
      *reinterpret_cast<uint4*>(dz + r * cols + c) = o;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[rl][cl * 8 + j] = acc[j];
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x];
    atomicAdd(dbias + cc, t);
  }
}

// =================================================================================================
// out[c] += sum_r x[r, c] * scale   (bias gradients).  x f16 or f32 [rows, ld]; one atomic / column / block
template <bool X_F32>
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ x_, int64_t rows, int cols, int64_t ld,
                                                     float scale, float* __restrict__ out) {
  // block = 32 column-lanes x 8 row-lanes; blockIdx.x = column block (32 cols), blockIdx.y = row slab
  __shared__ float sm[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < cols) {
    for (int64_t r = int64_t(blockIdx.y) * 8 + rl; r < rows; r += int64_t(gridDim.y) * 8) {
      if constexpr (X_F32) s += static_cast<const float*>(x_)[r * ld + c];
      else s += __half2float(static_cast<const __half*>(x_)[r * ld + c]);
    }
  }
  sm[rl][cl] = s;
  __syncthreads();
  if (rl == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][cl];
    atomicAdd(out + c, t * scale);
  }
}

// fp16 fast path: 16-byte loads, a warp covers 256 consecutive columns of one row, 8 rows per block pass
__global__ void __launch_bounds__(256) colsum_f16_vec_kernel(const __half* __restrict__ x, int64_t rows, int cols,
                                                             int64_t ld, float scale, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sm[8][256 + 8];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < cols) {
    for (int64_t r = int64_t(blockIdx.y) * 8 + rl; r < rows; r += int64_t(gridDim.y) * 8) {
      const uint4 q = *reinterpret_cast<const uint4*>(x + r * ld + c);
      const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        s[2 * j] += f.x;
        s[2 * j + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[rl][lane * 8 + j] = s[j];
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    atomicAdd(out + cc, t * scale);
  }
}

// =================================================================================================
// softmax cross-entropy backward (mean reduction): dlogits = (prob - onehot) * (loss_scale / B)
// written as the fp16 operand [B, ldd] (columns >= S zero-filled).
__global__ void softmax_ce_bwd_kernel(const float* __restrict__ prob, const int64_t* __restrict__ labels, float coef,
                                      __half* __restrict__ dl, int S, int ldd) {
  const int b = blockIdx.x;
  const int label = int(labels[b]);
  for (int i = threadIdx.x; i < ldd; i += blockDim.x) {
    float v = 0.f;
    if (i < S) v = (prob[int64_t(b) * S + i] - (i == label ? 1.0f : 0.0f)) * coef;
    dl[int64_t(b) * ldd + i] = __float2half_rn(v);
  }
}

// mean pooling backward: dh[b, t, :] = demb[b, :] / T.  Vector form (H % 4 == 0): a thread owns four columns and walks
// the frames -- no per-element 64-bit modulo, 16-byte stores (29.6 -> ~8 us for the 29 MB of cfg1).
__global__ void mean_pool_bwd_vec_kernel(const float* __restrict__ demb, float* __restrict__ dh, int T, int H) {
  const int b = blockIdx.z;
  const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (c4 * 4 >= H) return;
  const float inv = 1.0f / float(T);
  float4 v = *reinterpret_cast<const float4*>(demb + int64_t(b) * H + c4 * 4);
  v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
  float* base = dh + int64_t(b) * T * H + c4 * 4;
  for (int t = blockIdx.y; t < T; t += gridDim.y) *reinterpret_cast<float4*>(base + int64_t(t) * H) = v;
}
__global__ void mean_pool_bwd_kernel(const float* __restrict__ demb, float* __restrict__ dh, int T, int H) {
  const int b = blockIdx.y;
  const int64_t n = int64_t(T) * H;
  const float inv = 1.0f / float(T);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    dh[int64_t(b) * n + i] = demb[int64_t(b) * H + (i % H)] * inv;
}

// =================================================================================================
// Adam (torch.optim.Adam semantics, weight_decay = 0, amsgrad = False) over a flat fp32 parameter
// buffer: p -= lr * m_hat / (sqrt(v_hat) + eps); the gradient carries the loss scale (grad_scale = 1/scale).
// Vectorised (n and the pointers 16-byte aligned: the flat buffers are) with a scalar tail.  zero_grad: the
// gradient is cleared in the same pass (it is re-accumulated by the next backward), which saves the separate
// fill over the 0.38 GB buffer.  The launch uses 3 resident blocks per SM instead of a full machine: the
// update is HBM-bound and runs on its own stream under the next step's (tensor-bound) CNN forward, whose
// 227 KB / 320-thread GEMM CTAs must still fit next to it.
__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float lr_c, float beta1, float beta2,
                                         float eps, float rbc2, float grad_scale, bool zero_grad) {
  const float gr = g * grad_scale;
  m = beta1 * m + (1.0f - beta1) * gr;
  v = beta2 * v + (1.0f - beta2) * gr * gr;
  p -= lr_c * m / (sqrtf(v) * rbc2 + eps);
  if (zero_grad) g = 0.f;
}
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float lr, float beta1, float beta2,
                                                   float eps, float bc1, float bc2, float grad_scale, int zero_grad) {
  const float lr_c = lr / bc1, rbc2 = 1.0f / sqrtf(bc2);
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, lr_c, beta1, beta2, eps, rbc2, grad_scale, zero_grad);
    adam_one(pp.y, gg.y, mm.y, vv.y, lr_c, beta1, beta2, eps, rbc2, grad_scale, zero_grad);
    adam_one(pp.z, gg.z, mm.z, vv.z, lr_c, beta1, beta2, eps, rbc2, grad_scale, zero_grad);
    adam_one(pp.w, gg.w, mm.w, vv.w, lr_c, beta1, beta2, eps, rbc2, grad_scale, zero_grad);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
    if (zero_grad) g4[i] = gg;
  }
  if (blockIdx.x == 0) {
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x)
      adam_one(p[i], g[i], m[i], v[i], lr_c, beta1, beta2, eps, rbc2, grad_scale, zero_grad);
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" {

int w2v2_cast_f16_transpose(const float* w, void* wt16, int R, int C, int ldt, const float* row_scale, void* stream) {
  W2V2_REQUIRE(R > 0 && C > 0 && ldt >= R, "w2v2_cast_f16_transpose: bad shape R=%d C=%d ldt=%d", R, C, ldt);
  dim3 grid((C + 31) / 32, (ldt + 31) / 32), block(32, 8);
  cast_f16_transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(w, (__half*)wt16, R, C, ldt, row_scale);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_layernorm_bwd(const float* dy_a, const float* dy_b, const void* xa, int xa_dtype, const float* bias,
                       const float* residual, const float* gamma, float eps, float* dx32, void* dx16, float* dgamma,
                       float* dbeta, int64_t rows, int H, void* stream) {
  return w2v2_layernorm_bwd_ex(dy_a, dy_b, xa, xa_dtype, bias, residual, gamma, eps, dx32, dx16, dgamma, dbeta, nullptr,
                               rows, H, 0.f, 0, stream);
}

int w2v2_layernorm_bwd_from_output(const float* dy_a, const float* dy_b, const float* y32, const float* rstd,
                                   const float* gamma, const float* beta, float* dx32, void* dx16, float* dgamma,
                                   float* dbeta, float* dbias, int64_t rows, int H, float drop_p, uint64_t drop_seed,
                                   void* stream) {
  W2V2_REQUIRE(H == 512 || H == 768 || H == 1024, "w2v2_layernorm_bwd_from_output: H=%d not in {512, 768, 1024}", H);
  W2V2_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "w2v2_layernorm_bwd_from_output: drop_p=%f out of [0,1)", drop_p);
  W2V2_REQUIRE(y32 != nullptr && rstd != nullptr && beta != nullptr, "w2v2_layernorm_bwd_from_output: y, rstd, beta are required");
  if (rows == 0) return 0;
  const uint32_t thr = uint32_t(drop_p * 65536.0f + 0.5f);
  const float inv_keep = 1.0f / (1.0f - float(thr) / 65536.0f);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool ring = []() { const char* e = getenv("W2V2_LNB_RING"); return !(e != nullptr && e[0] == '0'); }();
  if (ring && dy_b == nullptr) {      // (one gradient stream: what the accumulating data-gradient GEMMs leave)
    int rc;
    if (H == 512) rc = launch_lnb_ring<4>(dy_a, y32, rstd, gamma, beta, dx32, dx16, dgamma, dbeta, dbias, rows, thr, inv_keep, drop_seed, st);
    else if (H == 768) rc = launch_lnb_ring<6>(dy_a, y32, rstd, gamma, beta, dx32, dx16, dgamma, dbeta, dbias, rows, thr, inv_keep, drop_seed, st);
    else rc = launch_lnb_ring<8>(dy_a, y32, rstd, gamma, beta, dx32, dx16, dgamma, dbeta, dbias, rows, thr, inv_keep, drop_seed, st);
    if (rc) return rc;
    count_launches(1);
    W2V2_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int64_t want = (rows + LNB_WARPS - 1) / LNB_WARPS, slots = int64_t(device_sm_count()) * (H <= 768 ? 4 : 3);
  const int64_t per_warp = (want + slots - 1) / slots;
  const int grid = int((want + per_warp - 1) / per_warp);
#define W2V2_LNY(NV) \
  launch_k(layernorm_bwd_kernel<true, NV, true, true>, dim3(grid), dim3(LNB_WARPS * 32), 0, st, 1, dy_a, dy_b, (const void*)y32, beta, rstd, gamma, 0.f, dx32, (__half*)dx16, dgamma, dbeta, dbias, rows, H, thr, inv_keep, drop_seed)
  if (H == 512) W2V2_LNY(4); else if (H == 768) W2V2_LNY(6); else W2V2_LNY(8);
#undef W2V2_LNY
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_layernorm_bwd_ex(const float* dy_a, const float* dy_b, const void* xa, int xa_dtype, const float* bias,
                          const float* residual, const float* gamma, float eps, float* dx32, void* dx16, float* dgamma,
                          float* dbeta, float* dbias, int64_t rows, int H, float drop_p, uint64_t drop_seed,
                          void* stream) {
  W2V2_REQUIRE(H % 4 == 0 && H <= 1024, "w2v2_layernorm_bwd: H=%d must be a multiple of 4 and <= 1024", H);
  W2V2_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "w2v2_layernorm_bwd: drop_p=%f out of [0,1)", drop_p);
  if (rows == 0) return 0;
  const uint32_t thr = uint32_t(drop_p * 65536.0f + 0.5f);
  const float inv_keep = 1.0f / (1.0f - float(thr) / 65536.0f);
  // one wave of resident blocks, every warp the same number of rows (+-1)
  const int64_t want = (rows + LNB_WARPS - 1) / LNB_WARPS, slots = int64_t(device_sm_count()) * (H <= 768 ? 4 : 3);
  const int64_t per_warp = (want + slots - 1) / slots;
  const int grid = int((want + per_warp - 1) / per_warp);
  cudaStream_t st = (cudaStream_t)stream;
#define W2V2_LNB(F32, NV, EX) \
  launch_k(layernorm_bwd_kernel<F32, NV, EX>, dim3(grid), dim3(LNB_WARPS * 32), 0, st, 1, dy_a, dy_b, xa, bias, residual, gamma, eps, dx32, (__half*)dx16, dgamma, dbeta, dbias, rows, H, thr, inv_keep, drop_seed)
  if (xa_dtype == 1) {
    if (H == 512) W2V2_LNB(true, 4, true); else if (H == 768) W2V2_LNB(true, 6, true);
    else if (H == 1024) W2V2_LNB(true, 8, true); else W2V2_LNB(true, 8, false);
  } else {
    if (H == 512) W2V2_LNB(false, 4, true); else if (H == 768) W2V2_LNB(false, 6, true);
    else if (H == 1024) W2V2_LNB(false, 8, true); else W2V2_LNB(false, 8, false);
  }
#undef W2V2_LNB
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_gelu_bwd(const void* dg16, const void* z16, void* dz16, int64_t n, void* stream) {
  W2V2_REQUIRE(n % 8 == 0, "w2v2_gelu_bwd: n=%lld must be a multiple of 8", (long long)n);
  if (n == 0) return 0;
  gelu_bwd_kernel<<<grid_cap((n / 8 + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>((const __half*)dg16, (const __half*)z16,
                                                                                     (__half*)dz16, n);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_gelu_bwd_colsum(const void* dg16, const void* z16, void* dz16, int64_t rows, int cols, float* dbias, void* stream) {
  W2V2_REQUIRE(cols % 8 == 0, "w2v2_gelu_bwd_colsum: cols=%d must be a multiple of 8", cols);
  W2V2_REQUIRE(dbias != nullptr, "w2v2_gelu_bwd_colsum: dbias is required (use w2v2_gelu_bwd without it)");
  if (rows == 0) return 0;
  dim3 grid((cols + 255) / 256, 1);
  int64_t want = (int64_t(device_sm_count()) * 6 + grid.x - 1) / grid.x, cap = (rows + 7) / 8;
  grid.y = unsigned(want < cap ? want : cap);
  W2V2_CHECK_CUDA(launch_k(gelu_bwd_colsum_kernel, grid, dim3(256), 0, (cudaStream_t)stream, 1, (const __half*)dg16,
                           (const __half*)z16, (__half*)dz16, rows, cols, dbias));
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_colsum(const void* x, int x_dtype, int64_t rows, int cols, int64_t ld, float scale, float* out, void* stream) {
  if (rows == 0) return 0;
  if (x_dtype == 0 && cols % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    dim3 vgrid((cols + 255) / 256, (unsigned)grid_cap((rows + 63) / 64, 1));
    const unsigned want = (unsigned)(2 * device_sm_count() / vgrid.x + 1);
    if (vgrid.y > want) vgrid.y = want;
    W2V2_CHECK_CUDA(launch_k(colsum_f16_vec_kernel, vgrid, dim3(256), 0, (cudaStream_t)stream, 1, (const __half*)x, rows, cols,
                             ld, scale, out));
    count_launches(1);
    W2V2_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  dim3 grid((cols + 31) / 32, (unsigned)grid_cap((rows + 63) / 64, 1));
  if (grid.y > 64) grid.y = 64;
  if (x_dtype == 1) colsum_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, scale, out);
  else colsum_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, scale, out);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_softmax_ce_bwd(const float* prob, const int64_t* labels, float coef, void* dlogits16, int B, int S, int ldd,
                        void* stream) {
  W2V2_REQUIRE(ldd >= S, "w2v2_softmax_ce_bwd: ldd=%d < S=%d", ldd, S);
  softmax_ce_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(prob, labels, coef, (__half*)dlogits16, S, ldd);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_mean_pool_bwd(const float* demb, float* dh, int B, int T, int H, void* stream) {
  if (H % 4 == 0 && ((reinterpret_cast<uintptr_t>(demb) | reinterpret_cast<uintptr_t>(dh)) & 15) == 0) {
    const int tx = 64, ty = T < 32 ? T : 32;
    dim3 vgrid((H / 4 + tx - 1) / tx, ty, B);
    mean_pool_bwd_vec_kernel<<<vgrid, tx, 0, (cudaStream_t)stream>>>(demb, dh, T, H);
    count_launches(1);
    W2V2_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  dim3 grid((unsigned)grid_cap((int64_t(T) * H + 255) / 256, 2), B);
  mean_pool_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(demb, dh, T, H);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int w2v2_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   int step, float grad_scale, void* stream) {
  return w2v2_adam_step_ex(p, const_cast<float*>(g), m, v, n, lr, beta1, beta2, eps, step, grad_scale, 0, stream);
}

int w2v2_adam_step_ex(float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                      int step, float grad_scale, int zero_grad, void* stream) {
  W2V2_REQUIRE(step >= 1, "w2v2_adam_step: step counts from 1");
  W2V2_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0, "w2v2_adam_step: buffers must be 16-byte aligned");
  const float bc1 = 1.0f - powf(beta1, float(step));
  const float bc2 = 1.0f - powf(beta2, float(step));
  adam_kernel<<<grid_cap((n / 4 + 255) / 256, 3), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1,
                                                                                 bc2, grad_scale, zero_grad);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
