// Common device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// UMMA descriptors, math.  Hand-written PTX; bit layouts follow the PTX ISA for sm_100a
// (cross-checked against CUTLASS 4.x cute/arch/mma_sm100_desc.hpp for field positions).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace w2v2 {

// --------------------------------------------------------------------------------------------
// error plumbing (no exceptions cross the C ABI)
void set_last_error(const char* fmt, ...);
// number of kernels launched by this library since load (bench.py's `gpu_launches` evidence)
void count_launches(int n);
#define W2V2_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::w2v2::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -2;                                                                           \
    }                                                                                      \
  } while (0)
#define W2V2_REQUIRE(cond, ...)                                                            \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      ::w2v2::set_last_error(__VA_ARGS__);                                                 \
      return -1;                                                                           \
    }                                                                                      \
  } while (0)

// --------------------------------------------------------------------------------------------
// Programmatic dependent launch: the hot kernels are launched with programmatic stream serialisation, so a
// kernel's CTAs are scheduled (and run their prologue: barrier init, TMEM allocation, tensor-map prefetch)
// while the previous kernel's last CTAs are still finishing.  Contract for every kernel launched through
// launch_k(): pdl_trigger() early, and pdl_wait() before the first access to memory another kernel wrote
// or will read -- the wait returns once the whole preceding grid has completed and its writes are visible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();     // W2V2_PDL=0 turns the attribute off (plain stream order)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                            Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}\n"
      : "+r"(pred));
  return pred != 0;
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n"
      "DONE:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- TMA -----------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1,
                                             int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA store / UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- named barriers ------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- tcgen05 / TMEM ------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued MMAs of this thread retire (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp gets TMEM lane (base_lane + i), columns [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA pairs (cta_group::2): two CTAs of one cluster, on the two SMs of a TPC, execute one MMA -----
// A 256 x N tile: each CTA stages its own 128 rows of A and its own N/2 columns of B, the leader CTA
// (cluster rank 0) issues the MMA, each CTA's TMEM receives the accumulator rows of its own A half.
// Per MMA every SM then reads 8 KB of operands instead of 12 KB, which is what lifts the shared-memory
// bandwidth ceiling of the single-CTA form (operand reads + TMA fills > 128 B/clk at full tensor rate).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// TMA load of a CTA pair: data lands in this CTA's smem, completion bytes are counted on `bar_cluster_addr`
// (a shared::cluster address: the leader CTA's barrier)
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                                int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- UMMA descriptors ----------------------------------------------------------------------
// Instruction descriptor, kind::f16: fp16 A/B (format 0), fp32 accumulate (c_format 1),
// a_major/b_major: 0 = K-major, 1 = MN-major.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4)                      // c_format = F32
         | (0u << 7) | (0u << 10)       // a_format = b_format = F16
         | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16)
         | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// Shared-memory matrix descriptor.  layout_type: 0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3fff);
  d |= uint64_t((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= uint64_t(1) << 46;               // descriptor version (Blackwell)
  d |= uint64_t(layout_type & 7) << 61;
  return d;
}
// K-major operand tile stored as rows of 128 bytes (64 fp16) with the 128B swizzle (what TMA
// SWIZZLE_128B writes): 8-row atoms of 1024 B => SBO = 1024; LBO unused for swizzled K-major.
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
  return make_smem_desc(saddr, 16, 1024, 2);
}

// ---- math ----------------------------------------------------------------------------------
// ---- packed fp32x2 arithmetic (Blackwell FFMA2: two fp32 FMAs per issue slot) -----------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 splat2(float v) { return pack2(v, v); }

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exact-erf GELU (HF "gelu" == torch F.gelu default):
//     gelu(x) = x Phi(x) = max(x, 0) - |x| Phi(-|x|),      Phi(-a) = 2^Q(a)
// Q = log2 of the normal tail, a degree-8 polynomial (weighted minimax fit on a in [0, 6.5], evaluated in
// u = 2a/6.5 - 1; |x| is clamped to 6.5 where the tail term is < 5e-10).  Max abs error 2.7e-7 in fp32
// (relative 4e-6 where |gelu| > 1e-2) -- three orders below the fp16 rounding every consumer applies.
// One MUFU.EX2 and, in the packed form, 8 issue slots per element (vs ~35 for erff()).
#define W2V2_GELU_C0 -10.75906753540039f
#define W2V2_GELU_C1 -16.487607955932617f
#define W2V2_GELU_C2 -7.1384711265563965f
#define W2V2_GELU_C3 -0.22706320881843567f
#define W2V2_GELU_C4 0.10637267678976059f
#define W2V2_GELU_C5 -0.03946828097105026f
#define W2V2_GELU_C6 0.02791619300842285f
#define W2V2_GELU_C7 -0.013774366118013859f
#define W2V2_GELU_C8 -0.004659503698348999f
#define W2V2_GELU_A 6.5f

__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fabsf(x);
  const float u = fmaf(fminf(a, W2V2_GELU_A), 2.0f / W2V2_GELU_A, -1.0f);
  float q = fmaf(W2V2_GELU_C8, u, W2V2_GELU_C7);
  q = fmaf(q, u, W2V2_GELU_C6);
  q = fmaf(q, u, W2V2_GELU_C5);
  q = fmaf(q, u, W2V2_GELU_C4);
  q = fmaf(q, u, W2V2_GELU_C3);
  q = fmaf(q, u, W2V2_GELU_C2);
  q = fmaf(q, u, W2V2_GELU_C1);
  q = fmaf(q, u, W2V2_GELU_C0);
  return fmaf(-a, fast_ex2(q), fmaxf(x, 0.f));
}

// two elements at once on the packed fp32x2 pipe
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  const f32x2 u = fma2(pack2(fminf(a0, W2V2_GELU_A), fminf(a1, W2V2_GELU_A)), splat2(2.0f / W2V2_GELU_A), splat2(-1.0f));
  f32x2 q = fma2(splat2(W2V2_GELU_C8), u, splat2(W2V2_GELU_C7));
  q = fma2(q, u, splat2(W2V2_GELU_C6));
  q = fma2(q, u, splat2(W2V2_GELU_C5));
  q = fma2(q, u, splat2(W2V2_GELU_C4));
  q = fma2(q, u, splat2(W2V2_GELU_C3));
  q = fma2(q, u, splat2(W2V2_GELU_C2));
  q = fma2(q, u, splat2(W2V2_GELU_C1));
  q = fma2(q, u, splat2(W2V2_GELU_C0));
  float q0, q1;
  unpack2(q, q0, q1);
  const f32x2 r = fma2(pack2(-a0, -a1), pack2(fast_ex2(q0), fast_ex2(q1)), pack2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  unpack2(r, x0, x1);
}

// gelu'(z) = Phi(z) + z phi(z) for two elements, the same log2-tail polynomial on the packed pipe; the value the
// GEMM epilogue of the FFN data gradient multiplies by (two MUFU.EX2 and ~9 issue slots per element).
__device__ __forceinline__ void gelu_grad2(float z0, float z1, float& d0, float& d1) {
  const float a0 = fabsf(z0), a1 = fabsf(z1);
  const f32x2 u = fma2(pack2(fminf(a0, W2V2_GELU_A), fminf(a1, W2V2_GELU_A)), splat2(2.0f / W2V2_GELU_A), splat2(-1.0f));
  f32x2 q = fma2(splat2(W2V2_GELU_C8), u, splat2(W2V2_GELU_C7));
  q = fma2(q, u, splat2(W2V2_GELU_C6));
  q = fma2(q, u, splat2(W2V2_GELU_C5));
  q = fma2(q, u, splat2(W2V2_GELU_C4));
  q = fma2(q, u, splat2(W2V2_GELU_C3));
  q = fma2(q, u, splat2(W2V2_GELU_C2));
  q = fma2(q, u, splat2(W2V2_GELU_C1));
  q = fma2(q, u, splat2(W2V2_GELU_C0));
  float q0, q1;
  unpack2(q, q0, q1);
  const float t0 = fast_ex2(q0), t1 = fast_ex2(q1);                  // Phi(-|z|)
  const float c0 = z0 >= 0.f ? 1.0f - t0 : t0, c1 = z1 >= 0.f ? 1.0f - t1 : t1;
  const float p0 = 0.3989422804014327f * fast_ex2(z0 * z0 * -0.72134752044448170f);
  const float p1 = 0.3989422804014327f * fast_ex2(z1 * z1 * -0.72134752044448170f);
  d0 = fmaf(z0, p0, c0);
  d1 = fmaf(z1, p1, c1);
}

// GELU and its derivative from ONE evaluation of the tail polynomial (the training forward of FFN1 can keep
// gelu'(z) instead of z: the backward epilogue then only multiplies).  x0/x1: in = z, out = gelu(z).
__device__ __forceinline__ void gelu_and_grad2(float& x0, float& x1, float& d0, float& d1) {
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  const f32x2 u = fma2(pack2(fminf(a0, W2V2_GELU_A), fminf(a1, W2V2_GELU_A)), splat2(2.0f / W2V2_GELU_A), splat2(-1.0f));
  f32x2 q = fma2(splat2(W2V2_GELU_C8), u, splat2(W2V2_GELU_C7));
  q = fma2(q, u, splat2(W2V2_GELU_C6));
  q = fma2(q, u, splat2(W2V2_GELU_C5));
  q = fma2(q, u, splat2(W2V2_GELU_C4));
  q = fma2(q, u, splat2(W2V2_GELU_C3));
  q = fma2(q, u, splat2(W2V2_GELU_C2));
  q = fma2(q, u, splat2(W2V2_GELU_C1));
  q = fma2(q, u, splat2(W2V2_GELU_C0));
  float q0, q1;
  unpack2(q, q0, q1);
  const float t0 = fast_ex2(q0), t1 = fast_ex2(q1);                  // Phi(-|z|)
  const float p0 = 0.3989422804014327f * fast_ex2(x0 * x0 * -0.72134752044448170f);
  const float p1 = 0.3989422804014327f * fast_ex2(x1 * x1 * -0.72134752044448170f);
  d0 = fmaf(x0, p0, x0 >= 0.f ? 1.0f - t0 : t0);
  d1 = fmaf(x1, p1, x1 >= 0.f ? 1.0f - t1 : t1);
  x0 = fmaf(-a0, t0, fmaxf(x0, 0.f));
  x1 = fmaf(-a1, t1, fmaxf(x1, 0.f));
}

// counter-based random bits for dropout: a 32-bit multiply / xor-shift mix (murmur3 finaliser) of the pair
// index with both seed halves folded in; low / high 16 bits decide the two elements of the pair.
// (32-bit on purpose: a 64-bit mix costs ~3x the instructions, and the attention kernels run it on a
// one-thread-per-row critical path.)
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
// The two 32-bit keys derived from the 64-bit seed.  Kernels derive them ONCE per thread (drop_keys at
// kernel entry): left inside the per-element call the compiler re-derives them for every pair.
struct DropKeys {
  uint32_t k0, k1;
};
__host__ __device__ __forceinline__ DropKeys drop_keys(uint64_t seed) {
  // nearby seeds -- the per-layer / per-site offsets of one step -- give unrelated masks
  DropKeys k;
  k.k0 = fmix32(uint32_t(seed) + 0x9E3779B9u);
  k.k1 = fmix32(uint32_t(seed >> 32) ^ k.k0 ^ 0x7F4A7C15u);
  return k;
}
// pair indices below 2^32 (the attention kernels check this on the host): 3 multiplies, 3 xor-shifts, 1 add
__host__ __device__ __forceinline__ uint32_t dropout_hash32(DropKeys k, uint32_t pair_idx) {
  uint32_t x = (pair_idx ^ k.k0) * 0x9E3779B1u;
  x ^= x >> 15;
  x += k.k1;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t dropout_hash_k(DropKeys k, uint64_t pair_idx) {
  uint32_t x = (uint32_t(pair_idx) ^ k.k0) * 0x9E3779B1u;
  x ^= x >> 15;
  x += k.k1 + uint32_t(pair_idx >> 32) * 0x7FEB352Du;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t dropout_hash(uint64_t seed, uint64_t pair_idx) {
  return dropout_hash_k(drop_keys(seed), pair_idx);
}

// inverted dropout of four consecutive elements starting at the (even) flat element index `idx`
__device__ __forceinline__ void dropout4(float4& a, DropKeys k, int64_t idx, uint32_t thr, float inv_keep) {
  const uint32_t h0 = dropout_hash_k(k, uint64_t(idx >> 1));
  const uint32_t h1 = dropout_hash_k(k, uint64_t(idx >> 1) + 1);
  a.x = (h0 & 0xffffu) >= thr ? a.x * inv_keep : 0.f;
  a.y = (h0 >> 16) >= thr ? a.y * inv_keep : 0.f;
  a.z = (h1 & 0xffffu) >= thr ? a.z * inv_keep : 0.f;
  a.w = (h1 >> 16) >= thr ? a.w * inv_keep : 0.f;
}
__device__ __forceinline__ void dropout4(float4& a, uint64_t seed, int64_t idx, uint32_t thr, float inv_keep) {
  dropout4(a, drop_keys(seed), idx, thr, inv_keep);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Column sums over the 32 rows a warp holds (lane = row, v[j] = that row's j-th column): lane j adds the sum of
// column j to dst[j].  Used by the attention backward to emit the q/k/v bias gradients while draining TMEM.
template <int N>
__device__ __forceinline__ void warp_colsum_atomic(const float (&v)[N], bool row_valid, float scale, float* dst) {
  static_assert(N <= 32, "one lane per column");
  const int lane = threadIdx.x & 31;
  float mine = 0.f;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float s = warp_sum(row_valid ? v[j] : 0.f);
    if (lane == j) mine = s;
  }
  if (lane < N) atomicAdd(dst + lane, mine * scale);
}

// lane = row, v[j] = column j of that row (N = 16 or 32).  Returns, in lane L, the sum of column (L mod N) over the
// warp's 32 rows: at each step a lane keeps one half of its columns and trades the other half with its partner.
template <int N>
__device__ __forceinline__ float warp_colsum_bfly(float (&v)[N]) {
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int h = N / 2; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float keep = up ? v[i + h] : v[i], send = up ? v[i] : v[i + h];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  float r = v[0];
  if (N == 16) r += __shfl_xor_sync(0xffffffffu, r, 16);
  return r;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace w2v2
