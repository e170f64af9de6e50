// Gradient all-reduce through the NVSwitch (NVLink SHARP / "NVLS" multicast), the data-parallel exchange step of the
// training path (R:config/trainer/trainer.yaml:6-9 `accelerator: ddp`: a sum over ranks of every gradient element).
//
// The flat fp32 gradient lives in symmetric memory that is also mapped as ONE multicast object over all ranks.  A
// two-shot all-reduce then needs no peer-to-peer reduction code at all: rank r owns the r-th slice of a span, pulls the
// switch-reduced sum of that slice with `multimem.ld_reduce` (the switch reads the eight replicas and adds them in fp32)
// and writes it back to all replicas with one `multimem.st`.  Per GPU and step that is 1/N of the gradient in and out
// of the SM instead of 2 (N-1)/N with a ring -- and, what matters more here, the kernel is tiny: a few hundred threads,
// ~24 registers, no shared memory.  The tensor-core GEMMs of the backward need a whole SM's shared memory (224 KB) per
// CTA, so an NCCL channel CTA resident on an SM keeps a GEMM CTA off it for the length of the collective (measured:
// the summed GEMM time of a step grows 6.2 -> 6.9 ms at 8 GPUs); this kernel fits NEXT TO a GEMM CTA.
// Ordering across ranks (every replica complete before the pull, every broadcast landed before anyone reads) is done by
// the caller with symmetric-memory barriers on the same stream (trainer.py).
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

__device__ __forceinline__ float4 multimem_ld_reduce_add_f32x4(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f32x4(float4* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// mc: multicast address of element 0 of the span (16-byte aligned); n4: float4 elements of the span.
constexpr int NVLS_UNROLL = 4;
__global__ void __launch_bounds__(256, 5) nvls_allreduce_kernel(float4* __restrict__ mc, int64_t n4, int rank, int world) {
  const int64_t per = (n4 + world - 1) / world;
  const int64_t lo = per * rank, hi = (lo + per < n4) ? lo + per : n4;
  const int64_t nth = int64_t(gridDim.x) * blockDim.x;
  int64_t i = lo + int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (NVLS_UNROLL - 1) * nth < hi; i += NVLS_UNROLL * nth) {
    float4 v[NVLS_UNROLL];
#pragma unroll
    for (int u = 0; u < NVLS_UNROLL; ++u) v[u] = multimem_ld_reduce_add_f32x4(mc + i + u * nth);
#pragma unroll
    for (int u = 0; u < NVLS_UNROLL; ++u) multimem_st_f32x4(mc + i + u * nth, v[u]);
  }
  for (; i < hi; i += nth) multimem_st_f32x4(mc + i, multimem_ld_reduce_add_f32x4(mc + i));
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_nvls_allreduce_f32(void* multicast_base, int64_t lo, int64_t hi, int rank, int world, int ctas,
                                       void* stream) {
  W2V2_REQUIRE(multicast_base != nullptr, "w2v2_nvls_allreduce_f32: no multicast mapping (NVLS unavailable)");
  W2V2_REQUIRE(lo >= 0 && hi >= lo && lo % 4 == 0 && hi % 4 == 0, "w2v2_nvls_allreduce_f32: span [%lld, %lld) must be 4-element aligned",
               (long long)lo, (long long)hi);
  W2V2_REQUIRE(world >= 1 && rank >= 0 && rank < world, "w2v2_nvls_allreduce_f32: bad rank %d of %d", rank, world);
  W2V2_REQUIRE((reinterpret_cast<uintptr_t>(multicast_base) & 15) == 0, "w2v2_nvls_allreduce_f32: base must be 16-byte aligned");
  if (hi == lo) return 0;
  if (ctas < 1) ctas = 128;
  float4* mc = reinterpret_cast<float4*>(static_cast<float*>(multicast_base) + lo);
  nvls_allreduce_kernel<<<ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(mc, (hi - lo) / 4, rank, world);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}
