// Trial scoring of the evaluation path (R:src/evaluation/speaker/cosine_distance.py:107-132, 249-262): cosine
// similarity of listed embedding pairs, with the evaluator's optional centring ((x - mean) / (std + 1e-12),
// R:src/evaluation/speaker/speaker_recognition_evaluator.py:162-167).  Length normalisation
// (speaker_recognition_evaluator.py:170-172) does not change a cosine and is therefore not a separate step.
// One warp per trial; the embedding table [N, E] stays in HBM / L2 and is gathered by index.
#include "common.cuh"
#include "w2v2_b200.h"

namespace w2v2 {

__global__ void __launch_bounds__(256) cosine_pairs_kernel(const float* __restrict__ emb, const float* __restrict__ mean,
                                                           const float* __restrict__ stdv, const int32_t* __restrict__ ia,
                                                           const int32_t* __restrict__ ib, float* __restrict__ scores,
                                                           int64_t P, int E) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t p = warp; p < P; p += nwarps) {
    const float* a = emb + int64_t(ia[p]) * E;
    const float* b = emb + int64_t(ib[p]) * E;
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int e = lane; e < E; e += 32) {
      float x = a[e], y = b[e];
      if (mean != nullptr) {
        const float m = mean[e], inv = 1.0f / (stdv[e] + 1e-12f);
        x = (x - m) * inv;
        y = (y - m) * inv;
      }
      dot = fmaf(x, y, dot);
      na = fmaf(x, x, na);
      nb = fmaf(y, y, nb);
    }
    dot = warp_sum(dot);
    na = warp_sum(na);
    nb = warp_sum(nb);
    // torch.nn.CosineSimilarity (eps = 1e-8): x.y / (max(|x|, eps) max(|y|, eps))
    if (lane == 0) scores[p] = dot / (fmaxf(sqrtf(na), 1e-8f) * fmaxf(sqrtf(nb), 1e-8f));
  }
}

int device_sm_count();

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_cosine_pairs(const float* emb, const float* mean, const float* stdv, const int32_t* idx_a,
                                 const int32_t* idx_b, float* scores, int64_t P, int E, void* stream) {
  W2V2_REQUIRE(P >= 0 && E >= 1, "w2v2_cosine_pairs: bad sizes");
  W2V2_REQUIRE((mean == nullptr) == (stdv == nullptr), "w2v2_cosine_pairs: mean and std come together");
  if (P == 0) return 0;
  int64_t blocks = (P + 7) / 8;
  const int64_t cap = int64_t(device_sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  cosine_pairs_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(emb, mean, stdv, idx_a, idx_b, scores, P, E);
  count_launches(1);
  W2V2_CHECK_CUDA(cudaGetLastError());
  return 0;
}
