// Frame-level outputs (nnet3-compute): the two small kernels either side of the
// frame pass when the result is a layer's activations at every frame instead of a pooled vector
// (egs/sre/v5/sid/nnet3_cvector/cvector/extract_am_embedding.sh:79-85 bottleneck features,
//  egs/sre/v5/sid/nnet3_cvector/cvector/extract_log_post.sh:80-85 log-posteriors).
//
//   pad_edges_kernel      compact rows -> padded rows: every utterance gets `left` copies of its
//                         first frame in front and `right` copies of its last frame behind, which
//                         is how upstream's DecodableNnetSimple supplies frames beyond the ends
//                         of the input [UPSTREAM-RECALL nnet3/nnet-am-decodable-simple.cc].
//   frame_output_kernel   padded fp32 activations of the output layer -> compact [rows, dim],
//                         optional row-wise log-softmax (LogSoftmaxComponent) and exp
//                         (nnet3-compute --apply-exp).
// Both are one warp per row, HBM-bound, and tiny next to the GEMMs.
#include <cfloat>

#include "kernels.h"

namespace cvx {
namespace {

// largest u in [0, n) with start[u] + u * pad <= row  (start is non-decreasing, start[0] + 0 <= row)
__device__ __forceinline__ int find_utt(const int64_t* __restrict__ start, int n, int64_t pad, int64_t row) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (start[mid] + mid * pad <= row) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) pad_edges_kernel(const float* __restrict__ src, int dim,
                                                         const int64_t* __restrict__ utt_start, int num_utts,
                                                         int left, int right, int64_t padded_rows,
                                                         float* __restrict__ dst) {
  const int64_t p = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= padded_rows) return;
  const int64_t pad = left + right;
  const int u = find_utt(utt_start, num_utts, pad, p);
  const int64_t r0 = utt_start[u], len = utt_start[u + 1] - r0;
  int64_t t = p - (r0 + u * pad) - left;
  t = t < 0 ? 0 : (t >= len ? len - 1 : t);
  const float* s = src + (r0 + t) * dim;
  float* d = dst + p * dim;
  for (int c = lane; c < dim; c += 32) d[c] = s[c];
}

__global__ void __launch_bounds__(256) frame_output_kernel(const float* __restrict__ act, int ld, int dim,
                                                            const int64_t* __restrict__ utt_start, int num_utts,
                                                            int left, int right, int64_t rows, int log_softmax,
                                                            int apply_exp, float* __restrict__ out) {
  const int64_t j = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= rows) return;
  // utterance of compact row j: largest u with utt_start[u] <= j that is not empty
  int lo = 0, hi = num_utts - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (utt_start[mid] <= j) lo = mid;
    else hi = mid - 1;
  }
  const int64_t pad = left + right;
  const float* x = act + (j + lo * pad + left) * ld;
  float* y = out + j * dim;
  float shift = 0.f;
  if (log_softmax) {
    float mx = -FLT_MAX;
    for (int c = lane; c < dim; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int c = lane; c < dim; c += 32) sum += expf(x[c] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    shift = mx + logf(sum);
  }
  for (int c = lane; c < dim; c += 32) {
    const float v = x[c] - shift;
    y[c] = apply_exp ? expf(v) : v;
  }
}

}  // namespace

cudaError_t LaunchPadEdges(const float* src, int32_t dim, const int64_t* utt_start, int32_t num_utts, int32_t left,
                           int32_t right, int64_t padded_rows, float* dst, cudaStream_t stream) {
  if (padded_rows <= 0) return cudaSuccess;
  const int64_t blocks = (padded_rows * 32 + 255) / 256;
  pad_edges_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, dim, utt_start, num_utts, left, right,
                                                                     padded_rows, dst);
  return cudaGetLastError();
}

cudaError_t LaunchFrameOutput(const float* act, int32_t ld, int32_t dim, const int64_t* utt_start, int32_t num_utts,
                              int32_t left, int32_t right, int64_t rows, int32_t log_softmax, int32_t apply_exp,
                              float* out, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  const int64_t blocks = (rows * 32 + 255) / 256;
  frame_output_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(act, ld, dim, utt_start, num_utts, left, right,
                                                                        rows, log_softmax, apply_exp, out);
  return cudaGetLastError();
}

}  // namespace cvx
