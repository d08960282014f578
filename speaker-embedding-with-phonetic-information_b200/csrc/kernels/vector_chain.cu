// Post-step on the embeddings: what the reference's back-end does to every
// x-vector / c-vector right after extraction, as three more processes on a pipe
//   ivector-subtract-global-mean mean.vec scp:xvector.scp ark:- |
//   transform-vec transform.mat ark:- ark:- | ivector-normalize-length ark:- ark:- |
// (egs/sre/v5/run_sre10.sh:118-129, egs/sre/v2/run_sre16.sh; upstream programs restated
// [UPSTREAM-RECALL ivectorbin/ivector-subtract-global-mean.cc, bin/transform-vec.cc,
//  ivectorbin/ivector-normalize-length.cc]):
//   x  = v - mean
//   y  = T x                     (T [rows, dim])   or   T [x; 1]   (T [rows, dim + 1])
//   y *= sqrt(rows) / ||y||      (--normalize=true --scaleup=true defaults; scaleup=false: 1 / ||y||)
// One CTA per vector; the 150 x 512 transform (300 KB) is read from L2 by every CTA.  The work
// is tiny (150 kFLOP per vector): the point is to leave the vectors on the device between the
// utterance average and the D2H copy instead of three more process hops.
#include "kernels.h"

namespace cvx {
namespace {

__global__ void __launch_bounds__(256) vector_chain_kernel(const float* __restrict__ in, int dim,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ tf, int t_rows, int t_cols,
                                                            int normalize, int scaleup, int keep_zero,
                                                            float* __restrict__ out, int out_dim,
                                                            float* __restrict__ ratios) {
  extern __shared__ float sm[];
  float* x = sm;        // [dim]
  float* y = sm + dim;  // [out_dim]
  __shared__ double red[8];
  const int64_t v = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* src = in + v * dim;
  int nonzero = 0;
  for (int c = tid; c < dim; c += 256) {
    const float s = src[c];
    nonzero |= s != 0.f;
    x[c] = s - (mean ? mean[c] : 0.f);
  }
  // keep_zero: an all-zero input is the extraction path's "utterance failed" marker and stays zero
  if (__syncthreads_or(nonzero) == 0 && keep_zero) {
    for (int c = tid; c < out_dim; c += 256) out[v * out_dim + c] = 0.f;
    if (ratios && tid == 0) ratios[v] = 0.f;
    return;
  }
  if (tf) {
    for (int r = warp; r < t_rows; r += 8) {
      const float* row = tf + static_cast<size_t>(r) * t_cols;
      float acc = 0.f;
      for (int c = lane; c < dim; c += 32) acc = fmaf(row[c], x[c], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) y[r] = acc + (t_cols == dim + 1 ? row[dim] : 0.f);
    }
  } else {
    for (int c = tid; c < dim; c += 256) y[c] = x[c];
  }
  __syncthreads();
  double ss = 0.0;
  for (int c = tid; c < out_dim; c += 256) ss += static_cast<double>(y[c]) * y[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  // ratio as ivector-normalize-length reports it: ||y|| / sqrt(dim) with scaleup, ||y|| without
  const float ratio = static_cast<float>(sqrt(tot) / (scaleup ? sqrt(static_cast<double>(out_dim)) : 1.0));
  if (ratios && tid == 0) ratios[v] = ratio;
  const float scale = (normalize && ratio != 0.f) ? 1.f / ratio : 1.f;
  float* dst = out + v * out_dim;
  for (int c = tid; c < out_dim; c += 256) dst[c] = y[c] * scale;
}

}  // namespace

cudaError_t LaunchVectorChain(const float* in, int64_t num_vectors, int32_t dim, const float* mean, const float* transform,
                              int32_t t_rows, int32_t t_cols, int32_t normalize, int32_t scaleup, int32_t keep_zero,
                              float* out, float* ratios, cudaStream_t stream) {
  if (num_vectors <= 0) return cudaSuccess;
  const int32_t out_dim = transform ? t_rows : dim;
  const size_t smem = static_cast<size_t>(dim + out_dim) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(vector_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  vector_chain_kernel<<<static_cast<unsigned>(num_vectors), 256, smem, stream>>>(in, dim, mean, transform, t_rows, t_cols,
                                                                               normalize, scaleup, keep_zero, out, out_dim,
                                                                               ratios);
  return cudaGetLastError();
}

}  // namespace cvx
