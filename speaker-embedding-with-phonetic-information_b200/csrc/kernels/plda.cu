// PLDA scoring (last step of the back-end): what `ivector-plda-scoring` computes for
// every trial of egs/sre/v5/run_sre10.sh:122-129.  Rules restated from upstream
// [UPSTREAM-RECALL ivector/plda.cc: Plda::TransformIvector, GetNormalizationFactor,
// LogLikelihoodRatio]; double precision like upstream (150-dim vectors: the work is tiny, the point
// is that enrolment / test vectors and the trial list never leave the device between the steps).
//
//   plda_transform_kernel  y = T (x - mean)  (upstream: offset + T x with offset = -T mean), then
//                          y *= sqrt(dim / sum_d y_d^2 / (psi_d + 1/n))   (n = utterances behind x)
//                          or, --simple-length-normalization, y *= sqrt(dim) / ||y||
//   plda_score_kernel      one warp per trial (enrolled speaker i with n_i utterances, test j):
//                          m_d = n psi_d / (n psi_d + 1) * train_d,  v_d = 1 + psi_d / (n psi_d + 1)
//                          llr = -0.5 [ sum_d log v_d + (test_d - m_d)^2 / v_d ]
//                                +0.5 [ sum_d log(psi_d + 1) + test_d^2 / (psi_d + 1) ]
#include "kernels.h"

namespace cvx {
namespace {

__global__ void __launch_bounds__(256) plda_transform_kernel(const float* __restrict__ in, int dim_in,
                                                              const double* __restrict__ mean,
                                                              const double* __restrict__ tf, int dim,
                                                              const double* __restrict__ psi,
                                                              const int32_t* __restrict__ num_examples,
                                                              int normalize_length, int simple_norm,
                                                              double* __restrict__ out) {
  extern __shared__ double smd[];
  double* x = smd;           // [dim_in]
  double* y = smd + dim_in;  // [dim]
  __shared__ double red[8];
  const int64_t v = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = tid; c < dim_in; c += 256) x[c] = static_cast<double>(in[v * dim_in + c]) - mean[c];
  __syncthreads();
  for (int r = warp; r < dim; r += 8) {
    const double* row = tf + static_cast<size_t>(r) * dim_in;
    double acc = 0.0;
    for (int c = lane; c < dim_in; c += 32) acc += row[c] * x[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[r] = acc;
  }
  __syncthreads();
  const double n = num_examples ? static_cast<double>(num_examples[v]) : 1.0;
  double s = 0.0;
  for (int c = tid; c < dim; c += 256) s += simple_norm ? y[c] * y[c] : y[c] * y[c] / (psi[c] + 1.0 / n);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const double factor = normalize_length ? sqrt(static_cast<double>(dim) / tot) : 1.0;
  for (int c = tid; c < dim; c += 256) out[v * dim + c] = y[c] * factor;
}

__global__ void __launch_bounds__(256) plda_score_kernel(const double* __restrict__ train,
                                                          const int32_t* __restrict__ train_n,
                                                          const double* __restrict__ test,
                                                          const double* __restrict__ psi, int dim,
                                                          const int32_t* __restrict__ trial_train,
                                                          const int32_t* __restrict__ trial_test, int64_t num_trials,
                                                          float* __restrict__ scores) {
  const int64_t t = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= num_trials) return;
  const int i = trial_train[t], j = trial_test[t];
  const double n = train_n ? static_cast<double>(train_n[i]) : 1.0;
  const double* a = train + static_cast<size_t>(i) * dim;
  const double* b = test + static_cast<size_t>(j) * dim;
  double given = 0.0, without = 0.0;
  for (int d = lane; d < dim; d += 32) {
    const double p = psi[d], np1 = n * p + 1.0;
    const double m = n * p / np1 * a[d], var = 1.0 + p / np1;
    const double diff = b[d] - m;
    given += log(var) + diff * diff / var;
    without += log(p + 1.0) + b[d] * b[d] / (p + 1.0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    given += __shfl_xor_sync(0xffffffffu, given, o);
    without += __shfl_xor_sync(0xffffffffu, without, o);
  }
  if (lane == 0) scores[t] = static_cast<float>(-0.5 * (given - without));  // the 2 pi terms cancel
}

}  // namespace

cudaError_t LaunchPldaTransform(const float* in, int64_t num_vectors, int32_t dim_in, const double* mean, const double* transform,
                                int32_t dim, const double* psi, const int32_t* num_examples, int32_t normalize_length,
                                int32_t simple_norm, double* out, cudaStream_t stream) {
  if (num_vectors <= 0) return cudaSuccess;
  const size_t smem = static_cast<size_t>(dim_in + dim) * sizeof(double);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(plda_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  plda_transform_kernel<<<static_cast<unsigned>(num_vectors), 256, smem, stream>>>(in, dim_in, mean, transform, dim, psi,
                                                                                 num_examples, normalize_length, simple_norm, out);
  return cudaGetLastError();
}

cudaError_t LaunchPldaScore(const double* train, const int32_t* train_n, const double* test, const double* psi, int32_t dim,
                            const int32_t* trial_train, const int32_t* trial_test, int64_t num_trials, float* scores,
                            cudaStream_t stream) {
  if (num_trials <= 0) return cudaSuccess;
  const int64_t blocks = (num_trials * 32 + 255) / 256;
  plda_score_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(train, train_n, test, psi, dim, trial_train, trial_test,
                                                                     num_trials, scores);
  return cudaGetLastError();
}

}  // namespace cvx
