// Embedding post-step (vector chain), PLDA scoring and the stand-alone pooling test entry.
#include <algorithm>
#include <cmath>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "engine.h"
#include "engine_detail.h"

namespace cvx {

using detail::ConvertToElem;
using detail::ElemToFloat;
using detail::Encode2D;
using detail::RoundUp;

// ---------------------------------------------------------- post-step ----
void ValidateVectorChain(const cvx_vector_chain& c, int32_t dim) {
  if (c.transform) {
    if (c.transform_rows <= 0) throw InvalidArgument("vector chain: transform_rows <= 0");
    if (c.transform_cols != dim && c.transform_cols != dim + 1)
      throw InvalidArgument("vector chain: the transform has " + std::to_string(c.transform_cols) + " columns, the vectors have dim " +
                            std::to_string(dim) + " (expected dim or dim + 1)");
  }
  if (static_cast<size_t>(dim + (c.transform ? c.transform_rows : dim)) * 4 > 200 * 1024)
    throw InvalidArgument("vector chain: dimensions too large");
}

void Engine::SetVectorChain(const cvx_vector_chain* chain) {
  CVX_CUDA_CHECK(cudaSetDevice(opts_.device));
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));
  if (plan_.FrameOutput()) throw InvalidArgument("the vector chain applies to extraction engines only");
  cudaFree(d_chain_mean_);
  cudaFree(d_chain_tf_);
  d_chain_mean_ = d_chain_tf_ = nullptr;
  chain_on_ = false;
  if (!chain) return;
  const int32_t E = plan_.EmbeddingDim();
  ValidateVectorChain(*chain, E);
  chain_ = *chain;
  if (chain->mean) {
    d_chain_mean_ = DevAlloc<float>(E);
    CVX_CUDA_CHECK(cudaMemcpy(d_chain_mean_, chain->mean, static_cast<size_t>(E) * 4, cudaMemcpyHostToDevice));
  }
  if (chain->transform) {
    const size_t n = static_cast<size_t>(chain->transform_rows) * chain->transform_cols;
    d_chain_tf_ = DevAlloc<float>(n);
    CVX_CUDA_CHECK(cudaMemcpy(d_chain_tf_, chain->transform, n * 4, cudaMemcpyHostToDevice));
  }
  chain_.mean = chain_.transform = nullptr;  // host pointers are not kept
  chain_out_dim_ = chain->transform ? chain->transform_rows : E;
  chain_on_ = true;
}

void TransformVectors(int32_t device, const float* in, int64_t n, int32_t dim, const cvx_vector_chain& c, float* out,
                      float* ratios) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw CudaError("no usable CUDA device; this path has no CPU fallback");
  if (device < 0 || device >= ndev) throw InvalidArgument("device ordinal out of range");
  if (n < 0 || dim <= 0) throw InvalidArgument("bad vector count / dim");
  if (n == 0) return;
  ValidateVectorChain(c, dim);
  CVX_CUDA_CHECK(cudaSetDevice(device));
  const int32_t od = c.transform ? c.transform_rows : dim;
  float *d_in = nullptr, *d_out = nullptr, *d_mean = nullptr, *d_tf = nullptr, *d_ratio = nullptr;
  auto cleanup = [&] {
    cudaFree(d_in);
    cudaFree(d_out);
    cudaFree(d_mean);
    cudaFree(d_tf);
    cudaFree(d_ratio);
  };
  try {
    CVX_CUDA_CHECK(cudaMalloc(&d_in, static_cast<size_t>(n) * dim * 4));
    CVX_CUDA_CHECK(cudaMalloc(&d_out, static_cast<size_t>(n) * od * 4));
    CVX_CUDA_CHECK(cudaMemcpy(d_in, in, static_cast<size_t>(n) * dim * 4, cudaMemcpyHostToDevice));
    if (c.mean) {
      CVX_CUDA_CHECK(cudaMalloc(&d_mean, static_cast<size_t>(dim) * 4));
      CVX_CUDA_CHECK(cudaMemcpy(d_mean, c.mean, static_cast<size_t>(dim) * 4, cudaMemcpyHostToDevice));
    }
    if (c.transform) {
      const size_t sz = static_cast<size_t>(c.transform_rows) * c.transform_cols * 4;
      CVX_CUDA_CHECK(cudaMalloc(&d_tf, sz));
      CVX_CUDA_CHECK(cudaMemcpy(d_tf, c.transform, sz, cudaMemcpyHostToDevice));
    }
    if (ratios) CVX_CUDA_CHECK(cudaMalloc(&d_ratio, static_cast<size_t>(n) * 4));
    CVX_CUDA_CHECK(LaunchVectorChain(d_in, n, dim, d_mean, d_tf, c.transform_rows, c.transform_cols, c.normalize_length,
                                     c.scaleup, /*keep_zero=*/0, d_out, d_ratio, nullptr));
    CVX_CUDA_CHECK(cudaMemcpy(out, d_out, static_cast<size_t>(n) * od * 4, cudaMemcpyDeviceToHost));
    if (ratios) CVX_CUDA_CHECK(cudaMemcpy(ratios, d_ratio, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost));
  } catch (...) {
    cleanup();
    throw;
  }
  cleanup();
}

namespace {
// RAII for the few device buffers of the stand-alone entry points
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  template <typename T>
  T* Upload(const T* host, size_t n) {
    CVX_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    if (host && n) CVX_CUDA_CHECK(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
    return static_cast<T*>(p);
  }
};
}  // namespace

void PldaScore(int32_t device, const cvx_plda& plda, const float* train, const int32_t* train_num_utts, int32_t num_train,
               const float* test, int32_t num_test, const int32_t* trial_train, const int32_t* trial_test, int64_t num_trials,
               float* scores) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw CudaError("no usable CUDA device; this path has no CPU fallback");
  if (device < 0 || device >= ndev) throw InvalidArgument("device ordinal out of range");
  if (plda.dim_in <= 0 || plda.dim <= 0 || num_train < 0 || num_test < 0 || num_trials < 0) throw InvalidArgument("bad PLDA sizes");
  if (num_trials == 0) return;
  for (int64_t t = 0; t < num_trials; ++t)
    if (trial_train[t] < 0 || trial_train[t] >= num_train || trial_test[t] < 0 || trial_test[t] >= num_test)
      throw InvalidArgument("trial " + std::to_string(t) + " refers to a vector that does not exist");
  for (int32_t i = 0; train_num_utts && i < num_train; ++i)
    if (train_num_utts[i] <= 0) throw InvalidArgument("train_num_utts must be positive");
  CVX_CUDA_CHECK(cudaSetDevice(device));
  const size_t di = plda.dim_in, d = plda.dim;
  DevBuf b_mean, b_tf, b_psi, b_train, b_test, b_n, b_ytrain, b_ytest, b_ti, b_tj, b_scores;
  const double* d_mean = b_mean.Upload(plda.mean, di);
  const double* d_tf = b_tf.Upload(plda.transform, d * di);
  const double* d_psi = b_psi.Upload(plda.psi, d);
  const float* d_train = b_train.Upload(train, static_cast<size_t>(num_train) * di);
  const float* d_test = b_test.Upload(test, static_cast<size_t>(num_test) * di);
  const int32_t* d_n = train_num_utts ? b_n.Upload(train_num_utts, static_cast<size_t>(num_train)) : nullptr;
  double* d_ytrain = b_ytrain.Upload<double>(nullptr, static_cast<size_t>(num_train) * d);
  double* d_ytest = b_ytest.Upload<double>(nullptr, static_cast<size_t>(num_test) * d);
  const int32_t* d_ti = b_ti.Upload(trial_train, static_cast<size_t>(num_trials));
  const int32_t* d_tj = b_tj.Upload(trial_test, static_cast<size_t>(num_trials));
  float* d_scores = b_scores.Upload<float>(nullptr, static_cast<size_t>(num_trials));
  CVX_CUDA_CHECK(LaunchPldaTransform(d_train, num_train, plda.dim_in, d_mean, d_tf, plda.dim, d_psi, d_n, plda.normalize_length,
                                     plda.simple_length_norm, d_ytrain, nullptr));
  CVX_CUDA_CHECK(LaunchPldaTransform(d_test, num_test, plda.dim_in, d_mean, d_tf, plda.dim, d_psi, nullptr, plda.normalize_length,
                                     plda.simple_length_norm, d_ytest, nullptr));
  CVX_CUDA_CHECK(LaunchPldaScore(d_ytrain, d_n, d_ytest, d_psi, plda.dim, d_ti, d_tj, num_trials, d_scores, nullptr));
  CVX_CUDA_CHECK(cudaMemcpy(scores, d_scores, static_cast<size_t>(num_trials) * 4, cudaMemcpyDeviceToHost));
}

// ------------------------------------------------------------ test hook ----
void TestStatsPool(int32_t device, int32_t precision, const float* in, int64_t rows, int32_t dim, const int64_t* begin,
                   const int64_t* end, int64_t num_chunks, float variance_floor, float* out) {
  CVX_CUDA_CHECK(cudaSetDevice(device));
  const int32_t elem = precision == CVX_PRECISION_FP16 ? kF16 : precision == CVX_PRECISION_BF16 ? kBF16 : kF32;
  const int32_t ld = RoundUp(dim, 64);
  std::vector<float> padded(static_cast<size_t>(rows) * ld, 0.f);
  for (int64_t r = 0; r < rows; ++r) std::memcpy(&padded[r * ld], in + r * dim, static_cast<size_t>(dim) * 4);
  std::vector<uint8_t> conv(padded.size() * ElemSize(elem));
  ConvertToElem(padded.data(), padded.size(), elem, conv.data());
  std::vector<PoolSlab> slabs;
  std::vector<int32_t> off(num_chunks + 1, 0);
  for (int64_t c = 0; c < num_chunks; ++c) {
    if (begin[c] < 0 || end[c] > rows || end[c] <= begin[c]) throw InvalidArgument("bad pooling range");
    off[c] = static_cast<int32_t>(slabs.size());
    for (int64_t r = begin[c]; r < end[c]; r += kPoolSlabRows)
      slabs.push_back({r, static_cast<int32_t>(std::min<int64_t>(kPoolSlabRows, end[c] - r)), static_cast<int32_t>(c)});
  }
  off[num_chunks] = static_cast<int32_t>(slabs.size());
  void *d_act = nullptr, *d_slabs = nullptr, *d_off = nullptr, *d_part = nullptr, *d_stats = nullptr;
  CVX_CUDA_CHECK(cudaMalloc(&d_act, std::max<size_t>(conv.size(), 16)));
  CVX_CUDA_CHECK(cudaMalloc(&d_slabs, std::max<size_t>(slabs.size(), 1) * sizeof(PoolSlab)));
  CVX_CUDA_CHECK(cudaMalloc(&d_off, off.size() * 4));
  CVX_CUDA_CHECK(cudaMalloc(&d_part, std::max<size_t>(slabs.size(), 1) * 2 * ld * 4));
  CVX_CUDA_CHECK(cudaMalloc(&d_stats, std::max<size_t>(num_chunks, 1) * 2 * dim * 4));
  CVX_CUDA_CHECK(cudaMemcpy(d_act, conv.data(), conv.size(), cudaMemcpyHostToDevice));
  CVX_CUDA_CHECK(cudaMemcpy(d_slabs, slabs.data(), slabs.size() * sizeof(PoolSlab), cudaMemcpyHostToDevice));
  CVX_CUDA_CHECK(cudaMemcpy(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
  CVX_CUDA_CHECK(LaunchStatsPoolPartial(d_act, elem, ld, static_cast<PoolSlab*>(d_slabs), static_cast<int32_t>(slabs.size()),
                                        static_cast<float*>(d_part), nullptr));
  CVX_CUDA_CHECK(LaunchStatsPoolFinalize(static_cast<float*>(d_part), ld, static_cast<PoolSlab*>(d_slabs),
                                         static_cast<int32_t*>(d_off), static_cast<int32_t>(num_chunks), dim, 1,
                                         variance_floor, static_cast<float*>(d_stats), 2 * dim, nullptr));
  CVX_CUDA_CHECK(cudaMemcpy(out, d_stats, static_cast<size_t>(num_chunks) * 2 * dim * 4, cudaMemcpyDeviceToHost));
  cudaFree(d_act);
  cudaFree(d_slabs);
  cudaFree(d_off);
  cudaFree(d_part);
  cudaFree(d_stats);
}


}  // namespace cvx
