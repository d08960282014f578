// K3 -- segment-level layers and the per-utterance frame-weighted average.
//
// Replaces the `[1,3000] x [3000,512]` affine of tdnn6[_xvec].affine that upstream
// Kaldi runs once per chunk (output node chosen by
// egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:58) by one batched
// [num_chunks, 3000] x [3000, 512] fp32 contraction, and the host-side
// `xvector_avg.AddVec(offset, xvector)` / `Scale(1/tot)` loop of
// nnet3-xvector-compute.cc by a segmented reduction on the device.
// fp32 CUDA-core math: this stage is ~0.02 % of the path's FLOPs for long chunks and
// its operands (pooled statistics) are kept in full precision.
#include "kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

namespace cvx {
namespace {

// x -> (hi, lo) with hi = round16(x), lo = round16(x - hi): the tensor-core segment GEMM
// computes hi*Wh + lo*Wh + hi*Wl, i.e. an fp32-accurate product from 16-bit operands.
template <typename T>
__global__ void __launch_bounds__(256) split_f32_kernel(const float* __restrict__ in, int ld_in, int64_t m, int k,
                                                         T* __restrict__ hi, T* __restrict__ lo, int ld_out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t row = idx / ld_out;
  const int c = static_cast<int>(idx % ld_out);
  if (row >= m) return;
  float x = c < k ? in[row * ld_in + c] : 0.f;
  T h, l;
  if constexpr (sizeof(T) == 2 && std::is_same<T, __half>::value) {
    x = fminf(fmaxf(x, -65504.f), 65504.f);
    h = __float2half_rn(x);
    l = __float2half_rn(x - __half2float(h));
  } else {
    h = __float2bfloat16_rn(x);
    l = __float2bfloat16_rn(x - __bfloat162float(h));
  }
  hi[row * ld_out + c] = h;
  lo[row * ld_out + c] = l;
}

constexpr int kTile = 64;
constexpr int kTileK = 16;

__global__ void __launch_bounds__(256) segment_affine_kernel(const float* __restrict__ in, int ld_in, int64_t m, int k,
                                                              const float* __restrict__ w,
                                                              const float* __restrict__ bias, int relu,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ offset, int n,
                                                              float* __restrict__ out, int ld_out) {
  __shared__ float As[kTileK][kTile + 1];
  __shared__ float Ws[kTileK][kTile + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kTile;
  const int n0 = blockIdx.y * kTile;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < k; k0 += kTileK) {
    for (int i = threadIdx.x; i < kTile * kTileK; i += 256) {
      const int r = i / kTileK, c = i % kTileK;
      const int64_t row = m0 + r;
      const int kk = k0 + c;
      As[c][r] = (row < m && kk < k) ? in[row * ld_in + kk] : 0.f;
      const int nn = n0 + r;
      Ws[c][r] = (nn < n && kk < k) ? w[static_cast<int64_t>(nn) * k + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTileK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = As[kk][ty * 4 + i];
        b[i] = Ws[kk][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = m0 + ty * 4 + i;
    if (row >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tx * 4 + j;
      if (nn >= n) continue;
      float v = acc[i][j] + (bias ? bias[nn] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      if (scale) v = v * scale[nn] + offset[nn];
      out[row * ld_out + nn] = v;
    }
  }
}

__global__ void __launch_bounds__(128) utt_average_kernel(const float* __restrict__ chunk_emb, int dim,
                                                           const int64_t* __restrict__ utt_chunk_off,
                                                           const float* __restrict__ chunk_weight,
                                                           float* __restrict__ out) {
  const int u = blockIdx.x;
  const int64_t c0 = utt_chunk_off[u], c1 = utt_chunk_off[u + 1];
  float tot = 0.f;
  for (int64_t c = c0; c < c1; ++c) tot += chunk_weight[c];
  for (int j = threadIdx.x; j < dim; j += blockDim.x) {
    float acc = 0.f;
    for (int64_t c = c0; c < c1; ++c) acc = fmaf(chunk_weight[c], chunk_emb[c * dim + j], acc);
    out[static_cast<int64_t>(u) * dim + j] = (c1 > c0) ? acc / tot : 0.f;
  }
}

}  // namespace

cudaError_t LaunchSplitF32(const float* in, int32_t ld_in, int64_t m, int32_t k, void* hi, void* lo,
                           int32_t ld_out, int32_t elem, cudaStream_t stream) {
  if (m <= 0) return cudaSuccess;
  const int64_t total = m * ld_out;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (elem == kF16)
    split_f32_kernel<__half><<<grid, 256, 0, stream>>>(in, ld_in, m, k, static_cast<__half*>(hi), static_cast<__half*>(lo), ld_out);
  else
    split_f32_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(in, ld_in, m, k, static_cast<__nv_bfloat16*>(hi),
                                                              static_cast<__nv_bfloat16*>(lo), ld_out);
  return cudaGetLastError();
}

cudaError_t LaunchSegmentAffine(const float* in, int32_t ld_in, int64_t m, int32_t k, const float* w,
                                const float* bias, int32_t relu, const float* scale, const float* offset,
                                int32_t n, float* out, int32_t ld_out, cudaStream_t stream) {
  if (m <= 0) return cudaSuccess;
  dim3 grid(static_cast<unsigned>((m + kTile - 1) / kTile), (n + kTile - 1) / kTile);
  segment_affine_kernel<<<grid, 256, 0, stream>>>(in, ld_in, m, k, w, bias, relu, scale, offset, n, out, ld_out);
  return cudaGetLastError();
}

cudaError_t LaunchUttAverage(const float* chunk_emb, int32_t dim, const int64_t* utt_chunk_off,
                             const float* chunk_weight, int32_t num_utts, float* out, cudaStream_t stream) {
  if (num_utts <= 0) return cudaSuccess;
  utt_average_kernel<<<num_utts, 128, 0, stream>>>(chunk_emb, dim, utt_chunk_off, chunk_weight, out);
  return cudaGetLastError();
}

}  // namespace cvx
