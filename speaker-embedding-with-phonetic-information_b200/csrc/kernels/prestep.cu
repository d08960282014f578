// Feature pre-step on the device: sliding-window cepstral mean normalisation and
// voiced-frame selection, i.e. what the reference's feature pipe does with two extra
// processes before the hot path
//   apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300 scp:feats.scp ark:- |
//   select-voiced-frames ark:- scp,s,cs:vad.scp ark:- |
// (egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:79; the frame-selection rule is the
// one src/ivectorbin/select-voiced-ali.cc:58-91 of the reference applies to labels: keep
// frames whose VAD value is non-zero, in order).  Window rule restated from upstream
// Kaldi's SlidingWindowCmn (src/feat/feature-functions.cc), sums in double like upstream.
//
//   cmn_prefix_kernel   one warp per utterance: P[t+1][c] = P[t][c] + x[t][c] (double),
//                       lanes = feature columns, so every row is one coalesced read
//   cmn_select_kernel   one thread per (kept frame, column): window [ws, we) of its source
//                       frame, out = x[t] - (P[we] - P[ws]) / (we - ws); writes the compacted
//                       fp32 rows the rest of the path consumes
// HBM-bound: 92 B read + 184 B of prefix written per raw frame, then ~280 B per kept frame.
#include "kernels.h"

namespace cvx {
namespace {

__global__ void __launch_bounds__(128) cmn_prefix_kernel(const float* __restrict__ x, int dim,
                                                          const int64_t* __restrict__ utt_start, int num_utts,
                                                          double* __restrict__ prefix) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= num_utts) return;
  const int64_t r0 = utt_start[warp], r1 = utt_start[warp + 1];
  double* p = prefix + (r0 + warp) * dim;  // utterance u owns rows [r0+u, r1+u+1): T+1 prefix rows
  for (int c = lane; c < dim; c += 32) {
    double s = 0.0;
    p[c] = 0.0;
    for (int64_t t = r0; t < r1; ++t) {
      s += static_cast<double>(x[t * dim + c]);
      p[(t - r0 + 1) * dim + c] = s;
    }
  }
}

__global__ void __launch_bounds__(256) cmn_select_kernel(const float* __restrict__ x, int dim,
                                                          const int64_t* __restrict__ utt_start,
                                                          const int32_t* __restrict__ kept_utt,
                                                          const int32_t* __restrict__ keep, int64_t total_kept,
                                                          const double* __restrict__ prefix, int cmn_window, int center,
                                                          int min_window, float* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t j = idx / dim;
  const int c = static_cast<int>(idx % dim);
  if (j >= total_kept) return;
  const int u = kept_utt[j];
  const int64_t r0 = utt_start[u];
  const int num_frames = static_cast<int>(utt_start[u + 1] - r0);
  const int t = keep ? keep[j] : static_cast<int>(j - r0);  // frame inside its utterance
  float v = x[(r0 + t) * dim + c];
  if (cmn_window > 0) {
    int ws, we;  // [ws, we)
    if (center) {
      ws = t - cmn_window / 2;
      we = ws + cmn_window;
    } else {
      ws = t - cmn_window;
      we = t + 1;
    }
    if (ws < 0) {
      we -= ws;
      ws = 0;
    }
    if (!center && we > t) we = max(t + 1, min_window);
    if (we > num_frames) {
      ws -= (we - num_frames);
      we = num_frames;
      if (ws < 0) ws = 0;
    }
    const double* p = prefix + (r0 + u) * dim;
    const double mean = (p[static_cast<int64_t>(we) * dim + c] - p[static_cast<int64_t>(ws) * dim + c]) / (we - ws);
    v = static_cast<float>(static_cast<double>(v) - mean);
  }
  out[j * dim + c] = v;
}

}  // namespace

cudaError_t LaunchCmnPrefix(const float* x, int32_t dim, const int64_t* utt_start, int32_t num_utts, double* prefix,
                            cudaStream_t stream) {
  if (num_utts <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((static_cast<int64_t>(num_utts) * 32 + 127) / 128);
  cmn_prefix_kernel<<<grid, 128, 0, stream>>>(x, dim, utt_start, num_utts, prefix);
  return cudaGetLastError();
}

cudaError_t LaunchCmnSelect(const float* x, int32_t dim, const int64_t* utt_start, const int32_t* kept_utt,
                            const int32_t* keep, int64_t total_kept, const double* prefix, int32_t cmn_window,
                            int32_t center, int32_t min_window, float* out, cudaStream_t stream) {
  if (total_kept <= 0) return cudaSuccess;
  const int64_t total = total_kept * dim;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  cmn_select_kernel<<<grid, 256, 0, stream>>>(x, dim, utt_start, kept_utt, keep, total_kept, prefix, cmn_window, center,
                                              min_window, out);
  return cudaGetLastError();
}

}  // namespace cvx
