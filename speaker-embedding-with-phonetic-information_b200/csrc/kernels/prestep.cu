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
//   cmn_sliding_kernel  one warp per run of 128 consecutive frames of one utterance, lanes = feature
//                       columns: the window sum of the run's first frame is summed directly
//                       (<= cmn_window coalesced row reads), then the window slides frame by frame
//                       (add the entering row, subtract the leaving one, in double) -- the same
//                       running sum upstream keeps; out = x[t] - sum / (we - ws) is written for
//                       the frames that survive selection only, at their compacted position
//   cmn_select_kernel   plain selection (no CMN for this stream): one thread per (kept frame, column)
// The rows a run re-reads are in L1/L2 (neighbouring runs share their windows): ~92 B of HBM reads
// per raw frame and 92 B written per kept frame.  (The first version kept a double prefix-sum
// array built by one warp per UTTERANCE: 64 warps per call, 270 ms for 9 M frames.)
#include "kernels.h"

namespace cvx {
namespace {

constexpr int kCmnRun = 128;     // frames per warp
constexpr int kCmnMaxColGroups = 4;  // feature dims up to 128

// [ws, we) of frame t in an utterance of num_frames frames (SlidingWindowCmn's rule)
__device__ __forceinline__ void cmn_window_bounds(int t, int num_frames, int cmn_window, int center, int min_window, int* ws_out,
                                                  int* we_out) {
  int ws, we;
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
  *ws_out = ws;
  *we_out = we;
}

__global__ void __launch_bounds__(128) cmn_sliding_kernel(const float* __restrict__ x, int dim,
                                                           const int64_t* __restrict__ utt_start,
                                                           const int64_t* __restrict__ utt_run_start,
                                                           const int64_t* __restrict__ utt_kept_start,
                                                           const int32_t* __restrict__ keep, int utt0, int num_utts,
                                                           int64_t run0, int64_t num_runs, int cmn_window, int center,
                                                           int min_window, float* __restrict__ out) {
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= num_runs) return;
  const int64_t run = run0 + w;
  int lo = utt0, hi = utt0 + num_utts - 1;  // largest u with utt_run_start[u] <= run (empty utterances have no runs)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (utt_run_start[mid] <= run) lo = mid;
    else hi = mid - 1;
  }
  const int u = lo;
  const int64_t r0 = utt_start[u];
  const int num_frames = static_cast<int>(utt_start[u + 1] - r0);
  const int t0 = static_cast<int>(run - utt_run_start[u]) * kCmnRun, t1 = min(num_frames, t0 + kCmnRun);
  // position of the run's first surviving frame among the kept rows
  int64_t kp = utt_kept_start[u], ke = utt_kept_start[u + 1];
  if (keep) {
    int64_t a = kp, b = ke;
    while (a < b) {
      const int64_t mid = (a + b) >> 1;
      if (keep[mid] >= t0) b = mid;
      else a = mid + 1;
    }
    kp = a;
  } else {
    kp += t0;
  }
  const float* xu = x + r0 * dim;
  double sum[kCmnMaxColGroups] = {0.0, 0.0, 0.0, 0.0};
  int ws, we;
  cmn_window_bounds(t0, num_frames, cmn_window, center, min_window, &ws, &we);
  for (int r = ws; r < we; ++r)
#pragma unroll
    for (int k = 0; k < kCmnMaxColGroups; ++k)
      if (lane + 32 * k < dim) sum[k] += static_cast<double>(xu[static_cast<int64_t>(r) * dim + lane + 32 * k]);
  for (int t = t0; t < t1; ++t) {
    int nws, nwe;
    cmn_window_bounds(t, num_frames, cmn_window, center, min_window, &nws, &nwe);
    for (; we < nwe; ++we)
#pragma unroll
      for (int k = 0; k < kCmnMaxColGroups; ++k)
        if (lane + 32 * k < dim) sum[k] += static_cast<double>(xu[static_cast<int64_t>(we) * dim + lane + 32 * k]);
    for (; ws < nws; ++ws)
#pragma unroll
      for (int k = 0; k < kCmnMaxColGroups; ++k)
        if (lane + 32 * k < dim) sum[k] -= static_cast<double>(xu[static_cast<int64_t>(ws) * dim + lane + 32 * k]);
    const bool kept = keep ? (kp < ke && keep[kp] == t) : true;
    if (kept) {
      const double inv = 1.0 / static_cast<double>(we - ws);
#pragma unroll
      for (int k = 0; k < kCmnMaxColGroups; ++k)
        if (lane + 32 * k < dim) {
          const int c = lane + 32 * k;
          out[kp * dim + c] = static_cast<float>(static_cast<double>(xu[static_cast<int64_t>(t) * dim + c]) - sum[k] * inv);
        }
      ++kp;
    }
  }
}

__global__ void __launch_bounds__(256) cmn_select_kernel(const float* __restrict__ x, int dim,
                                                          const int64_t* __restrict__ utt_start,
                                                          const int32_t* __restrict__ kept_utt,
                                                          const int32_t* __restrict__ keep, int64_t kept0,
                                                          int64_t num_kept, float* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c = static_cast<int>(idx % dim);
  if (idx / dim >= num_kept) return;
  const int64_t j = kept0 + idx / dim;  // kept rows [kept0, kept0 + num_kept) of the call
  const int64_t r0 = utt_start[kept_utt[j]];
  const int t = keep ? keep[j] : static_cast<int>(j - r0);  // frame inside its utterance
  out[j * dim + c] = x[(r0 + t) * dim + c];
}

}  // namespace

cudaError_t LaunchCmnSliding(const float* x, int32_t dim, const int64_t* utt_start, const int64_t* utt_run_start,
                             const int64_t* utt_kept_start, const int32_t* keep, int32_t utt0, int32_t num_utts, int64_t run0,
                             int64_t num_runs, int32_t cmn_window, int32_t center, int32_t min_window, float* out,
                             cudaStream_t stream) {
  if (num_runs <= 0 || num_utts <= 0) return cudaSuccess;
  if (dim > 32 * kCmnMaxColGroups) return cudaErrorInvalidValue;
  const unsigned grid = static_cast<unsigned>((num_runs * 32 + 127) / 128);
  cmn_sliding_kernel<<<grid, 128, 0, stream>>>(x, dim, utt_start, utt_run_start, utt_kept_start, keep, utt0, num_utts, run0,
                                               num_runs, cmn_window, center, min_window, out);
  return cudaGetLastError();
}

int32_t CmnRunFrames() { return kCmnRun; }

cudaError_t LaunchCmnSelect(const float* x, int32_t dim, const int64_t* utt_start, const int32_t* kept_utt,
                            const int32_t* keep, int64_t kept0, int64_t num_kept, float* out, cudaStream_t stream) {
  if (num_kept <= 0) return cudaSuccess;
  const int64_t total = num_kept * dim;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  cmn_select_kernel<<<grid, 256, 0, stream>>>(x, dim, utt_start, kept_utt, keep, kept0, num_kept, out);
  return cudaGetLastError();
}

}  // namespace cvx
