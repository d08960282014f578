// Engine::ComputeFrames -- frame-level outputs (nnet3-compute).
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

// -------------------------------------------------------- frame outputs ----
// nnet3-compute's product: the output layer at every frame of every utterance.
// Utterances are packed back to back with `left` + `right` replicated edge frames each, so every
// real frame is an interior row of its padded utterance and one frame pass serves the whole batch.
void Engine::ComputeFrames(const float* const* feats, int32_t num_streams, int64_t total_rows, const int64_t* utt_start,
                           int32_t num_utts, float* out, uint32_t flags) {
  CVX_CUDA_CHECK(cudaSetDevice(opts_.device));
  if (!plan_.FrameOutput())
    throw InvalidArgument("this engine was built from an extraction model; load it with cvx_model_load_frames");
  if (num_streams != static_cast<int32_t>(plan_.input_dims.size()))
    throw InvalidArgument("model has " + std::to_string(plan_.input_dims.size()) + " input stream(s), got " +
                          std::to_string(num_streams));
  if (num_utts < 0 || total_rows < 0) throw InvalidArgument("negative size");
  if (num_utts == 0) return;
  if (utt_start[0] < 0 || utt_start[num_utts] > total_rows) throw InvalidArgument("utt_start outside [0, total_rows]");
  for (int32_t u = 0; u < num_utts; ++u)
    if (utt_start[u + 1] <= utt_start[u]) throw InvalidArgument("utterance " + std::to_string(u) + " has no frames");
  for (int32_t s = 0; s < num_streams; ++s)
    if (!feats[s]) throw InvalidArgument("feats[" + std::to_string(s) + "] is NULL");
  const bool dev_in = (flags & CVX_FLAG_DEVICE_INPUT) != 0, dev_out = (flags & CVX_FLAG_DEVICE_OUTPUT) != 0;
  const Layer& OL = plan_.frame[plan_.frame_out];
  const int32_t left = OL.left, right = OL.right, pad = left + right, E = OL.out_dim;
  const int32_t ld32 = RoundUp(E, 4);
  // batch = whole utterances; padded rows per batch bounded by max_batch_rows and by 4 GB of fp32 output
  const int64_t cap = std::min<int64_t>(max_batch_rows_, std::max<int64_t>((int64_t{1} << 30) / ld32, kGemmBlockM));
  struct FBatch {
    int32_t u0, u1;
    int64_t rows_c, rows_p;
    size_t off_pos;  // position of its utt_start copy in `rel`
  };
  std::vector<FBatch> batches;
  std::vector<int64_t> rel;  // per batch: utterance starts relative to the batch, n + 1 entries
  int64_t max_c = 0, max_p = 0;
  for (int32_t u0 = 0; u0 < num_utts;) {
    int32_t u1 = u0;
    int64_t rows_p = 0;
    while (u1 < num_utts) {
      const int64_t add = utt_start[u1 + 1] - utt_start[u1] + pad;
      if (u1 > u0 && rows_p + add > cap) break;
      rows_p += add;
      ++u1;
    }
    FBatch fb{u0, u1, utt_start[u1] - utt_start[u0], rows_p, rel.size()};
    for (int32_t u = u0; u <= u1; ++u) rel.push_back(utt_start[u] - utt_start[u0]);
    max_c = std::max(max_c, fb.rows_c);
    max_p = std::max(max_p, fb.rows_p);
    batches.push_back(fb);
    u0 = u1;
  }
  EnsureRows(max_p);
  if (!d2h_) {
    CVX_CUDA_CHECK(cudaStreamCreateWithFlags(&d2h_, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CVX_CUDA_CHECK(cudaEventCreateWithFlags(&out_ready_[i], cudaEventDisableTiming));
      CVX_CUDA_CHECK(cudaEventCreateWithFlags(&out_free_[i], cudaEventDisableTiming));
    }
  }
  Grow(&d_utt_off_, &utt_off_cap_, rel.size() * 8);
  CVX_CUDA_CHECK(cudaMemcpyAsync(d_utt_off_, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, compute_));
  for (int32_t s = 0; s < num_streams; ++s) {
    if (!dev_in)
      for (int i = 0; i < 2; ++i) Grow(&d_feat_[i][s], &d_feat_cap_[i][s], static_cast<size_t>(max_c) * plan_.input_dims[s] * 4);
    if (pad > 0) Grow(&d_sel_[s], &sel_cap_[s], static_cast<size_t>(max_p) * plan_.input_dims[s] * 4);
  }
  if (elem_ != kF32) Grow(&d_stats_, &stats_cap_, static_cast<size_t>(max_p) * ld32 * 4);
  if (!dev_out)
    for (int i = 0; i < 2; ++i) Grow(&d_frame_out_[i], &frame_out_cap_[i], static_cast<size_t>(max_c) * E * 4);
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));  // `rel` is pageable: its copy must have run

  // Three streams: H2D of batch i+1 and D2H of batch i-1 overlap the frame pass of batch i.
  CVX_CUDA_CHECK(cudaEventRecord(call_begin_, compute_));
  for (size_t bi = 0; bi < batches.size(); ++bi) {
    const FBatch& b = batches[bi];
    const int slot = static_cast<int>(bi & 1);
    const int32_t nu = b.u1 - b.u0;
    const int64_t r0 = utt_start[b.u0];
    const int64_t* d_rel = static_cast<const int64_t*>(d_utt_off_) + b.off_pos;
    const float* dev[kMaxASources] = {};
    if (!dev_in) {
      if (bi >= 2) CVX_CUDA_CHECK(cudaStreamWaitEvent(copy_, feat_free_[slot], 0));
      for (int32_t s = 0; s < num_streams; ++s) {
        const int32_t dim = plan_.input_dims[s];
        CVX_CUDA_CHECK(cudaMemcpyAsync(d_feat_[slot][s], feats[s] + r0 * dim, static_cast<size_t>(b.rows_c) * dim * 4,
                                       cudaMemcpyHostToDevice, copy_));
      }
      CVX_CUDA_CHECK(cudaEventRecord(feat_ready_[slot], copy_));
      CVX_CUDA_CHECK(cudaStreamWaitEvent(compute_, feat_ready_[slot], 0));
    }
    for (int32_t s = 0; s < num_streams; ++s) {
      const int32_t dim = plan_.input_dims[s];
      const float* src = dev_in ? feats[s] + r0 * dim : static_cast<const float*>(d_feat_[slot][s]);
      if (pad > 0) {
        ProfBegin(kProfStage, 0, static_cast<double>(b.rows_c + b.rows_p) * dim * 4.0);
        CVX_CUDA_CHECK(LaunchPadEdges(src, dim, d_rel, nu, left, right, b.rows_p, static_cast<float*>(d_sel_[s]), compute_));
        ProfEnd();
        src = static_cast<const float*>(d_sel_[s]);
      }
      dev[s] = src;
    }
    // with padding the frame pass reads d_sel_, so the raw rows are free as soon as the pad kernels ran
    if (!dev_in && pad > 0) CVX_CUDA_CHECK(cudaEventRecord(feat_free_[slot], compute_));
    const float* act = nullptr;
    int32_t act_ld = 0;
    if (elem_ == kF32) {
      RunFrames(dev, b.rows_p, 0, b.rows_p);
      const Buffer& ob = buffers_[frame_ops_.back().out_buffer];
      act = static_cast<const float*>(ob.ptr);
      act_ld = ob.ld;
    } else {
      RunFrames(dev, b.rows_p, 0, b.rows_p, nullptr, static_cast<float*>(d_stats_), ld32);
      act = static_cast<const float*>(d_stats_);
      act_ld = ld32;
    }
    if (!dev_in && pad == 0) CVX_CUDA_CHECK(cudaEventRecord(feat_free_[slot], compute_));
    float* dst = dev_out ? out + r0 * E : static_cast<float*>(d_frame_out_[slot]);
    if (!dev_out && bi >= 2) CVX_CUDA_CHECK(cudaStreamWaitEvent(compute_, out_free_[slot], 0));
    ProfBegin(kProfAverage, 0, static_cast<double>(b.rows_c) * E * 8.0);
    CVX_CUDA_CHECK(LaunchFrameOutput(act, act_ld, E, d_rel, nu, left, right, b.rows_c, plan_.frame_log_softmax ? 1 : 0,
                                     (flags & CVX_FLAG_APPLY_EXP) ? 1 : 0, dst, compute_));
    ProfEnd();
    if (!dev_out) {
      CVX_CUDA_CHECK(cudaEventRecord(out_ready_[slot], compute_));
      CVX_CUDA_CHECK(cudaStreamWaitEvent(d2h_, out_ready_[slot], 0));
      CVX_CUDA_CHECK(cudaMemcpyAsync(out + r0 * E, d_frame_out_[slot], static_cast<size_t>(b.rows_c) * E * 4,
                                     cudaMemcpyDeviceToHost, d2h_));
      CVX_CUDA_CHECK(cudaEventRecord(out_free_[slot], d2h_));
    }
  }
  if (!dev_out) {  // the call ends when the last results are on the host
    CVX_CUDA_CHECK(cudaEventRecord(out_ready_[0], d2h_));
    CVX_CUDA_CHECK(cudaStreamWaitEvent(compute_, out_ready_[0], 0));
  }
  CVX_CUDA_CHECK(cudaEventRecord(call_end_, compute_));
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));
  CVX_CUDA_CHECK(cudaStreamSynchronize(copy_));
  CVX_CUDA_CHECK(cudaStreamSynchronize(d2h_));
  CVX_CUDA_CHECK(cudaEventElapsedTime(&last_call_ms_, call_begin_, call_end_));
  ProfResolve();
}


}  // namespace cvx
