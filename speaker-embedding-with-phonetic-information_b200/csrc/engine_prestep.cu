// Feature pre-step on the device (sliding CMN + voiced-frame selection) and Engine::ExtractRaw.
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

// ------------------------------------------------------------ pre-step ----
// Pre-step kernels of utterances [u0, u1) (raw rows already on the device): the selected /
// normalised rows [kept0, kept0 + num_kept) of stream s.
void Engine::RunPrestep(int32_t s, int32_t dim, const cvx_prestep& pre, int32_t u0, int32_t u1, int64_t run0, int64_t num_runs,
                        int64_t kept0, int64_t num_kept, bool has_keep) {
  const int32_t* d_keep = has_keep ? static_cast<const int32_t*>(d_pre_keep_) : nullptr;
  const int64_t* aux = static_cast<const int64_t*>(d_pre_aux_);  // [n + 1] run starts, then [n + 1] kept starts
  ProfBegin(kProfStage, 0, static_cast<double>(num_runs) * CmnRunFrames() * dim * 4.0 + static_cast<double>(num_kept) * dim * 4.0);
  if (pre.cmn_window > 0)
    CVX_CUDA_CHECK(LaunchCmnSliding(static_cast<const float*>(d_raw_[s]), dim, static_cast<const int64_t*>(d_pre_utt_start_), aux,
                                    aux + pre_num_utts_ + 1, d_keep, u0, u1 - u0, run0, num_runs, pre.cmn_window,
                                    pre.cmn_center, pre.min_cmn_window, static_cast<float*>(d_sel_[s]), compute_));
  else
    CVX_CUDA_CHECK(LaunchCmnSelect(static_cast<const float*>(d_raw_[s]), dim, static_cast<const int64_t*>(d_pre_utt_start_),
                                   static_cast<const int32_t*>(d_pre_kept_utt_), d_keep, kept0, num_kept,
                                   static_cast<float*>(d_sel_[s]), compute_));
  ProfEnd();
}

// Uploads the per-call index arrays of the pre-step; returns the run offsets (host copy).
std::vector<int64_t> Engine::UploadPrestepIndex(const int64_t* utt_raw_start, const int32_t* keep, const int64_t* utt_kept_start,
                                                int32_t num_utts, const std::vector<int32_t>& kept_utt) {
  const int64_t total_kept = utt_kept_start[num_utts];
  const int32_t run = CmnRunFrames();
  std::vector<int64_t> aux(2 * (static_cast<size_t>(num_utts) + 1));
  aux[0] = 0;
  for (int32_t u = 0; u < num_utts; ++u) aux[u + 1] = aux[u] + (utt_raw_start[u + 1] - utt_raw_start[u] + run - 1) / run;
  for (int32_t u = 0; u <= num_utts; ++u) aux[num_utts + 1 + u] = utt_kept_start[u];
  pre_num_utts_ = num_utts;
  Grow(&d_pre_utt_start_, &pre_utt_start_cap_, static_cast<size_t>(num_utts + 1) * 8);
  Grow(&d_pre_kept_utt_, &pre_kept_utt_cap_, std::max<size_t>(kept_utt.size(), 1) * 4);
  Grow(&d_pre_aux_, &pre_aux_cap_, aux.size() * 8);
  CVX_CUDA_CHECK(cudaMemcpyAsync(d_pre_utt_start_, utt_raw_start, static_cast<size_t>(num_utts + 1) * 8, cudaMemcpyHostToDevice, compute_));
  CVX_CUDA_CHECK(cudaMemcpyAsync(d_pre_kept_utt_, kept_utt.data(), kept_utt.size() * 4, cudaMemcpyHostToDevice, compute_));
  CVX_CUDA_CHECK(cudaMemcpyAsync(d_pre_aux_, aux.data(), aux.size() * 8, cudaMemcpyHostToDevice, compute_));
  if (keep) {
    Grow(&d_pre_keep_, &pre_keep_cap_, std::max<size_t>(static_cast<size_t>(total_kept), 1) * 4);
    CVX_CUDA_CHECK(cudaMemcpyAsync(d_pre_keep_, keep, static_cast<size_t>(total_kept) * 4, cudaMemcpyHostToDevice, compute_));
  }
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));  // the host vectors have been consumed
  aux.resize(static_cast<size_t>(num_utts) + 1);
  return aux;
}

static void CheckPrestepIndex(int64_t total_raw_rows, const int64_t* utt_raw_start, const int32_t* keep,
                              const int64_t* utt_kept_start, int32_t num_utts, std::vector<int32_t>* kept_utt) {
  if (num_utts < 0 || !utt_raw_start || !utt_kept_start) throw InvalidArgument("pre-step: NULL utterance offsets");
  if (utt_raw_start[0] != 0 || utt_kept_start[0] != 0 || utt_raw_start[num_utts] != total_raw_rows)
    throw InvalidArgument("pre-step: utterance offsets must start at 0 and end at the row count");
  kept_utt->clear();
  kept_utt->reserve(static_cast<size_t>(utt_kept_start[num_utts]));
  for (int32_t u = 0; u < num_utts; ++u) {
    const int64_t T = utt_raw_start[u + 1] - utt_raw_start[u], K = utt_kept_start[u + 1] - utt_kept_start[u];
    if (T < 0 || K < 0 || K > T || T > INT32_MAX) throw InvalidArgument("pre-step: bad offsets for utterance " + std::to_string(u));
    if (!keep && K != T) throw InvalidArgument("pre-step: keep == NULL requires kept == raw row counts");
    int32_t prev = -1;
    for (int64_t j = utt_kept_start[u]; j < utt_kept_start[u + 1]; ++j) {
      if (keep) {
        if (keep[j] <= prev || keep[j] >= T) throw InvalidArgument("pre-step: keep[] must be increasing frame indices inside the utterance");
        prev = keep[j];
      }
      kept_utt->push_back(u);
    }
  }
}

void Engine::ExtractRaw(const float* const* raw, const cvx_prestep* pre, int32_t num_streams, int64_t total_raw_rows,
                        const int64_t* utt_raw_start, const int32_t* keep, const int64_t* utt_kept_start, int32_t num_utts,
                        const int64_t* chunk_start, const int32_t* chunk_len, const int32_t* chunk_utt,
                        const float* chunk_weight, int64_t num_chunks, float* out, uint32_t flags) {
  CVX_CUDA_CHECK(cudaSetDevice(opts_.device));
  if (flags & CVX_FLAG_DEVICE_INPUT) throw InvalidArgument("cvx_extract_raw takes host features");
  if (num_streams != static_cast<int32_t>(plan_.input_dims.size())) throw InvalidArgument("wrong number of input streams");
  if (!raw || !pre) throw InvalidArgument("pre-step: NULL argument");
  std::vector<int32_t> kept_utt;
  CheckPrestepIndex(total_raw_rows, utt_raw_start, keep, utt_kept_start, num_utts, &kept_utt);
  const int64_t total_kept = utt_kept_start[num_utts];
  const std::vector<int64_t> run_start = UploadPrestepIndex(utt_raw_start, keep, utt_kept_start, num_utts, kept_utt);
  // all Grow()s first: they synchronise and may free buffers earlier kernels of this call read
  for (int32_t s = 0; s < num_streams; ++s) {
    if (plan_.input_dims[s] > 128 && pre[s].cmn_window > 0) throw InvalidArgument("sliding CMN supports feature dims up to 128");
    Grow(&d_raw_[s], &raw_cap_[s], static_cast<size_t>(total_raw_rows) * plan_.input_dims[s] * 4);
    Grow(&d_sel_[s], &sel_cap_[s], static_cast<size_t>(std::max<int64_t>(total_kept, 1)) * plan_.input_dims[s] * 4);
  }
  const float* sel[kMaxASources] = {};
  for (int32_t s = 0; s < num_streams; ++s) {
    if (!raw[s]) throw InvalidArgument("pre-step: NULL feature stream");
    if (pre[s].norm_vars) throw InvalidArgument("sliding CMN with --norm-vars=true is not on this path (the recipes use false)");
    if (pre[s].cmn_window < 0) throw InvalidArgument("negative cmn_window");
    sel[s] = static_cast<const float*>(d_sel_[s]);
  }

  // Pieces of whole utterances (>= kPieceRows raw rows each): all raw copies are queued on the copy
  // stream now; the pre-step kernels of a piece are queued on the compute stream right before the
  // first device batch that reads its rows, so the copy of piece p+1 overlaps the GEMMs of piece p.
  struct Piece {
    int32_t u0, u1;
  };
  constexpr int64_t kPieceRows = 1 << 18;
  std::vector<Piece> pieces;
  for (int32_t u = 0; u < num_utts;) {
    int32_t v = u;
    while (v < num_utts && utt_raw_start[v] - utt_raw_start[u] < kPieceRows) ++v;
    pieces.push_back({u, v});
    u = v;
  }
  while (piece_events_.size() < pieces.size()) {
    cudaEvent_t ev;
    CVX_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    piece_events_.push_back(ev);
  }
  for (size_t p = 0; p < pieces.size(); ++p) {
    const int64_t r0 = utt_raw_start[pieces[p].u0], r1 = utt_raw_start[pieces[p].u1];
    for (int32_t s = 0; s < num_streams; ++s) {
      const int32_t dim = plan_.input_dims[s];
      CVX_CUDA_CHECK(cudaMemcpyAsync(static_cast<float*>(d_raw_[s]) + r0 * dim, raw[s] + r0 * dim,
                                     static_cast<size_t>(r1 - r0) * dim * 4, cudaMemcpyHostToDevice, copy_));
    }
    CVX_CUDA_CHECK(cudaEventRecord(piece_events_[p], copy_));
  }
  size_t next_piece = 0;
  before_batch_ = [&](int64_t row_end) {
    while (next_piece < pieces.size() && utt_kept_start[pieces[next_piece].u0] < row_end) {
      const Piece& pc = pieces[next_piece];
      CVX_CUDA_CHECK(cudaStreamWaitEvent(compute_, piece_events_[next_piece], 0));
      const int64_t k0 = utt_kept_start[pc.u0], k1 = utt_kept_start[pc.u1];
      for (int32_t s = 0; s < num_streams; ++s)
        RunPrestep(s, plan_.input_dims[s], pre[s], pc.u0, pc.u1, run_start[pc.u0], run_start[pc.u1] - run_start[pc.u0], k0,
                   k1 - k0, keep != nullptr);
      ++next_piece;
    }
  };
  try {
    Extract(sel, num_streams, total_kept, chunk_start, chunk_len, chunk_utt, chunk_weight, num_chunks, num_utts, out,
            flags | CVX_FLAG_DEVICE_INPUT);
  } catch (...) {
    before_batch_ = nullptr;
    cudaStreamSynchronize(copy_);
    throw;
  }
  before_batch_ = nullptr;
  CVX_CUDA_CHECK(cudaStreamSynchronize(copy_));  // pieces no batch needed (utterances without chunks)
}

void Engine::DebugPrestep(const float* raw, int32_t dim, const cvx_prestep& pre, int64_t total_raw_rows,
                          const int64_t* utt_raw_start, const int32_t* keep, const int64_t* utt_kept_start,
                          int32_t num_utts, float* out) {
  CVX_CUDA_CHECK(cudaSetDevice(opts_.device));
  std::vector<int32_t> kept_utt;
  CheckPrestepIndex(total_raw_rows, utt_raw_start, keep, utt_kept_start, num_utts, &kept_utt);
  const int64_t total_kept = utt_kept_start[num_utts];
  const std::vector<int64_t> run_start = UploadPrestepIndex(utt_raw_start, keep, utt_kept_start, num_utts, kept_utt);
  if (dim > 128 && pre.cmn_window > 0) throw InvalidArgument("sliding CMN supports feature dims up to 128");
  Grow(&d_raw_[0], &raw_cap_[0], static_cast<size_t>(total_raw_rows) * dim * 4);
  Grow(&d_sel_[0], &sel_cap_[0], static_cast<size_t>(std::max<int64_t>(total_kept, 1)) * dim * 4);
  if (pre.norm_vars) throw InvalidArgument("sliding CMN with --norm-vars=true is not on this path (the recipes use false)");
  if (pre.cmn_window < 0) throw InvalidArgument("negative cmn_window");
  CVX_CUDA_CHECK(cudaMemcpyAsync(d_raw_[0], raw, static_cast<size_t>(total_raw_rows) * dim * 4, cudaMemcpyHostToDevice, compute_));
  RunPrestep(0, dim, pre, 0, num_utts, 0, run_start[num_utts], 0, total_kept, keep != nullptr);
  CVX_CUDA_CHECK(cudaMemcpyAsync(out, d_sel_[0], static_cast<size_t>(total_kept) * dim * 4, cudaMemcpyDeviceToHost, compute_));
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));
  ProfResolve();
}


}  // namespace cvx
