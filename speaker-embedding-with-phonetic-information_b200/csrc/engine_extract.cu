// Engine::Extract -- the hot path: packed ragged batches of chunks -> per-utterance embeddings.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
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

// A device batch runs the sparse (offset-forwarding) form only if every utterance with a chunk in it pools at
// least this many frames: the form's slightly larger activation rounding error needs the pooling average (engine.cu).
constexpr int32_t kSparseMinPooledFrames = 32;

// -------------------------------------------------------------- extract ----
void Engine::Extract(const float* const* feats, int32_t num_streams, int64_t total_rows, const int64_t* chunk_start,
                     const int32_t* chunk_len, const int32_t* chunk_utt, const float* chunk_weight,
                     int64_t num_chunks, int32_t num_utts, float* out, uint32_t flags) {
  // CVX_TRACE=1: host-side phases of every call on stderr (where a call's wall time goes beyond its device time)
  static const bool trace = std::getenv("CVX_TRACE") != nullptr;
  const auto t_enter = std::chrono::steady_clock::now();
  auto ms_since = [](std::chrono::steady_clock::time_point t) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
  };
  CVX_CUDA_CHECK(cudaSetDevice(opts_.device));
  if (num_streams != static_cast<int32_t>(plan_.input_dims.size()))
    throw InvalidArgument("model has " + std::to_string(plan_.input_dims.size()) + " input stream(s), got " +
                          std::to_string(num_streams));
  if (plan_.FrameOutput())
    throw InvalidArgument("this engine was built from a frame-output model (cvx_model_load_frames); use cvx_compute_frames");
  if (num_utts < 0 || num_chunks < 0 || total_rows < 0) throw InvalidArgument("negative size");
  if (!out && num_utts > 0) throw InvalidArgument("out is NULL");
  const int32_t E = plan_.EmbeddingDim();
  const bool dev_in = (flags & CVX_FLAG_DEVICE_INPUT) != 0, dev_out = (flags & CVX_FLAG_DEVICE_OUTPUT) != 0;
  if (num_utts == 0) return;
  if (num_chunks == 0) {
    if (dev_out) CVX_CUDA_CHECK(cudaMemset(out, 0, static_cast<size_t>(num_utts) * OutputDim() * 4));
    else std::memset(out, 0, static_cast<size_t>(num_utts) * OutputDim() * 4);
    return;
  }
  for (int32_t s = 0; s < num_streams; ++s)
    if (!feats || !feats[s]) throw InvalidArgument("feats[" + std::to_string(s) + "] is NULL");

  const Layer& PL = plan_.frame[plan_.pool_src];
  const int32_t L = PL.left, R = PL.right;
  const int32_t lc = plan_.pool_left_context, rc = plan_.pool_right_context;
  (void)lc;  // left >= 0 >= -lc: the window's left edge never binds

  // ---- validate chunks, cut batches, build pooling slabs ----
  struct Batch {
    int64_t c0, c1, row0, row1;
    int64_t slab0, slab1;
  };
  std::vector<Batch> batches;
  std::vector<PoolSlab> slabs;
  std::vector<int64_t> utt_off(static_cast<size_t>(num_utts) + 1, 0);
  std::vector<float> weights(num_chunks);
  int64_t prev_end = 0;
  int32_t prev_utt = 0;
  int64_t max_chunk_rows = 0;
  for (int64_t c = 0; c < num_chunks; ++c) {
    const int64_t st = chunk_start[c];
    const int32_t len = chunk_len[c];
    if (st < prev_end || len <= 0 || st + len > total_rows)
      throw InvalidArgument("chunk " + std::to_string(c) + " is out of order, overlapping or out of range");
    if (chunk_utt[c] < prev_utt || chunk_utt[c] >= num_utts)
      throw InvalidArgument("chunk_utt must be non-decreasing and < num_utts");
    if (std::min<int64_t>(len - 1 - R, rc) < L)
      throw InvalidArgument("chunk " + std::to_string(c) + " has " + std::to_string(len) + " rows but the network needs " +
                            std::to_string(L + R + 1) + " (left " + std::to_string(L) + ", right " + std::to_string(R) + ")");
    prev_end = st + len;
    prev_utt = chunk_utt[c];
    utt_off[chunk_utt[c] + 1] += 1;
    weights[c] = chunk_weight ? chunk_weight[c] : static_cast<float>(len);
    max_chunk_rows = std::max<int64_t>(max_chunk_rows, len);
  }
  for (int32_t u = 0; u < num_utts; ++u) utt_off[u + 1] += utt_off[u];
  const int64_t cap = std::max(max_batch_rows_, (max_chunk_rows + kGemmBlockM - 1) / kGemmBlockM * kGemmBlockM);
  // valid (poolable) rows of every chunk, relative to its batch; hi inclusive
  std::vector<int32_t> chunk_lo(num_chunks), chunk_hi(num_chunks);
  {
    int64_t c = 0;
    while (c < num_chunks) {
      Batch b;
      b.c0 = c;
      b.row0 = chunk_start[c];
      b.slab0 = b.slab1 = 0;
      while (c < num_chunks && chunk_start[c] + chunk_len[c] - b.row0 <= cap) {
        chunk_lo[c] = static_cast<int32_t>(chunk_start[c] - b.row0 + L);
        chunk_hi[c] = static_cast<int32_t>(chunk_start[c] - b.row0 + std::min<int64_t>(chunk_len[c] - 1 - R, rc));
        ++c;
      }
      b.c1 = c;
      b.row1 = chunk_start[c - 1] + chunk_len[c - 1];
      batches.push_back(b);
    }
  }
  // pooled frames per utterance (all of its chunks): decides the activation form of its batches
  std::vector<int64_t> utt_pooled(static_cast<size_t>(num_utts), 0);
  for (int64_t c = 0; c < num_chunks; ++c) utt_pooled[chunk_utt[c]] += chunk_hi[c] - chunk_lo[c] + 1;
  // ---- fused pooling: per batch, every 32-row group holds at most two runs of valid rows (tail of
  // one chunk, head of the next); slot = one run.  The host only numbers the slots per chunk, the
  // per-group table itself is built on the device (pool_meta_kernel), so this is O(chunks). ----
  bool fused = fused_pool_;
  std::vector<int32_t> slot_begin_packed;  // per batch: [nchunks + 1] slot offsets
  std::vector<int64_t> meta_pos(batches.size()), slot_pos(batches.size()), slot_begin_pos(batches.size()),
      batch_groups(batches.size());
  int64_t max_slots = 0, total_groups = 0, total_slots = 0;
  for (size_t bi = 0; bi < batches.size() && fused; ++bi) {
    const Batch& b = batches[bi];
    const int64_t rows = b.row1 - b.row0;
    batch_groups[bi] = (rows + 2 * kGemmBlockM - 1) / (2 * kGemmBlockM) * (2 * kGemmBlockM) / 32;
    meta_pos[bi] = total_groups;
    slot_pos[bi] = total_slots;
    slot_begin_pos[bi] = static_cast<int64_t>(slot_begin_packed.size());
    int32_t nslots = 0;
    for (int64_t c = b.c0; c < b.c1; ++c) {
      // three runs in one group (tiny chunks): use the stand-alone pooling kernels
      if (c + 2 < b.c1 && chunk_lo[c + 2] / 32 == chunk_hi[c] / 32) fused = false;
      slot_begin_packed.push_back(nslots);
      nslots += chunk_hi[c] / 32 - chunk_lo[c] / 32 + 1;
    }
    slot_begin_packed.push_back(nslots);
    max_slots = std::max<int64_t>(max_slots, nslots);
    total_groups += batch_groups[bi];
    total_slots += nslots;
  }

  // stand-alone pooling: slabs of <= kPoolSlabRows valid rows cut relative to each chunk;
  // chunk_slab_off is stored per batch as [c1-c0+1] entries relative to the batch's first slab
  std::vector<int32_t> slab_off_packed;
  std::vector<int64_t> slab_off_pos(batches.size());
  if (!fused) {
    for (size_t bi = 0; bi < batches.size(); ++bi) {
      Batch& b = batches[bi];
      b.slab0 = static_cast<int64_t>(slabs.size());
      slab_off_pos[bi] = static_cast<int64_t>(slab_off_packed.size());
      for (int64_t c = b.c0; c < b.c1; ++c) {
        slab_off_packed.push_back(static_cast<int32_t>(slabs.size() - b.slab0));
        for (int64_t r = chunk_lo[c]; r <= chunk_hi[c]; r += kPoolSlabRows) {
          PoolSlab sl;
          sl.row_begin = r;
          sl.nrows = static_cast<int32_t>(std::min<int64_t>(kPoolSlabRows, chunk_hi[c] - r + 1));
          sl.chunk = static_cast<int32_t>(c - b.c0);
          slabs.push_back(sl);
        }
      }
      b.slab1 = static_cast<int64_t>(slabs.size());
      slab_off_packed.push_back(static_cast<int32_t>(b.slab1 - b.slab0));
    }
  }

  // ---- device scratch ----
  const double ms_tables = ms_since(t_enter);
  int64_t max_rows = 0, max_chunks = 0, max_slabs = 0;
  for (const Batch& b : batches) {
    max_rows = std::max(max_rows, b.row1 - b.row0);
    max_chunks = std::max(max_chunks, b.c1 - b.c0);
    max_slabs = std::max(max_slabs, b.slab1 - b.slab0);
  }
  EnsureRows(max_rows);
  const Buffer& pool_buf = buffers_[layer_buffer_[plan_.pool_src]];
  const int32_t D = PL.out_dim, pooled = plan_.PooledDim();
  Grow(&d_slabs_, &slabs_cap_, std::max<size_t>(slabs.size(), 1) * sizeof(PoolSlab));
  Grow(&d_chunk_slab_off_, &chunk_slab_off_cap_, std::max<size_t>(slab_off_packed.size(), 1) * 4);
  Grow(&d_weights_, &weights_cap_, weights.size() * 4);
  Grow(&d_utt_off_, &utt_off_cap_, utt_off.size() * 8);
  Grow(&d_partials_, &partials_cap_, static_cast<size_t>(std::max(max_slabs, fused ? max_slots : 0)) * 2 * pool_buf.ld * 4);
  if (fused) {
    Grow(&d_pool_meta_, &pool_meta_cap_, static_cast<size_t>(total_groups) * sizeof(PoolGroupMeta));
    Grow(&d_slot_n_, &slot_n_cap_, static_cast<size_t>(std::max<int64_t>(total_slots, 1)) * 4);
    Grow(&d_slot_begin_, &slot_begin_cap_, slot_begin_packed.size() * 4);
    Grow(&d_chunk_rng_, &chunk_rng_cap_, static_cast<size_t>(num_chunks) * 8);
  }
  // pooled statistics of several device batches are collected and the segment layers run on
  // up to kSegFlushChunks chunks at once (a [few chunks, 3000] GEMM per batch is latency bound)
  constexpr int64_t kSegFlushChunks = 16384;
  const int64_t seg_cap_chunks = std::max(max_chunks, std::min<int64_t>(num_chunks, kSegFlushChunks));
  Grow(&d_stats_, &stats_cap_, static_cast<size_t>(seg_cap_chunks) * pooled * 4);
  int32_t max_seg = 0;
  for (size_t i = 0; i + 1 < seg_ops_.size(); ++i) max_seg = std::max(max_seg, seg_ops_[i].out_dim);
  for (int i = 0; i < 2; ++i) Grow(&d_seg_[i], &seg_cap_[i], static_cast<size_t>(seg_cap_chunks) * std::max(max_seg, 1) * 4);
  int32_t max_kp = 0;
  for (const SegmentOp& so : seg_ops_)
    if (so.tensor) max_kp = std::max(max_kp, so.kp);
  if (max_kp > 0) {
    const int64_t split_rows = (seg_cap_chunks + 2 * kGemmBlockM - 1) / (2 * kGemmBlockM) * (2 * kGemmBlockM);
    Grow(&d_seg_hi_, &seg_hi_cap_, static_cast<size_t>(split_rows) * max_kp * 2);
    Grow(&d_seg_lo_, &seg_lo_cap_, static_cast<size_t>(split_rows) * max_kp * 2);
    seg_split_rows_ = split_rows;
  }
  Grow(&d_chunk_emb_, &chunk_emb_cap_, static_cast<size_t>(num_chunks) * E * 4);
  Grow(&d_out_, &out_cap_, static_cast<size_t>(num_utts) * E * 4);
  if (!dev_in)
    for (int i = 0; i < 2; ++i)
      for (int32_t s = 0; s < num_streams; ++s)
        Grow(&d_feat_[i][s], &d_feat_cap_[i][s], static_cast<size_t>(max_rows) * plan_.input_dims[s] * 4);

  const double ms_scratch = ms_since(t_enter);
  if (fused) {
    int32_t* rng = static_cast<int32_t*>(d_chunk_rng_);
    CVX_CUDA_CHECK(cudaMemcpyAsync(rng, chunk_lo.data(), static_cast<size_t>(num_chunks) * 4, cudaMemcpyHostToDevice, compute_));
    CVX_CUDA_CHECK(cudaMemcpyAsync(rng + num_chunks, chunk_hi.data(), static_cast<size_t>(num_chunks) * 4, cudaMemcpyHostToDevice, compute_));
    CVX_CUDA_CHECK(cudaMemcpyAsync(d_slot_begin_, slot_begin_packed.data(), slot_begin_packed.size() * 4, cudaMemcpyHostToDevice, compute_));
  } else {
    CVX_CUDA_CHECK(cudaMemcpyAsync(d_slabs_, slabs.data(), slabs.size() * sizeof(PoolSlab), cudaMemcpyHostToDevice, compute_));
    CVX_CUDA_CHECK(cudaMemcpyAsync(d_chunk_slab_off_, slab_off_packed.data(), slab_off_packed.size() * 4,
                                   cudaMemcpyHostToDevice, compute_));
  }
  CVX_CUDA_CHECK(cudaMemcpyAsync(d_weights_, weights.data(), weights.size() * 4, cudaMemcpyHostToDevice, compute_));
  CVX_CUDA_CHECK(cudaMemcpyAsync(d_utt_off_, utt_off.data(), utt_off.size() * 8, cudaMemcpyHostToDevice, compute_));
  // the pageable host vectors above must stay alive until these copies ran
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));

  // Results leave flush by flush: once the segment layers ran on a group of chunks, the utterances
  // whose chunks are all done are averaged (+ post-step) and copied out on their own stream while the
  // next batches compute -- with 10^5 short segments per call the embeddings are hundreds of MB.
  // Needs a destination an asynchronous copy can write: device memory or pinned host memory (a copy
  // into pageable memory would stall the enqueueing thread); otherwise everything is delivered at the end.
  bool stream_out = dev_out;
  if (!dev_out) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, out) == cudaSuccess) stream_out = attr.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  if (!d2h_) {
    CVX_CUDA_CHECK(cudaStreamCreateWithFlags(&d2h_, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CVX_CUDA_CHECK(cudaEventCreateWithFlags(&out_ready_[i], cudaEventDisableTiming));
      CVX_CUDA_CHECK(cudaEventCreateWithFlags(&out_free_[i], cudaEventDisableTiming));
    }
  }
  if (chain_on_) Grow(&d_chain_out_, &chain_out_cap_, static_cast<size_t>(num_utts) * chain_out_dim_ * 4);
  const int32_t OD = OutputDim();
  int32_t utt_done = 0;
  // average (+ post-step) utterances [u0, u1) on the compute stream, then hand them to the caller
  auto deliver = [&](int32_t u0, int32_t u1) {
    if (u1 <= u0) return;
    const int32_t n = u1 - u0;
    float* avg_dst = (dev_out && !chain_on_) ? out + static_cast<size_t>(u0) * E : static_cast<float*>(d_out_) + static_cast<size_t>(u0) * E;
    ProfBegin(kProfAverage, 0, static_cast<double>(n) * E * 4 * 2);
    CVX_CUDA_CHECK(LaunchUttAverage(static_cast<const float*>(d_chunk_emb_), E, static_cast<const int64_t*>(d_utt_off_) + u0,
                                    static_cast<const float*>(d_weights_), n, avg_dst, compute_));
    ProfEnd();
    const float* result = avg_dst;
    if (chain_on_) {
      // post-step on the device: the vectors leave as the back-end wants them
      float* dst = dev_out ? out + static_cast<size_t>(u0) * OD : static_cast<float*>(d_chain_out_) + static_cast<size_t>(u0) * OD;
      ProfBegin(kProfAverage, 0, static_cast<double>(n) * (E + OD) * 4);
      CVX_CUDA_CHECK(LaunchVectorChain(avg_dst, n, E, d_chain_mean_, d_chain_tf_, chain_.transform_rows, chain_.transform_cols,
                                       chain_.normalize_length, chain_.scaleup, /*keep_zero=*/1, dst, nullptr, compute_));
      ProfEnd();
      result = dst;
    }
    if (!dev_out) {
      CVX_CUDA_CHECK(cudaEventRecord(out_ready_[0], compute_));
      CVX_CUDA_CHECK(cudaStreamWaitEvent(d2h_, out_ready_[0], 0));
      CVX_CUDA_CHECK(cudaMemcpyAsync(out + static_cast<size_t>(u0) * OD, result, static_cast<size_t>(n) * OD * 4,
                                     cudaMemcpyDeviceToHost, d2h_));
    }
  };

  // ---- batches ----
  const double ms_meta = ms_since(t_enter);
  CVX_CUDA_CHECK(cudaEventRecord(call_begin_, compute_));
  int64_t pending = 0, pending_c0 = 0;  // chunks whose statistics wait for the segment layers
  for (size_t bi = 0; bi < batches.size(); ++bi) {
    const Batch& b = batches[bi];
    const int64_t rows = b.row1 - b.row0;
    const int fb = static_cast<int>(bi & 1);
    const float* dev[kMaxASources] = {};
    int64_t src_min, src_max;
    if (dev_in) {
      if (before_batch_) before_batch_(b.row1);  // cvx_extract_raw: pre-step of the pieces this batch reads
      for (int32_t s = 0; s < num_streams; ++s) dev[s] = feats[s] + b.row0 * plan_.input_dims[s];
      src_min = -b.row0;
      src_max = total_rows - b.row0;
    } else {
      // H2D of this batch on the copy stream, overlapping the previous batch's compute
      if (bi >= 2) CVX_CUDA_CHECK(cudaStreamWaitEvent(copy_, feat_free_[fb], 0));
      for (int32_t s = 0; s < num_streams; ++s) {
        CVX_CUDA_CHECK(cudaMemcpyAsync(d_feat_[fb][s], feats[s] + b.row0 * plan_.input_dims[s],
                                       static_cast<size_t>(rows) * plan_.input_dims[s] * 4, cudaMemcpyHostToDevice, copy_));
        dev[s] = static_cast<const float*>(d_feat_[fb][s]);
      }
      CVX_CUDA_CHECK(cudaEventRecord(feat_ready_[fb], copy_));
      CVX_CUDA_CHECK(cudaStreamWaitEvent(compute_, feat_ready_[fb], 0));
      src_min = 0;
      src_max = rows;
    }
    if (fused) {
      // per-group run table and rows per slot of this batch, from the chunk ranges
      const int32_t* rng = static_cast<const int32_t*>(d_chunk_rng_);
      ProfBegin(kProfPoolFin, 0, static_cast<double>(batch_groups[bi]) * sizeof(PoolGroupMeta));
      CVX_CUDA_CHECK(LaunchPoolMeta(rng + b.c0, rng + num_chunks + b.c0,
                                    static_cast<const int32_t*>(d_slot_begin_) + slot_begin_pos[bi],
                                    static_cast<int32_t>(b.c1 - b.c0), batch_groups[bi],
                                    static_cast<PoolGroupMeta*>(d_pool_meta_) + meta_pos[bi],
                                    static_cast<int32_t*>(d_slot_n_) + slot_pos[bi], compute_));
      ProfEnd();
    }
    const int32_t batch_slots = fused ? slot_begin_packed[slot_begin_pos[bi] + (b.c1 - b.c0)] : 0;
    // sparse form (engine.cu, "offset forwarding") unless an utterance with chunks in this batch pools only a
    // handful of frames in total (a short tail chunk of a long utterance carries a tiny weight in its average)
    bool sparse = true;
    for (int64_t c = b.c0; c < b.c1 && sparse; ++c) sparse = utt_pooled[chunk_utt[c]] >= kSparseMinPooledFrames;
    RunFrames(dev, rows, src_min, src_max, fused ? static_cast<const PoolGroupMeta*>(d_pool_meta_) + meta_pos[bi] : nullptr,
              nullptr, 0, batch_slots, sparse);
    if (!dev_in) CVX_CUDA_CHECK(cudaEventRecord(feat_free_[fb], compute_));

    const int32_t nslabs = static_cast<int32_t>(b.slab1 - b.slab0);
    const int32_t nchunks = static_cast<int32_t>(b.c1 - b.c0);
    const PoolSlab* dsl = static_cast<const PoolSlab*>(d_slabs_) + b.slab0;
    double pooled_rows = 0;
    for (int64_t c = b.c0; c < b.c1; ++c) pooled_rows += chunk_hi[c] - chunk_lo[c] + 1;
    if (fused) {
      // the pooled layer's epilogue already left per-slot sums in d_partials_
      const int32_t bslots = slot_begin_packed[slot_begin_pos[bi] + nchunks];
      ProfBegin(kProfPoolFin, 0, static_cast<double>(bslots) * 2 * D * 4 + static_cast<double>(nchunks) * pooled * 4);
      CVX_CUDA_CHECK(LaunchStatsPoolFinalizeFused(static_cast<const float*>(d_partials_), pool_buf.ld,
                                                  static_cast<const int32_t*>(d_slot_begin_) + slot_begin_pos[bi],
                                                  static_cast<const int32_t*>(d_slot_n_) + slot_pos[bi], nchunks, bslots, D, 1,
                                                  plan_.pool_variance_floor,
                                                  static_cast<float*>(d_stats_) + pending * pooled, pooled, compute_));
      ProfEnd();
    } else {
      ProfBegin(kProfPool, 0, pooled_rows * D * ElemSize(elem_));
      CVX_CUDA_CHECK(LaunchStatsPoolPartial(pool_buf.ptr, elem_, pool_buf.ld, dsl, nslabs,
                                            static_cast<float*>(d_partials_), compute_));
      ProfEnd();
      ProfBegin(kProfPoolFin, 0, static_cast<double>(nslabs) * 2 * D * 4 + static_cast<double>(nchunks) * pooled * 4);
      CVX_CUDA_CHECK(LaunchStatsPoolFinalize(static_cast<const float*>(d_partials_), pool_buf.ld, dsl,
                                             static_cast<const int32_t*>(d_chunk_slab_off_) + slab_off_pos[bi], nchunks, D,
                                             plan_.pool_output_stddevs ? 1 : 0, plan_.pool_variance_floor,
                                             static_cast<float*>(d_stats_) + pending * pooled, pooled, compute_));
      ProfEnd();
    }

    pending += nchunks;
    const bool last_batch = bi + 1 == batches.size();
    if (last_batch || pending + (batches[bi + 1].c1 - batches[bi + 1].c0) > seg_cap_chunks) {
      const float* cur = static_cast<const float*>(d_stats_);
      int32_t cur_ld = pooled;
      float* emb = static_cast<float*>(d_chunk_emb_) + pending_c0 * E;
      if (seg_ops_.empty())
        CVX_CUDA_CHECK(cudaMemcpyAsync(emb, cur, static_cast<size_t>(pending) * E * 4, cudaMemcpyDeviceToDevice, compute_));
      for (size_t si = 0; si < seg_ops_.size(); ++si) {
        const SegmentOp& s = seg_ops_[si];
        const bool last = si + 1 == seg_ops_.size();
        float* dst = last ? emb : static_cast<float*>(d_seg_[si & 1]);
        const int32_t dst_ld = s.out_dim;
        ProfBegin(kProfSegment, 2.0 * pending * s.in_dim * s.out_dim, 0);
        RunSegmentOp(s, elem_, num_sms_, cur, cur_ld, pending, dst, dst_ld, d_seg_hi_, d_seg_lo_, seg_split_rows_, compute_);
        ProfEnd();
        cur = dst;
        cur_ld = dst_ld;
      }
      pending_c0 += pending;
      pending = 0;
      if (stream_out && !last_batch) {
        int32_t u1 = utt_done;
        while (u1 < num_utts && utt_off[u1 + 1] <= pending_c0) ++u1;
        deliver(utt_done, u1);
        utt_done = u1;
      }
    }
  }
  deliver(utt_done, num_utts);
  if (!dev_out) {  // the caller's buffer is complete when the call returns; the copy tail is inside the timed bracket
    CVX_CUDA_CHECK(cudaEventRecord(out_free_[0], d2h_));
    CVX_CUDA_CHECK(cudaStreamWaitEvent(compute_, out_free_[0], 0));
  }
  CVX_CUDA_CHECK(cudaEventRecord(call_end_, compute_));
  const double ms_enqueued = ms_since(t_enter);
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));
  CVX_CUDA_CHECK(cudaEventElapsedTime(&last_call_ms_, call_begin_, call_end_));
  ProfResolve();
  last_num_chunks_ = num_chunks;
  if (trace)
    std::fprintf(stderr,
                 "cvx trace: extract rows=%lld chunks=%lld batches=%zu | tables %.2f ms, device scratch %.2f, table upload + sync %.2f, "
                 "enqueue %.2f, wait %.2f | device %.2f ms, rows_cap %lld\n",
                 static_cast<long long>(total_rows), static_cast<long long>(num_chunks), batches.size(), ms_tables,
                 ms_scratch - ms_tables, ms_meta - ms_scratch, ms_enqueued - ms_meta, ms_since(t_enter) - ms_enqueued,
                 static_cast<double>(last_call_ms_), static_cast<long long>(rows_cap_));
}

void Engine::LastChunkEmbeddings(float* out, int64_t num_chunks) {
  CVX_CUDA_CHECK(cudaSetDevice(opts_.device));
  if (num_chunks != last_num_chunks_) throw InvalidArgument("num_chunks does not match the last cvx_extract call");
  CVX_CUDA_CHECK(cudaMemcpy(out, d_chunk_emb_, static_cast<size_t>(num_chunks) * plan_.EmbeddingDim() * 4,
                            cudaMemcpyDeviceToHost));
}


}  // namespace cvx
