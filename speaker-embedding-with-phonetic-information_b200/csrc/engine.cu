// Engine: lowers a Plan to device resources and runs the hot path.
//
// HBM layout (all row-major, one packed ragged batch of chunks at a time):
//   features     fp32 [rows, D_s] per input stream (device copy of the caller's rows)
//   staging      elem [rows_cap, roundup(sum tap dims, 64)]  spliced first-layer operand
//   activations  elem [rows_cap, roundup(N_l, 64)] per frame layer; pad columns are zero
//   weights      elem [roundup(N_l,64), K_pad] per layer, tap-major, each tap padded to the
//                source buffer's row pitch; BatchNorm scale folded into rows
//   stats        fp32 [chunks_in_batch, 2*D]   pooled mean | stddev
//   chunk_emb    fp32 [num_chunks, E]          per-chunk embeddings of the whole call
// Row r of every frame-level buffer is the same time step of the same chunk: edge
// rows of a chunk read their neighbours' rows and hold garbage that pooling never
// reads (a layer's valid rows are [start+left, start+len-right)).
#include "engine.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "engine_detail.h"

namespace cvx {

namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn GetEncodeTiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CVX_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !p) throw CudaError("cuTensorMapEncodeTiled is not available in this driver");
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D map over a row-major [rows, ld] matrix of 2-byte elements, box = box_cols x box_rows,
// 128-byte swizzle (box_cols * 2 bytes == 128), out-of-bounds reads give zeros.
}  // namespace

namespace detail {
void Encode2D(CUtensorMap* map, int32_t elem, void* base, int64_t rows, int64_t ld, uint32_t box_cols,
              uint32_t box_rows, int64_t cols, bool swizzle) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols < 0 ? ld : cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * ElemSize(elem)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : elem == kF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = GetEncodeTiled()(map, dt, 2,
                                base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
}

void ConvertToElem(const float* src, size_t n, int32_t elem, void* dst) {
  if (elem == kF32) {
    std::memcpy(dst, src, n * 4);
  } else if (elem == kF16) {
    __half* d = static_cast<__half*>(dst);
    for (size_t i = 0; i < n; ++i) d[i] = __float2half_rn(std::min(std::max(src[i], -65504.f), 65504.f));
  } else {
    __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dst);
    for (size_t i = 0; i < n; ++i) d[i] = __float2bfloat16_rn(src[i]);
  }
}

float ElemToFloat(const void* p, size_t i, int32_t elem) {
  if (elem == kF32) return static_cast<const float*>(p)[i];
  if (elem == kF16) return __half2float(static_cast<const __half*>(p)[i]);
  return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
}

}  // namespace detail

namespace {
const char* kProfNames[] = {"stage_input", "tdnn_gemm", "stats_pool_partial", "stats_pool_finalize",
                            "segment_affine", "utt_average"};

}  // namespace

using detail::ConvertToElem;
using detail::ElemToFloat;
using detail::Encode2D;
using detail::RoundUp;

void Engine::Grow(void** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return;
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));
  CVX_CUDA_CHECK(cudaStreamSynchronize(copy_));
  if (*p) CVX_CUDA_CHECK(cudaFree(*p));
  *p = nullptr;
  size_t want = bytes + bytes / 8;
  CVX_CUDA_CHECK(cudaMalloc(p, want));
  *cap = want;
}

Engine::Engine(const Plan& plan, const cvx_engine_options& opts) : plan_(plan), opts_(opts) {
  try {
    Build();
  } catch (...) {
    Release();  // a constructor that throws runs no destructor: free what was already allocated
    throw;
  }
}

void Engine::Build() {
  const cvx_engine_options& opts = opts_;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw CudaError(std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); this path has no CPU fallback");
  if (opts.device < 0 || opts.device >= ndev) throw InvalidArgument("device ordinal out of range");
  CVX_CUDA_CHECK(cudaSetDevice(opts.device));
  cudaDeviceProp prop;
  CVX_CUDA_CHECK(cudaGetDeviceProperties(&prop, opts.device));
  num_sms_ = prop.multiProcessorCount;
  if (opts.precision == CVX_PRECISION_FP16) elem_ = kF16;
  else if (opts.precision == CVX_PRECISION_BF16) elem_ = kBF16;
  else if (opts.precision == CVX_PRECISION_FP32) elem_ = kF32;
  else throw InvalidArgument("unknown precision");
  if (elem_ != kF32 && prop.major != 10)
    throw CudaError("the tcgen05 path needs an sm_100a device (Blackwell B200); found sm_" + std::to_string(prop.major) +
                    std::to_string(prop.minor));
  debug_ = static_cast<uint32_t>(opts.reserved[0]);
  if (const char* env = std::getenv("CVX_GEMM_HALO")) {
    if (env[0] == '0') debug_ |= kDebugNoHaloGemm;  // A/B aid: plain CTA-pair kernel
  }
  if (const char* env = std::getenv("CVX_INPUT_TAPS")) {
    if (env[0] == '1') debug_ |= kDebugForceTapPath;  // A/B aid: input layers read per-tap tiles instead of a compact splice
  }
  if (const char* env = std::getenv("CVX_GEMM_1CTA")) {
    if (env[0] == '1') debug_ |= kDebugSingleCtaGemm;  // A/B aid: same results, 1-CTA tiles
  }
  profile_ = opts.profile != 0;
  check_sat_ = opts.reserved[1] != 0 && elem_ == kF16;
  max_batch_rows_ = opts.max_batch_rows > 0 ? opts.max_batch_rows : (int64_t{1} << 19);
  max_batch_rows_ = (max_batch_rows_ + kGemmBlockM - 1) / kGemmBlockM * kGemmBlockM;

  CVX_CUDA_CHECK(cudaStreamCreateWithFlags(&compute_, cudaStreamNonBlocking));
  CVX_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CVX_CUDA_CHECK(cudaEventCreateWithFlags(&feat_ready_[i], cudaEventDisableTiming));
    CVX_CUDA_CHECK(cudaEventCreateWithFlags(&feat_free_[i], cudaEventDisableTiming));
  }
  CVX_CUDA_CHECK(cudaEventCreate(&call_begin_));
  CVX_CUDA_CHECK(cudaEventCreate(&call_end_));
  if (plan_.input_dims.size() > kMaxASources) throw InvalidArgument("too many input streams");

  // ---- lower frame layers ----
  // Offset forwarding ("sparse form"): a hidden layer's output is y = relu(.) + o (BatchNorm offset; its scale
  // is already folded into the layer's own weights).  Stored as y it is dense; stored as r = relu(.) about
  // half of it is exact zeros, and W (r + o) + b = W r + (W o + b): the offset moves into the CONSUMERS' biases
  // (computed once, in double), the activation operand of every hidden GEMM becomes half zeros, and on a
  // power-capped part a tensor core fed zeros draws less, i.e. clocks higher: +7 % on the headline workload
  // (same box, CVX_FORWARD_OFFSETS=0/1: 238.7-241.8 -> 223.5-224.7 ms per step).  Same algebra as Kaldi's
  // CollapseModel, applied forward; the weights are the same in both forms, only the {b', o} tables differ, so
  // the form is chosen per device batch.  r is larger in magnitude than r + o (o < 0 for a positive BatchNorm
  // mean), so its 16-bit rounding error is ~25 % larger: pooling averages that away at the bench's shapes
  // (2.0e-4 -> 2.5e-4 of the 1e-3 tolerance) but not for a chunk that pools a handful of frames, so a batch
  // holding a chunk of an utterance that pools fewer than kSparseMinPooledFrames frames in total runs the dense form, as do
  // frame-level outputs (nnet3-compute: no average at all).  Never for the pooled / output layer itself.
  forward_offsets_ = !plan_.FrameOutput();
  if (const char* env = std::getenv("CVX_FORWARD_OFFSETS")) forward_offsets_ = forward_offsets_ && env[0] != '0';
  const bool forward_offsets = forward_offsets_;
  std::vector<bool> forwarded(plan_.frame.size(), false);
  layer_buffer_.assign(plan_.frame.size(), -1);
  for (size_t li = 0; li < plan_.frame.size(); ++li) {
    const Layer& L = plan_.frame[li];
    forwarded[li] = forward_offsets && L.relu && !L.offset.empty() && static_cast<int32_t>(li) != plan_.pool_src &&
                    static_cast<int32_t>(li) != plan_.frame_out;
    FrameOp op;
    op.layer = static_cast<int32_t>(li);
    op.n_pad = RoundUp(L.out_dim, 64);
    bool all_inputs = true;
    for (const Tap& t : L.taps) all_inputs = all_inputs && t.src < 0;
    std::vector<std::pair<int32_t, int32_t>> col_map;  // (logical k begin, padded k begin) per tap; dims via plan
    std::vector<int32_t> tap_dims;
    if (all_inputs && !(debug_ & kDebugForceTapPath)) {
      // compact splice: one staging buffer holds Append(...) exactly as W's columns expect
      int32_t b = StagingFor(L.taps);
      op.srcs.push_back(b);
      op.k_pad = buffers_[b].ld;
      const int32_t k_real = RoundUp(L.in_dim, 16);
      for (int32_t j = 0; j * kGemmBlockK < k_real; ++j)
        op.kb.push_back({0, j * kGemmBlockK, 0, std::min(kGemmBlockK, k_real - j * kGemmBlockK) / 16});
      col_map.push_back({0, 0});
      tap_dims.push_back(L.in_dim);
    } else {
      int32_t klog = 0, kpad = 0;
      for (const Tap& t : L.taps) {
        int32_t b;
        if (t.src < 0) b = StagingFor({Tap{t.src, 0}});  // plain converted copy of the stream
        else b = layer_buffer_[t.src];
        int32_t si = -1;
        for (size_t s = 0; s < op.srcs.size(); ++s)
          if (op.srcs[s] == b) si = static_cast<int32_t>(s);
        if (si < 0) {
          if (op.srcs.size() >= kMaxASources) throw InvalidArgument("layer " + L.name + " appends more than 4 distinct sources");
          si = static_cast<int32_t>(op.srcs.size());
          op.srcs.push_back(b);
        }
        const int32_t d = buffers_[b].dim, dp = buffers_[b].ld;
        const int32_t d_real = RoundUp(d, 16);  // columns [d, d_real) of a source are exact zeros
        for (int32_t j = 0; j * kGemmBlockK < d_real; ++j)
          op.kb.push_back({si, j * kGemmBlockK, t.offset, std::min(kGemmBlockK, d_real - j * kGemmBlockK) / 16});
        col_map.push_back({klog, kpad});
        tap_dims.push_back(d);
        klog += d;
        kpad += dp;
      }
      op.k_pad = kpad;
    }
    if (op.kb.size() > kMaxKBlocks)
      throw InvalidArgument("layer " + L.name + ": padded K " + std::to_string(op.k_pad) + " exceeds " +
                            std::to_string(kMaxKBlocks * kGemmBlockK));

    // weights: fold BatchNorm scale into rows, re-pack columns, convert
    std::vector<float> wp(static_cast<size_t>(op.n_pad) * op.k_pad, 0.f);
    std::vector<float2> post(op.n_pad, make_float2(0.f, 0.f)), post_sparse(op.n_pad, make_float2(0.f, 0.f));
    // bias correction for sources that store r instead of r + o: sum over their taps of W[:, tap columns] o_src
    std::vector<double> fwd(L.out_dim, 0.0);
    {
      int32_t klog = 0;
      for (const Tap& t : L.taps) {
        const int32_t d = plan_.SrcDim(t.src);
        if (t.src >= 0 && forwarded[t.src]) {
          const std::vector<float>& o = plan_.frame[t.src].offset;
          for (int32_t n = 0; n < L.out_dim; ++n) {
            const float* w = &L.weight[static_cast<size_t>(n) * L.in_dim + klog];
            double acc = 0.0;
            for (int32_t c = 0; c < d; ++c) acc += static_cast<double>(w[c]) * o[c];
            fwd[n] += acc;
          }
        }
        klog += d;
      }
    }
    for (int32_t n = 0; n < L.out_dim; ++n) {
      float s = L.scale.empty() ? 1.f : L.scale[n];
      if (L.relu && !(s > 0.f))
        throw InvalidArgument("layer " + L.name + ": BatchNorm scale <= 0 after ReLU cannot be folded");
      for (size_t t = 0; t < col_map.size(); ++t)
        for (int32_t c = 0; c < tap_dims[t]; ++c)
          wp[static_cast<size_t>(n) * op.k_pad + col_map[t].second + c] =
              L.weight[static_cast<size_t>(n) * L.in_dim + col_map[t].first + c] * s;
      const double b = L.bias.empty() ? 0.0 : static_cast<double>(L.bias[n]);
      post[n] = make_float2(static_cast<float>(b * s), L.offset.empty() ? 0.f : L.offset[n]);
      post_sparse[n] = make_float2(static_cast<float>((b + fwd[n]) * s), (L.offset.empty() || forwarded[li]) ? 0.f : L.offset[n]);
    }
    std::vector<uint8_t> wconv(wp.size() * ElemSize(elem_));
    ConvertToElem(wp.data(), wp.size(), elem_, wconv.data());
    CVX_CUDA_CHECK(cudaMalloc(&op.d_w, wconv.size()));
    CVX_CUDA_CHECK(cudaMemcpy(op.d_w, wconv.data(), wconv.size(), cudaMemcpyHostToDevice));
    op.d_post = DevAlloc<float2>(post.size());
    CVX_CUDA_CHECK(cudaMemcpy(op.d_post, post.data(), post.size() * sizeof(float2), cudaMemcpyHostToDevice));
    op.d_post_sparse = DevAlloc<float2>(post_sparse.size());
    CVX_CUDA_CHECK(cudaMemcpy(op.d_post_sparse, post_sparse.data(), post_sparse.size() * sizeof(float2), cudaMemcpyHostToDevice));

    Buffer ob;
    ob.ld = op.n_pad;
    ob.dim = L.out_dim;
    op.out_buffer = static_cast<int32_t>(buffers_.size());
    buffers_.push_back(ob);
    layer_buffer_[li] = op.out_buffer;
    std::memset(&op.targs, 0, sizeof(op.targs));
    frame_ops_.push_back(std::move(op));
  }

  // ---- segment layers (fp32 in / out; hi-lo split on the tensor path) ----
  for (const Layer& L : plan_.segment) {
    seg_ops_.emplace_back();
    BuildSegmentOp(L, elem_, &seg_ops_.back());
  }

  // Fused pooling: possible when the pooled layer feeds nothing but the pooling, runs on the plain
  // CTA-pair kernel (single-tap sources) and the tensor path is in use.  CVX_FUSED_POOL=0 keeps the
  // stand-alone pooling kernels (whose slabs are cut relative to each chunk, so per-chunk results
  // do not depend on where a chunk sits in the batch).
  fused_pool_ = !plan_.FrameOutput() && elem_ != kF32 && !(debug_ & kDebugSingleCtaGemm) && plan_.pool_output_stddevs;
  for (const Layer& L : plan_.frame)
    for (const Tap& t : L.taps) fused_pool_ = fused_pool_ && t.src != plan_.pool_src;
  if (!plan_.FrameOutput())
    for (const Tap& t : plan_.frame[plan_.pool_src].taps) fused_pool_ = fused_pool_ && t.offset == 0;
  if (const char* env = std::getenv("CVX_FUSED_POOL")) fused_pool_ = fused_pool_ && env[0] != '0';
  if (check_sat_) {
    fused_pool_ = false;  // every layer's output must exist in memory to be inspected
    d_sat_ = DevAlloc<unsigned long long>(1 + plan_.frame.size());
    CVX_CUDA_CHECK(cudaMemset(d_sat_, 0, (1 + plan_.frame.size()) * sizeof(unsigned long long)));
    sat_counts_.assign(1 + plan_.frame.size(), 0);
  }
}

void Engine::CollectSaturation() {
  if (!check_sat_) return;
  std::vector<unsigned long long> h(sat_counts_.size());
  CVX_CUDA_CHECK(cudaMemcpy(h.data(), d_sat_, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < h.size(); ++i) sat_counts_[i] = static_cast<int64_t>(h[i]);
}

int Engine::SaturationCounts(int64_t* counts, int32_t max_counts) {
  if (!check_sat_) return 0;
  CVX_CUDA_CHECK(cudaSetDevice(opts_.device));
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));
  CollectSaturation();
  const int n = std::min<int>(max_counts, static_cast<int>(sat_counts_.size()));
  for (int i = 0; i < n; ++i) counts[i] = sat_counts_[i];
  return n;
}

Engine::~Engine() { Release(); }

void Engine::Release() noexcept {
  if (cudaSetDevice(opts_.device) != cudaSuccess) return;  // no usable device: nothing was allocated
  cudaDeviceSynchronize();
  for (Buffer& b : buffers_)
    if (b.ptr) cudaFree(b.ptr);
  cudaFree(d_chain_mean_);
  cudaFree(d_chain_tf_);
  cudaFree(d_chain_out_);
  cudaFree(d_sat_);
  cudaFree(d_frame_out_[0]);
  cudaFree(d_frame_out_[1]);
  if (d2h_) {
    cudaStreamDestroy(d2h_);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(out_ready_[i]);
      cudaEventDestroy(out_free_[i]);
    }
  }
  for (FrameOp& op : frame_ops_) {
    cudaFree(op.d_w);
    cudaFree(op.d_post);
    cudaFree(op.d_post_sparse);
  }
  for (Staging& s : stagings_) cudaFree(s.d_colmap);
  for (int s = 0; s < kMaxASources; ++s) {
    cudaFree(d_raw_[s]);
    cudaFree(d_sel_[s]);
  }
  void* pre_scratch[] = {d_pre_aux_, d_pre_utt_start_, d_pre_kept_utt_, d_pre_keep_};
  for (void* p : pre_scratch) cudaFree(p);
  cudaFree(d_pool_meta_);
  cudaFree(d_slot_n_);
  cudaFree(d_slot_begin_);
  cudaFree(d_chunk_rng_);
  for (cudaEvent_t ev : piece_events_) cudaEventDestroy(ev);
  piece_events_.clear();
  cudaFree(d_seg_hi_);
  cudaFree(d_seg_lo_);
  for (SegmentOp& s : seg_ops_) FreeSegmentOp(&s);
  for (int i = 0; i < 2; ++i)
    for (int s = 0; s < kMaxASources; ++s) cudaFree(d_feat_[i][s]);
  void* scratch[] = {d_slabs_, d_chunk_slab_off_, d_weights_, d_utt_off_, d_partials_, d_stats_,
                     d_seg_[0], d_seg_[1], d_chunk_emb_, d_out_};
  for (void* p : scratch) cudaFree(p);
  for (cudaEvent_t ev : event_pool_) cudaEventDestroy(ev);
  for (PendingEvent& pe : pending_) {
    cudaEventDestroy(pe.a);
    cudaEventDestroy(pe.b);
  }
  for (int i = 0; i < 2; ++i) {
    if (feat_ready_[i]) cudaEventDestroy(feat_ready_[i]);
    if (feat_free_[i]) cudaEventDestroy(feat_free_[i]);
  }
  if (call_begin_) cudaEventDestroy(call_begin_);
  if (call_end_) cudaEventDestroy(call_end_);
  if (compute_) cudaStreamDestroy(compute_);
  if (copy_) cudaStreamDestroy(copy_);
}

int32_t Engine::StagingFor(const std::vector<Tap>& taps) {
  for (const Staging& s : stagings_) {
    if (s.taps.size() != taps.size()) continue;
    bool same = true;
    for (size_t i = 0; i < taps.size(); ++i)
      same = same && s.taps[i].stream == -1 - taps[i].src && s.taps[i].offset == taps[i].offset;
    if (same) return s.buffer;
  }
  if (taps.size() > kMaxStageTaps) throw InvalidArgument("first-layer splice has more than 16 taps");
  Staging s;
  int32_t col = 0;
  for (const Tap& t : taps) {
    StageTap st;
    st.stream = -1 - t.src;
    st.offset = t.offset;
    st.dim = plan_.input_dims[st.stream];
    st.col0 = col;
    col += st.dim;
    s.taps.push_back(st);
  }
  Buffer b;
  b.dim = col;
  b.ld = RoundUp(col, 64);
  std::vector<uint32_t> colmap(b.ld, 0u);
  for (const StageTap& st : s.taps) {
    if (st.offset < -kStageMaxOffset || st.offset > kStageMaxOffset || st.dim > 0xFFFF)
      throw InvalidArgument("input splice offset or dim out of range");
    for (int32_t c = 0; c < st.dim; ++c)
      colmap[st.col0 + c] = 0x80000000u | (static_cast<uint32_t>(st.stream) << 28) |
                            (static_cast<uint32_t>(st.offset + 2048) << 16) | static_cast<uint32_t>(c);
  }
  s.d_colmap = DevAlloc<uint32_t>(colmap.size());
  CVX_CUDA_CHECK(cudaMemcpy(s.d_colmap, colmap.data(), colmap.size() * 4, cudaMemcpyHostToDevice));
  s.buffer = static_cast<int32_t>(buffers_.size());
  buffers_.push_back(b);
  stagings_.push_back(s);
  return s.buffer;
}

void Engine::EnsureRows(int64_t rows) {
  int64_t need = (rows + kGemmBlockM - 1) / kGemmBlockM * kGemmBlockM;
  if (need <= rows_cap_) return;
  // A job of many calls cuts device batches of slightly different sizes just below max_batch_rows_: growing to
  // exactly `need` re-allocated every activation buffer (GBs: free + malloc + clear + new tensor maps) whenever
  // a batch beat the previous record.  Anything beyond a small job gets the full batch size at once.
  static const bool grow_exact = std::getenv("CVX_GROW_EXACT") != nullptr;  // A/B aid: the old policy
  if (need * 4 > max_batch_rows_ && !grow_exact) need = std::max(need, max_batch_rows_);
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));
  CVX_CUDA_CHECK(cudaStreamSynchronize(copy_));
  for (Buffer& b : buffers_) {
    if (b.ptr) CVX_CUDA_CHECK(cudaFree(b.ptr));
    b.ptr = nullptr;
  }
  rows_cap_ = need;
  for (Buffer& b : buffers_) {
    size_t bytes = static_cast<size_t>(rows_cap_) * b.ld * ElemSize(elem_);
    CVX_CUDA_CHECK(cudaMalloc(&b.ptr, bytes));
    CVX_CUDA_CHECK(cudaMemsetAsync(b.ptr, 0, bytes, compute_));
  }
  EncodeMaps();
}

void Engine::EncodeMaps() {
  if (elem_ == kF32) return;
  for (FrameOp& op : frame_ops_) {
    TdnnGemmArgs& a = op.targs;
    std::memset(&a, 0, sizeof(a));
    for (size_t s = 0; s < op.srcs.size(); ++s) {
      const Buffer& b = buffers_[op.srcs[s]];
      Encode2D(&a.a_maps[s], elem_, b.ptr, rows_cap_, b.ld, kGemmBlockK, kGemmBlockM);
    }
    for (size_t s = op.srcs.size(); s < kMaxASources; ++s) a.a_maps[s] = a.a_maps[0];
    a.b_box_rows = std::min(kGemmBlockN, op.n_pad);
    Encode2D(&a.b_map, elem_, op.d_w, op.n_pad, op.k_pad, kGemmBlockK, a.b_box_rows);
    const Buffer& ob = buffers_[op.out_buffer];
    // Columns the MMAs produce: out_dim rounded up to the instruction's granularity of 16 (650 -> 656, not 672:
    // 2.4 % fewer tensor-core cycles on the 650-wide layers).  The epilogue works in 32-column units, so the
    // last tile's final unit may hold 16 columns no MMA wrote: the store maps end at n_mma and clip them.
    // (The pooled layer keeps whole 32-column units for the fused pooling epilogue.)
    a.n_mma = RoundUp(plan_.frame[op.layer].out_dim, op.layer == plan_.pool_src ? 32 : 16);
    if (const char* env = std::getenv("CVX_N_GRANULE"))  // A/B aid: 32 restores whole units everywhere
      if (std::atoi(env) == 32) a.n_mma = RoundUp(plan_.frame[op.layer].out_dim, 32);
    Encode2D(&a.c_map, elem_, ob.ptr, rows_cap_, ob.ld, 64, 32, a.n_mma);
    Encode2D(&a.c_map_half, elem_, ob.ptr, rows_cap_, ob.ld, 32, 32, a.n_mma, /*swizzle=*/false);
    a.num_kb = static_cast<int32_t>(op.kb.size());
    for (size_t i = 0; i < op.kb.size(); ++i) a.kb[i] = op.kb[i];
    a.n_pad = op.n_pad;
    {  // equal N tiles (e.g. 672 = 3 x 224 instead of 256 + 256 + 160): narrow tiles are operand-read bound
      const int32_t nt = (a.n_mma + kGemmBlockN - 1) / kGemmBlockN;
      a.tile_n = RoundUp((a.n_mma + nt - 1) / nt, 32);
    }
    a.num_n_tiles = (a.n_mma + a.tile_n - 1) / a.tile_n;
    // each CTA of a pair stages half of a W tile: the box holds exactly tile_n / 2 rows
    a.b_box_rows_pair = a.tile_n / 2;
    Encode2D(&a.b_map_pair, elem_, op.d_w, op.n_pad, op.k_pad, kGemmBlockK, a.b_box_rows_pair);
    a.post = op.d_post;
    a.relu = plan_.frame[op.layer].relu ? 1 : 0;
    a.ab_format = elem_ == kBF16 ? 1 : 0;
    a.num_stages = 64;
    if (const char* env = std::getenv("CVX_GEMM_STAGES")) a.num_stages = std::max(1, std::atoi(env));
    // first layers (K <= 128): weight half-tiles stay in shared memory for the whole launch
    a.resident_b = a.num_kb <= 2 ? 1 : 0;
    if (const char* env = std::getenv("CVX_RESIDENT_W")) a.resident_b = a.resident_b && env[0] != '0';
    a.wide_epilogue = 1;
    if (const char* env = std::getenv("CVX_WIDE_EPILOGUE")) a.wide_epilogue = env[0] != '0';
    BuildHaloArgs(op);
  }
}

// Groups the layer's K blocks by (source, column block): taps of one source whose offsets span
// <= kMaxHaloSpan rows share one activation tile (tdnn_gemm_halo.cu).
void Engine::BuildHaloArgs(FrameOp& op) {
  const TdnnGemmArgs& a = op.targs;
  TdnnGemmHaloArgs& h = op.hargs;
  std::memset(&h, 0, sizeof(h));
  op.halo_ok = false;
  // blocks of one tap are consecutive in op.kb; taps appear in Append order
  struct TapRun { int32_t src, rowoff, first_kb, nblocks; };
  std::vector<TapRun> runs;
  for (size_t i = 0; i < op.kb.size(); ++i) {
    const KBlockDesc& d = op.kb[i];
    if (d.col == 0) runs.push_back({d.src, d.rowoff, static_cast<int32_t>(i), 0});
    if (runs.empty()) return;
    runs.back().nblocks++;
  }
  std::vector<HaloGroup> groups;
  std::vector<bool> used(runs.size(), false);
  int32_t span_max = 0;
  for (size_t r = 0; r < runs.size(); ++r) {
    if (used[r]) continue;
    // collect later runs of the same source with the same block count that fit the span / tap limits
    std::vector<size_t> members{r};
    int32_t lo = runs[r].rowoff, hi = runs[r].rowoff;
    for (size_t q = r + 1; q < runs.size() && members.size() < kMaxHaloTaps; ++q) {
      if (used[q] || runs[q].src != runs[r].src || runs[q].nblocks != runs[r].nblocks) continue;
      const int32_t nlo = std::min(lo, runs[q].rowoff), nhi = std::max(hi, runs[q].rowoff);
      if (nhi - nlo > kMaxHaloSpan) continue;
      lo = nlo;
      hi = nhi;
      members.push_back(q);
    }
    for (size_t m : members) used[m] = true;
    span_max = std::max(span_max, hi - lo);
    for (int32_t j = 0; j < runs[r].nblocks; ++j) {
      HaloGroup g;
      std::memset(&g, 0, sizeof(g));
      g.src = runs[r].src;
      g.col = j * kGemmBlockK;
      g.rowoff = lo;
      g.ntaps = static_cast<int32_t>(members.size());
      g.k16 = op.kb[runs[r].first_kb + j].k16;
      for (size_t t = 0; t < members.size(); ++t) {
        g.shift[t] = runs[members[t]].rowoff - lo;
        g.wkb[t] = runs[members[t]].first_kb + j;
      }
      groups.push_back(g);
    }
  }
  if (groups.size() > kMaxHaloGroups) return;
  h.num_groups = static_cast<int32_t>(groups.size());
  for (size_t i = 0; i < groups.size(); ++i) h.groups[i] = groups[i];
  h.a_box_rows = kGemmBlockM + span_max;
  for (size_t s = 0; s < op.srcs.size(); ++s) {
    const Buffer& b = buffers_[op.srcs[s]];
    Encode2D(&h.a_maps[s], elem_, b.ptr, rows_cap_, b.ld, kGemmBlockK, h.a_box_rows);
  }
  for (size_t s = op.srcs.size(); s < kMaxASources; ++s) h.a_maps[s] = h.a_maps[0];
  h.b_map_pair = a.b_map_pair;
  h.c_map = a.c_map;
  h.b_box_rows_pair = a.b_box_rows_pair;
  h.n_pad = a.n_pad;
  h.n_mma = a.n_mma;
  h.tile_n = a.tile_n;
  h.c_map_half = a.c_map_half;
  h.num_n_tiles = a.num_n_tiles;
  h.post = a.post;
  h.relu = a.relu;
  h.ab_format = a.ab_format;
  h.out_f32 = 0;
  if (const char* env = std::getenv("CVX_HALO_BASEOFF")) h.base_offset_mode = env[0] == '1';
  // single-tap layers gain nothing from the halo tile and prefer the deeper 6-stage ring of the
  // plain CTA-pair kernel
  bool multi = false;
  for (const HaloGroup& g : groups) multi = multi || g.ntaps > 1;
  op.halo_ok = multi;
}

// ------------------------------------------------------------- profile ----
void Engine::ProfBegin(int cls, double flops, double bytes, int layer) {
  ++launches_;
  if (!profile_) return;
  PendingEvent pe;
  pe.cls = cls;
  pe.layer = layer;
  pe.flops = flops;
  pe.bytes = bytes;
  for (cudaEvent_t* ev : {&pe.a, &pe.b}) {
    if (!event_pool_.empty()) {
      *ev = event_pool_.back();
      event_pool_.pop_back();
    } else {
      CVX_CUDA_CHECK(cudaEventCreate(ev));
    }
  }
  CVX_CUDA_CHECK(cudaEventRecord(pe.a, compute_));
  pending_.push_back(pe);
}

void Engine::ProfEnd() {
  if (!profile_) return;
  CVX_CUDA_CHECK(cudaEventRecord(pending_.back().b, compute_));
}

void Engine::ProfResolve() {
  if (!profile_) return;
  for (PendingEvent& pe : pending_) {
    float ms = 0.f;
    CVX_CUDA_CHECK(cudaEventElapsedTime(&ms, pe.a, pe.b));
    prof_ms_[pe.cls] += ms;
    prof_flops_[pe.cls] += pe.flops;
    prof_bytes_[pe.cls] += pe.bytes;
    prof_launches_[pe.cls] += 1;
    if (pe.layer >= 0) {
      if (layer_ms_.size() != plan_.frame.size()) {
        layer_ms_.assign(plan_.frame.size(), 0.0);
        layer_flops_.assign(plan_.frame.size(), 0.0);
        layer_launches_.assign(plan_.frame.size(), 0);
      }
      layer_ms_[pe.layer] += ms;
      layer_flops_[pe.layer] += pe.flops;
      layer_launches_[pe.layer] += 1;
    }
    event_pool_.push_back(pe.a);
    event_pool_.push_back(pe.b);
  }
  pending_.clear();
}

int Engine::GetProfile(cvx_kernel_profile* recs, int32_t max_recs) {
  int n = 0;
  for (int c = 0; c < kNumProf && n < max_recs; ++c) {
    std::memset(&recs[n], 0, sizeof(recs[n]));
    std::strncpy(recs[n].name, kProfNames[c], sizeof(recs[n].name) - 1);
    recs[n].launches = prof_launches_[c];
    recs[n].total_ms = prof_ms_[c];
    recs[n].flops = prof_flops_[c];
    recs[n].bytes = prof_bytes_[c];
    ++n;
  }
  // then one record per frame layer: "layer:<name>"
  for (size_t l = 0; l < layer_ms_.size() && n < max_recs; ++l) {
    std::memset(&recs[n], 0, sizeof(recs[n]));
    std::snprintf(recs[n].name, sizeof(recs[n].name), "layer:%s", plan_.frame[l].name.c_str());
    recs[n].launches = layer_launches_[l];
    recs[n].total_ms = layer_ms_[l];
    recs[n].flops = layer_flops_[l];
    ++n;
  }
  return n;
}

void Engine::ResetProfile() {
  for (int c = 0; c < kNumProf; ++c) prof_ms_[c] = prof_flops_[c] = prof_bytes_[c] = 0, prof_launches_[c] = 0;
  layer_ms_.clear();
  layer_flops_.clear();
  layer_launches_.clear();
}

// ---------------------------------------------------------- frame pass ----
void Engine::RunFrames(const float* const* dev_feats, int64_t rows, int64_t src_row_min, int64_t src_row_max,
                       const PoolGroupMeta* pool_meta, float* f32_out, int32_t f32_ld, int32_t pool_slots, bool sparse) {
  sparse = sparse && forward_offsets_;
  for (const Staging& s : stagings_) {
    StageArgs a;
    std::memset(&a, 0, sizeof(a));
    for (size_t i = 0; i < plan_.input_dims.size(); ++i) {
      a.feats[i] = dev_feats[i];
      a.feat_dim[i] = plan_.input_dims[i];
    }
    a.colmap = s.d_colmap;
    // tiled fast path: one stream, |offset| <= 16, narrow rows (the shared-memory tile stays < 48 KB)
    a.tile_stream = -1;
    if (!s.taps.empty()) {
      int32_t lo = s.taps[0].offset, hi = s.taps[0].offset;
      bool one = true;
      for (const StageTap& t : s.taps) {
        one = one && t.stream == s.taps[0].stream;
        lo = std::min(lo, t.offset);
        hi = std::max(hi, t.offset);
      }
      if (one && lo >= -16 && hi <= 16 && plan_.input_dims[s.taps[0].stream] <= 64) {
        a.tile_stream = s.taps[0].stream;
        a.off_min = lo;
        a.off_max = hi;
      }
    }
    a.out = buffers_[s.buffer].ptr;
    a.ld_out = buffers_[s.buffer].ld;
    a.elem = elem_;
    a.rows = rows;
    a.src_row_min = src_row_min;
    a.src_row_max = src_row_max;
    ProfBegin(kProfStage, 0, static_cast<double>(rows) * (buffers_[s.buffer].dim * 4.0 + a.ld_out * ElemSize(elem_)));
    CVX_CUDA_CHECK(LaunchStageInput(a, compute_));
    ProfEnd();
    if (check_sat_) CVX_CUDA_CHECK(LaunchCountSaturated(a.out, rows, a.ld_out, buffers_[s.buffer].dim, d_sat_, compute_));
  }
  for (FrameOp& op : frame_ops_) {
    const Layer& L = plan_.frame[op.layer];
    const double flops = 2.0 * static_cast<double>(rows) * L.in_dim * L.out_dim;
    ProfBegin(kProfGemm, flops, 0, op.layer);
    if (elem_ == kF32) {
      TdnnGemmSimtArgs a;
      std::memset(&a, 0, sizeof(a));
      for (size_t s = 0; s < op.srcs.size(); ++s) {
        a.a_src[s] = static_cast<const float*>(buffers_[op.srcs[s]].ptr);
        a.a_ld[s] = buffers_[op.srcs[s]].ld;
      }
      a.num_kb = static_cast<int32_t>(op.kb.size());
      for (size_t i = 0; i < op.kb.size(); ++i) a.kb[i] = op.kb[i];
      a.w = static_cast<const float*>(op.d_w);
      a.k_pad = op.k_pad;
      a.n_pad = op.n_pad;
      a.post = sparse ? op.d_post_sparse : op.d_post;
      a.relu = L.relu ? 1 : 0;
      a.out = static_cast<float*>(buffers_[op.out_buffer].ptr);
      a.ld_out = buffers_[op.out_buffer].ld;
      a.rows = rows;
      a.rows_cap = rows_cap_;
      CVX_CUDA_CHECK(LaunchTdnnGemmSimt(a, compute_));
    } else {
      op.targs.num_m_tiles = static_cast<int32_t>((rows + kGemmBlockM - 1) / kGemmBlockM);
      op.targs.post = op.hargs.post = sparse ? op.d_post_sparse : op.d_post;
      if (pool_meta && op.layer == plan_.pool_src) {
        TdnnGemmArgs fa = op.targs;  // same layer, epilogue reduces instead of storing
        fa.pool_meta = pool_meta;
        fa.pool_partials = static_cast<float*>(d_partials_);
        fa.pool_ld = buffers_[op.out_buffer].ld;
        fa.pool_slots = pool_slots;
        CVX_CUDA_CHECK(LaunchTdnnGemmPair(fa, num_sms_, compute_));
      } else if (f32_out && &op == &frame_ops_.back()) {
        // frame-output mode: the last layer's epilogue writes fp32 rows (no 16-bit rounding of the result);
        // rows beyond `rows` and columns beyond out_dim are clipped by the map
        TdnnGemmArgs fa = op.targs;
        Encode2D(&fa.c_map, kF32, f32_out, rows, f32_ld, 32, 32, L.out_dim);
        fa.c_map_half = fa.c_map;  // (tile widths stay the layer's own: the fp32 epilogue stores 32-column chunks)
        fa.out_f32 = 1;
        CVX_CUDA_CHECK(LaunchTdnnGemmPair(fa, num_sms_, compute_));
      } else if (debug_ & kDebugSingleCtaGemm) {
        CVX_CUDA_CHECK(LaunchTdnnGemm(op.targs, num_sms_, compute_));
      } else if (op.halo_ok && !(debug_ & kDebugNoHaloGemm)) {
        op.hargs.num_m_tiles = op.targs.num_m_tiles;
        CVX_CUDA_CHECK(LaunchTdnnGemmHalo(op.hargs, num_sms_, compute_));
      } else {
        CVX_CUDA_CHECK(LaunchTdnnGemmPair(op.targs, num_sms_, compute_));
      }
    }
    ProfEnd();
    if (check_sat_ && !(f32_out && &op == &frame_ops_.back()))
      CVX_CUDA_CHECK(LaunchCountSaturated(buffers_[op.out_buffer].ptr, rows, buffers_[op.out_buffer].ld, L.out_dim,
                                          d_sat_ + 1 + op.layer, compute_));
  }
}

void Engine::DebugForwardFrames(const float* const* feats, int64_t rows, float* out) {
  CVX_CUDA_CHECK(cudaSetDevice(opts_.device));
  if (rows <= 0) throw InvalidArgument("rows must be positive");
  EnsureRows(rows);
  const float* dev[kMaxASources] = {};
  for (size_t s = 0; s < plan_.input_dims.size(); ++s) {
    size_t bytes = static_cast<size_t>(rows) * plan_.input_dims[s] * 4;
    Grow(&d_feat_[0][s], &d_feat_cap_[0][s], bytes);
    CVX_CUDA_CHECK(cudaMemcpyAsync(d_feat_[0][s], feats[s], bytes, cudaMemcpyHostToDevice, compute_));
    dev[s] = static_cast<const float*>(d_feat_[0][s]);
  }
  RunFrames(dev, rows, 0, rows);
  const Buffer& b = buffers_[frame_ops_.back().out_buffer];
  std::vector<uint8_t> host(static_cast<size_t>(rows) * b.ld * ElemSize(elem_));
  CVX_CUDA_CHECK(cudaMemcpyAsync(host.data(), b.ptr, host.size(), cudaMemcpyDeviceToHost, compute_));
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));
  ProfResolve();
  for (int64_t r = 0; r < rows; ++r)
    for (int32_t c = 0; c < b.dim; ++c) out[r * b.dim + c] = ElemToFloat(host.data(), static_cast<size_t>(r) * b.ld + c, elem_);
}

double Engine::DebugBenchFrames(const float* const* feats, int64_t rows, int32_t iters) {
  CVX_CUDA_CHECK(cudaSetDevice(opts_.device));
  if (rows <= 0 || iters <= 0) throw InvalidArgument("rows and iters must be positive");
  EnsureRows(rows);
  const float* dev[kMaxASources] = {};
  for (size_t s = 0; s < plan_.input_dims.size(); ++s) {
    size_t bytes = static_cast<size_t>(rows) * plan_.input_dims[s] * 4;
    Grow(&d_feat_[0][s], &d_feat_cap_[0][s], bytes);
    CVX_CUDA_CHECK(cudaMemcpyAsync(d_feat_[0][s], feats[s], bytes, cudaMemcpyHostToDevice, compute_));
    dev[s] = static_cast<const float*>(d_feat_[0][s]);
  }
  const bool was = profile_;
  profile_ = false;
  for (int i = 0; i < 3; ++i) RunFrames(dev, rows, 0, rows);  // warm-up
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));
  profile_ = true;
  ResetProfile();
  for (int i = 0; i < iters; ++i) RunFrames(dev, rows, 0, rows);
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));
  ProfResolve();
  const double ms = prof_ms_[kProfGemm] / static_cast<double>(std::max<int64_t>(prof_launches_[kProfGemm], 1));
  ResetProfile();
  profile_ = was;
  return ms;
}


}  // namespace cvx
