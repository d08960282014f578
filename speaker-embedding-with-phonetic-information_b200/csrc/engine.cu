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
void Encode2D(CUtensorMap* map, int32_t elem, void* base, int64_t rows, int64_t ld, uint32_t box_cols,
              uint32_t box_rows, int64_t cols = -1, bool swizzle = true) {
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

inline int32_t RoundUp(int32_t v, int32_t m) { return (v + m - 1) / m * m; }

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

const char* kProfNames[] = {"stage_input", "tdnn_gemm", "stats_pool_partial", "stats_pool_finalize",
                            "segment_affine", "utt_average"};

}  // namespace

template <typename T>
T* Engine::DevAlloc(size_t n) {
  void* p = nullptr;
  CVX_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  return static_cast<T*>(p);
}

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
  layer_buffer_.assign(plan_.frame.size(), -1);
  for (size_t li = 0; li < plan_.frame.size(); ++li) {
    const Layer& L = plan_.frame[li];
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
    std::vector<float2> post(op.n_pad, make_float2(0.f, 0.f));
    for (int32_t n = 0; n < L.out_dim; ++n) {
      float s = L.scale.empty() ? 1.f : L.scale[n];
      if (L.relu && !(s > 0.f))
        throw InvalidArgument("layer " + L.name + ": BatchNorm scale <= 0 after ReLU cannot be folded");
      for (size_t t = 0; t < col_map.size(); ++t)
        for (int32_t c = 0; c < tap_dims[t]; ++c)
          wp[static_cast<size_t>(n) * op.k_pad + col_map[t].second + c] =
              L.weight[static_cast<size_t>(n) * L.in_dim + col_map[t].first + c] * s;
      post[n] = make_float2((L.bias.empty() ? 0.f : L.bias[n]) * s, L.offset.empty() ? 0.f : L.offset[n]);
    }
    std::vector<uint8_t> wconv(wp.size() * ElemSize(elem_));
    ConvertToElem(wp.data(), wp.size(), elem_, wconv.data());
    CVX_CUDA_CHECK(cudaMalloc(&op.d_w, wconv.size()));
    CVX_CUDA_CHECK(cudaMemcpy(op.d_w, wconv.data(), wconv.size(), cudaMemcpyHostToDevice));
    op.d_post = DevAlloc<float2>(post.size());
    CVX_CUDA_CHECK(cudaMemcpy(op.d_post, post.data(), post.size() * sizeof(float2), cudaMemcpyHostToDevice));

    Buffer ob;
    ob.ld = op.n_pad;
    ob.dim = L.out_dim;
    op.out_buffer = static_cast<int32_t>(buffers_.size());
    buffers_.push_back(ob);
    layer_buffer_[li] = op.out_buffer;
    std::memset(&op.targs, 0, sizeof(op.targs));
    frame_ops_.push_back(std::move(op));
  }

  // ---- segment layers (fp32) ----
  for (const Layer& L : plan_.segment) {
    SegOp s;
    s.in_dim = L.in_dim;
    s.out_dim = L.out_dim;
    s.relu = L.relu;
    s.d_w = DevAlloc<float>(L.weight.size());
    CVX_CUDA_CHECK(cudaMemcpy(s.d_w, L.weight.data(), L.weight.size() * 4, cudaMemcpyHostToDevice));
    if (!L.bias.empty()) {
      s.d_b = DevAlloc<float>(L.bias.size());
      CVX_CUDA_CHECK(cudaMemcpy(s.d_b, L.bias.data(), L.bias.size() * 4, cudaMemcpyHostToDevice));
    }
    if (!L.scale.empty()) {
      s.d_scale = DevAlloc<float>(L.scale.size());
      s.d_offset = DevAlloc<float>(L.offset.size());
      CVX_CUDA_CHECK(cudaMemcpy(s.d_scale, L.scale.data(), L.scale.size() * 4, cudaMemcpyHostToDevice));
      CVX_CUDA_CHECK(cudaMemcpy(s.d_offset, L.offset.data(), L.offset.size() * 4, cudaMemcpyHostToDevice));
    }
    bool scale_ok = true;
    for (float v : L.scale) scale_ok = scale_ok && (!L.relu || v > 0.f);
    s.kp = RoundUp(L.in_dim, kGemmBlockK);
    if (elem_ != kF32 && L.out_dim % 4 == 0 && scale_ok && 3 * (s.kp / kGemmBlockK) <= kMaxKBlocks) {
      s.tensor = true;
      s.n_pad = RoundUp(L.out_dim, 64);
      const size_t kc = static_cast<size_t>(3) * s.kp;
      std::vector<float> whi(static_cast<size_t>(s.n_pad) * kc, 0.f), wlo_src(static_cast<size_t>(s.n_pad) * s.kp, 0.f);
      std::vector<float2> post(s.n_pad, make_float2(0.f, 0.f));
      for (int32_t n = 0; n < L.out_dim; ++n) {
        const float sc = L.scale.empty() ? 1.f : L.scale[n];
        for (int32_t k = 0; k < L.in_dim; ++k) wlo_src[static_cast<size_t>(n) * s.kp + k] = L.weight[static_cast<size_t>(n) * L.in_dim + k] * sc;
        post[n] = make_float2((L.bias.empty() ? 0.f : L.bias[n]) * sc, L.offset.empty() ? 0.f : L.offset[n]);
      }
      // Wh = round16(W), Wl = round16(W - Wh)
      std::vector<uint8_t> h16(wlo_src.size() * 2), l16(wlo_src.size() * 2);
      ConvertToElem(wlo_src.data(), wlo_src.size(), elem_, h16.data());
      std::vector<float> resid(wlo_src.size());
      for (size_t i = 0; i < resid.size(); ++i) resid[i] = wlo_src[i] - ElemToFloat(h16.data(), i, elem_);
      ConvertToElem(resid.data(), resid.size(), elem_, l16.data());
      std::vector<uint8_t> wcat(static_cast<size_t>(s.n_pad) * kc * 2);
      for (int32_t n = 0; n < s.n_pad; ++n) {
        uint8_t* row = wcat.data() + static_cast<size_t>(n) * kc * 2;
        std::memcpy(row, h16.data() + static_cast<size_t>(n) * s.kp * 2, static_cast<size_t>(s.kp) * 2);
        std::memcpy(row + static_cast<size_t>(s.kp) * 2, h16.data() + static_cast<size_t>(n) * s.kp * 2, static_cast<size_t>(s.kp) * 2);
        std::memcpy(row + static_cast<size_t>(s.kp) * 4, l16.data() + static_cast<size_t>(n) * s.kp * 2, static_cast<size_t>(s.kp) * 2);
      }
      CVX_CUDA_CHECK(cudaMalloc(&s.d_wcat, wcat.size()));
      CVX_CUDA_CHECK(cudaMemcpy(s.d_wcat, wcat.data(), wcat.size(), cudaMemcpyHostToDevice));
      s.d_post = DevAlloc<float2>(post.size());
      CVX_CUDA_CHECK(cudaMemcpy(s.d_post, post.data(), post.size() * sizeof(float2), cudaMemcpyHostToDevice));
      const int32_t k_real = RoundUp(L.in_dim, 16);
      for (int pass = 0; pass < 3; ++pass)
        for (int32_t j = 0; j * kGemmBlockK < k_real; ++j)
          s.kb.push_back({pass == 1 ? 1 : 0, j * kGemmBlockK, 0, std::min(kGemmBlockK, k_real - j * kGemmBlockK) / 16});
      // every pass must cover the same kp/64 blocks so that W's column offset is kb*64
      if (static_cast<int32_t>(s.kb.size()) != 3 * (s.kp / kGemmBlockK)) s.tensor = false;
    }
    seg_ops_.push_back(s);
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
  for (SegOp& s : seg_ops_) {
    cudaFree(s.d_wcat);
    cudaFree(s.d_post);
    cudaFree(s.d_w);
    cudaFree(s.d_b);
    cudaFree(s.d_scale);
    cudaFree(s.d_offset);
  }
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
    Encode2D(&a.c_map, elem_, ob.ptr, rows_cap_, ob.ld, 64, 32);
    Encode2D(&a.c_map_half, elem_, ob.ptr, rows_cap_, ob.ld, 32, 32, -1, /*swizzle=*/false);
    a.num_kb = static_cast<int32_t>(op.kb.size());
    for (size_t i = 0; i < op.kb.size(); ++i) a.kb[i] = op.kb[i];
    a.n_pad = op.n_pad;
    a.n_mma = RoundUp(plan_.frame[op.layer].out_dim, 32);
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
                       const PoolGroupMeta* pool_meta, float* f32_out, int32_t f32_ld, int32_t pool_slots) {
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
      a.post = op.d_post;
      a.relu = L.relu ? 1 : 0;
      a.out = static_cast<float*>(buffers_[op.out_buffer].ptr);
      a.ld_out = buffers_[op.out_buffer].ld;
      a.rows = rows;
      a.rows_cap = rows_cap_;
      CVX_CUDA_CHECK(LaunchTdnnGemmSimt(a, compute_));
    } else {
      op.targs.num_m_tiles = static_cast<int32_t>((rows + kGemmBlockM - 1) / kGemmBlockM);
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

// -------------------------------------------------------------- extract ----
void Engine::Extract(const float* const* feats, int32_t num_streams, int64_t total_rows, const int64_t* chunk_start,
                     const int32_t* chunk_len, const int32_t* chunk_utt, const float* chunk_weight,
                     int64_t num_chunks, int32_t num_utts, float* out, uint32_t flags) {
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
  for (const SegOp& so : seg_ops_)
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

  // ---- batches ----
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
    RunFrames(dev, rows, src_min, src_max, fused ? static_cast<const PoolGroupMeta*>(d_pool_meta_) + meta_pos[bi] : nullptr,
              nullptr, 0, batch_slots);
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
        const SegOp& s = seg_ops_[si];
        const bool last = si + 1 == seg_ops_.size();
        float* dst = last ? emb : static_cast<float*>(d_seg_[si & 1]);
        const int32_t dst_ld = s.out_dim;
        ProfBegin(kProfSegment, 2.0 * pending * s.in_dim * s.out_dim, 0);
        if (s.tensor) {
          CVX_CUDA_CHECK(LaunchSplitF32(cur, cur_ld, pending, s.in_dim, d_seg_hi_, d_seg_lo_, s.kp, elem_, compute_));
          TdnnGemmArgs a;
          std::memset(&a, 0, sizeof(a));
          Encode2D(&a.a_maps[0], elem_, d_seg_hi_, seg_split_rows_, s.kp, kGemmBlockK, kGemmBlockM);
          Encode2D(&a.a_maps[1], elem_, d_seg_lo_, seg_split_rows_, s.kp, kGemmBlockK, kGemmBlockM);
          a.a_maps[2] = a.a_maps[3] = a.a_maps[0];
          a.b_box_rows = std::min(kGemmBlockN, s.n_pad);
          a.b_box_rows_pair = std::min(kGemmBlockN / 2, s.n_pad);
          Encode2D(&a.b_map, elem_, s.d_wcat, s.n_pad, 3 * s.kp, kGemmBlockK, a.b_box_rows);
          Encode2D(&a.b_map_pair, elem_, s.d_wcat, s.n_pad, 3 * s.kp, kGemmBlockK, a.b_box_rows_pair);
          // fp32 output; rows beyond `pending` and columns beyond out_dim are clipped by the map
          Encode2D(&a.c_map, kF32, dst, pending, dst_ld, 32, 32, s.out_dim);
          a.num_kb = static_cast<int32_t>(s.kb.size());
          for (size_t i = 0; i < s.kb.size(); ++i) a.kb[i] = s.kb[i];
          a.n_pad = s.n_pad;
          a.n_mma = RoundUp(s.out_dim, 32);
          a.tile_n = kGemmBlockN;
          a.c_map_half = a.c_map;
          a.num_m_tiles = static_cast<int32_t>((pending + kGemmBlockM - 1) / kGemmBlockM);
          a.num_n_tiles = (a.n_mma + kGemmBlockN - 1) / kGemmBlockN;
          a.post = s.d_post;
          a.relu = s.relu ? 1 : 0;
          a.ab_format = elem_ == kBF16 ? 1 : 0;
          a.out_f32 = 1;
          a.num_stages = 64;
          CVX_CUDA_CHECK(LaunchTdnnGemmPair(a, num_sms_, compute_));
        } else {
          CVX_CUDA_CHECK(LaunchSegmentAffine(cur, cur_ld, pending, s.in_dim, s.d_w, s.d_b, s.relu ? 1 : 0, s.d_scale,
                                             s.d_offset, s.out_dim, dst, dst_ld, compute_));
        }
        ProfEnd();
        cur = dst;
        cur_ld = dst_ld;
      }
      pending_c0 += pending;
      pending = 0;
    }
  }
  ProfBegin(kProfAverage, 0, static_cast<double>(num_chunks + num_utts) * E * 4);
  CVX_CUDA_CHECK(LaunchUttAverage(static_cast<const float*>(d_chunk_emb_), E, static_cast<const int64_t*>(d_utt_off_),
                                  static_cast<const float*>(d_weights_), num_utts,
                                  dev_out && !chain_on_ ? out : static_cast<float*>(d_out_), compute_));
  ProfEnd();
  const void* result = d_out_;
  if (chain_on_) {
    // post-step on the device: the vectors leave as the back-end wants them
    Grow(&d_chain_out_, &chain_out_cap_, static_cast<size_t>(num_utts) * chain_out_dim_ * 4);
    float* dst = dev_out ? out : static_cast<float*>(d_chain_out_);
    ProfBegin(kProfAverage, 0, static_cast<double>(num_utts) * (E + chain_out_dim_) * 4);
    CVX_CUDA_CHECK(LaunchVectorChain(static_cast<const float*>(d_out_), num_utts, E, d_chain_mean_, d_chain_tf_,
                                     chain_.transform_rows, chain_.transform_cols, chain_.normalize_length,
                                     chain_.scaleup, /*keep_zero=*/1, dst, nullptr, compute_));
    ProfEnd();
    result = d_chain_out_;
  }
  if (!dev_out)
    CVX_CUDA_CHECK(cudaMemcpyAsync(out, result, static_cast<size_t>(num_utts) * OutputDim() * 4, cudaMemcpyDeviceToHost, compute_));
  CVX_CUDA_CHECK(cudaEventRecord(call_end_, compute_));
  CVX_CUDA_CHECK(cudaStreamSynchronize(compute_));
  CVX_CUDA_CHECK(cudaEventElapsedTime(&last_call_ms_, call_begin_, call_end_));
  ProfResolve();
  last_num_chunks_ = num_chunks;
}

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

void Engine::LastChunkEmbeddings(float* out, int64_t num_chunks) {
  CVX_CUDA_CHECK(cudaSetDevice(opts_.device));
  if (num_chunks != last_num_chunks_) throw InvalidArgument("num_chunks does not match the last cvx_extract call");
  CVX_CUDA_CHECK(cudaMemcpy(out, d_chunk_emb_, static_cast<size_t>(num_chunks) * plan_.EmbeddingDim() * 4,
                            cudaMemcpyDeviceToHost));
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
