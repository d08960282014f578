// C ABI (include/cvx.h): exception firewall around the C++ host side.
#include <cuda_runtime.h>

#include <cstring>
#include <memory>
#include <new>
#include <string>

#include "../../include/cvx.h"
#include "api_error.h"
#include "engine.h"
#include "kaldi_io.h"
#include "nnet3_model.h"
#include "plan.h"

struct cvx_model {
  cvx::Plan plan;
};
struct cvx_engine {
  std::unique_ptr<cvx::Engine> impl;
};

namespace {

thread_local std::string g_last_error;

int Fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

template <typename F>
int Guard(F&& f) {
  try {
    f();
    g_last_error.clear();
    return CVX_OK;
  } catch (const cvx::UnsupportedError& e) {
    return Fail(CVX_ERR_UNSUPPORTED, e.what());
  } catch (const cvx::FormatError& e) {
    return Fail(CVX_ERR_FORMAT, e.what());
  } catch (const cvx::KaldiError& e) {
    return Fail(CVX_ERR_IO, e.what());
  } catch (const cvx::CudaError& e) {
    return Fail(CVX_ERR_CUDA, e.what());
  } catch (const cvx::InvalidArgument& e) {
    return Fail(CVX_ERR_INVALID, e.what());
  } catch (const std::bad_alloc&) {
    return Fail(CVX_ERR_NOMEM, "out of host memory");
  } catch (const std::exception& e) {
    return Fail(CVX_ERR_INVALID, e.what());
  } catch (...) {
    return Fail(CVX_ERR_INVALID, "unknown error");
  }
}

cvx::Layer CopyLayer(const cvx_layer& in, bool frame) {
  cvx::Layer L;
  L.name = in.name ? in.name : "";
  if (in.in_dim <= 0 || in.out_dim <= 0 || !in.weight) throw cvx::InvalidArgument("layer with bad dims or NULL weight");
  L.in_dim = in.in_dim;
  L.out_dim = in.out_dim;
  if (frame) {
    if (in.num_taps <= 0 || !in.taps) throw cvx::InvalidArgument("frame layer without taps");
    for (int32_t i = 0; i < in.num_taps; ++i) L.taps.push_back({in.taps[i].src, in.taps[i].offset});
  }
  L.weight.assign(in.weight, in.weight + static_cast<size_t>(in.in_dim) * in.out_dim);
  if (in.bias) L.bias.assign(in.bias, in.bias + in.out_dim);
  L.relu = in.relu != 0;
  if ((in.scale == nullptr) != (in.offset == nullptr)) throw cvx::InvalidArgument("scale and offset must come together");
  if (in.scale) {
    L.scale.assign(in.scale, in.scale + in.out_dim);
    L.offset.assign(in.offset, in.offset + in.out_dim);
  }
  return L;
}

}  // namespace

namespace cvx {
int SetApiError(int code, const std::string& msg) { return Fail(code, msg); }
}  // namespace cvx

extern "C" {

const char* cvx_last_error(void) { return g_last_error.c_str(); }
const char* cvx_version(void) { return "cvx-b200 0.1 (abi 1)"; }

int cvx_model_load(const char* rxfilename, const char* output_node, cvx_model** out) {
  if (!rxfilename || !out) return Fail(CVX_ERR_INVALID, "cvx_model_load: NULL argument");
  *out = nullptr;
  return Guard([&] {
    std::unique_ptr<cvx_model> m(new cvx_model);
    m->plan = cvx::LoadNnet3Plan(rxfilename, output_node ? output_node : "");
    *out = m.release();
  });
}

int cvx_model_load_frames(const char* rxfilename, const char* output_node, cvx_model** out) {
  if (!rxfilename || !out) return Fail(CVX_ERR_INVALID, "cvx_model_load_frames: NULL argument");
  *out = nullptr;
  return Guard([&] {
    std::unique_ptr<cvx_model> m(new cvx_model);
    m->plan = cvx::LoadNnet3Plan(rxfilename, output_node ? output_node : "", /*frame_output=*/true);
    *out = m.release();
  });
}

int cvx_model_from_plan(const cvx_plan* p, cvx_model** out) {
  if (!p || !out) return Fail(CVX_ERR_INVALID, "cvx_model_from_plan: NULL argument");
  *out = nullptr;
  return Guard([&] {
    std::unique_ptr<cvx_model> m(new cvx_model);
    cvx::Plan& plan = m->plan;
    if (p->num_inputs <= 0 || !p->input_dims) throw cvx::InvalidArgument("plan without inputs");
    for (int32_t i = 0; i < p->num_inputs; ++i) {
      plan.input_dims.push_back(p->input_dims[i]);
      plan.input_names.push_back(p->input_names && p->input_names[i] ? p->input_names[i] : "input" + std::to_string(i));
    }
    if (p->num_frame_layers <= 0 || !p->frame_layers) throw cvx::InvalidArgument("plan without frame layers");
    for (int32_t i = 0; i < p->num_frame_layers; ++i) plan.frame.push_back(CopyLayer(p->frame_layers[i], true));
    plan.pool_src = p->pool_src;
    plan.pool_left_context = p->pool_left_context;
    plan.pool_right_context = p->pool_right_context;
    plan.pool_output_stddevs = p->pool_output_stddevs != 0;
    plan.pool_variance_floor = p->pool_variance_floor;
    for (int32_t i = 0; i < p->num_segment_layers; ++i) plan.segment.push_back(CopyLayer(p->segment_layers[i], false));
    plan.Finalize();
    *out = m.release();
  });
}

void cvx_model_free(cvx_model* model) { delete model; }

int cvx_model_get_info(const cvx_model* model, cvx_model_info* info) {
  if (!model || !info) return Fail(CVX_ERR_INVALID, "cvx_model_get_info: NULL argument");
  const cvx::Plan& p = model->plan;
  info->num_inputs = static_cast<int32_t>(p.input_dims.size());
  info->embedding_dim = p.EmbeddingDim();
  const int32_t ctx_layer = p.FrameOutput() ? p.frame_out : p.pool_src;
  info->left_context = p.frame[ctx_layer].left;
  info->right_context = p.frame[ctx_layer].right;
  info->pooled_dim = p.PooledDim();
  info->num_frame_layers = static_cast<int32_t>(p.frame.size());
  info->num_segment_layers = static_cast<int32_t>(p.segment.size());
  info->mac_per_frame = p.MacPerFrame();
  info->mac_per_chunk = p.MacPerChunk();
  return CVX_OK;
}

const char* cvx_model_input_name(const cvx_model* model, int32_t i) {
  if (!model || i < 0 || i >= static_cast<int32_t>(model->plan.input_names.size())) return nullptr;
  return model->plan.input_names[i].c_str();
}

int32_t cvx_model_input_dim(const cvx_model* model, int32_t i) {
  if (!model || i < 0 || i >= static_cast<int32_t>(model->plan.input_dims.size())) return -1;
  return model->plan.input_dims[i];
}

const char* cvx_model_describe(const cvx_model* model) { return model ? model->plan.description.c_str() : nullptr; }

void cvx_engine_options_default(cvx_engine_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->precision = CVX_PRECISION_FP16;
}

int cvx_engine_create(const cvx_model* model, const cvx_engine_options* opts, cvx_engine** out) {
  if (!model || !out) return Fail(CVX_ERR_INVALID, "cvx_engine_create: NULL argument");
  *out = nullptr;
  return Guard([&] {
    cvx_engine_options o;
    if (opts) o = *opts;
    else cvx_engine_options_default(&o);
    std::unique_ptr<cvx_engine> e(new cvx_engine);
    e->impl.reset(new cvx::Engine(model->plan, o));
    *out = e.release();
  });
}

void cvx_engine_destroy(cvx_engine* engine) { delete engine; }

int cvx_extract(cvx_engine* engine, const float* const* feats, int32_t num_streams, int64_t total_rows,
                const int64_t* chunk_start, const int32_t* chunk_len, const int32_t* chunk_utt,
                const float* chunk_weight, int64_t num_chunks, int32_t num_utts, float* out, uint32_t flags) {
  if (!engine) return Fail(CVX_ERR_INVALID, "cvx_extract: NULL engine");
  if (num_chunks > 0 && (!chunk_start || !chunk_len || !chunk_utt)) return Fail(CVX_ERR_INVALID, "cvx_extract: NULL chunk arrays");
  return Guard([&] {
    engine->impl->Extract(feats, num_streams, total_rows, chunk_start, chunk_len, chunk_utt, chunk_weight, num_chunks,
                          num_utts, out, flags);
  });
}

int cvx_extract_raw(cvx_engine* engine, const float* const* raw_feats, const cvx_prestep* prestep, int32_t num_streams,
                    int64_t total_raw_rows, const int64_t* utt_raw_start, const int32_t* keep,
                    const int64_t* utt_kept_start, int32_t num_utts, const int64_t* chunk_start,
                    const int32_t* chunk_len, const int32_t* chunk_utt, const float* chunk_weight, int64_t num_chunks,
                    float* out, uint32_t flags) {
  if (!engine) return Fail(CVX_ERR_INVALID, "cvx_extract_raw: NULL engine");
  if (num_chunks > 0 && (!chunk_start || !chunk_len || !chunk_utt)) return Fail(CVX_ERR_INVALID, "cvx_extract_raw: NULL chunk arrays");
  return Guard([&] {
    engine->impl->ExtractRaw(raw_feats, prestep, num_streams, total_raw_rows, utt_raw_start, keep, utt_kept_start,
                             num_utts, chunk_start, chunk_len, chunk_utt, chunk_weight, num_chunks, out, flags);
  });
}

int cvx_engine_set_vector_chain(cvx_engine* engine, const cvx_vector_chain* chain) {
  if (!engine) return Fail(CVX_ERR_INVALID, "cvx_engine_set_vector_chain: NULL engine");
  return Guard([&] { engine->impl->SetVectorChain(chain); });
}

int32_t cvx_engine_output_dim(const cvx_engine* engine) { return engine ? engine->impl->OutputDim() : -1; }

int cvx_transform_vectors(int32_t device, const float* in, int64_t num_vectors, int32_t dim, const cvx_vector_chain* chain,
                          float* out, float* ratios) {
  if (!chain || (num_vectors > 0 && (!in || !out))) return Fail(CVX_ERR_INVALID, "cvx_transform_vectors: NULL argument");
  return Guard([&] { cvx::TransformVectors(device, in, num_vectors, dim, *chain, out, ratios); });
}

int cvx_plda_score(int32_t device, const cvx_plda* plda, const float* train, const int32_t* train_num_utts, int32_t num_train,
                   const float* test, int32_t num_test, const int32_t* trial_train, const int32_t* trial_test,
                   int64_t num_trials, float* scores) {
  if (!plda || !plda->mean || !plda->transform || !plda->psi) return Fail(CVX_ERR_INVALID, "cvx_plda_score: NULL model");
  if (num_trials > 0 && (!train || !test || !trial_train || !trial_test || !scores))
    return Fail(CVX_ERR_INVALID, "cvx_plda_score: NULL argument");
  return Guard([&] {
    cvx::PldaScore(device, *plda, train, train_num_utts, num_train, test, num_test, trial_train, trial_test, num_trials, scores);
  });
}

int cvx_compute_frames(cvx_engine* engine, const float* const* feats, int32_t num_streams, int64_t total_rows,
                       const int64_t* utt_start, int32_t num_utts, float* out, uint32_t flags) {
  if (!engine) return Fail(CVX_ERR_INVALID, "cvx_compute_frames: NULL engine");
  if (num_utts > 0 && (!feats || !utt_start || !out)) return Fail(CVX_ERR_INVALID, "cvx_compute_frames: NULL argument");
  return Guard([&] { engine->impl->ComputeFrames(feats, num_streams, total_rows, utt_start, num_utts, out, flags); });
}

int cvx_test_prestep(cvx_engine* engine, const float* raw, int32_t dim, const cvx_prestep* prestep,
                     int64_t total_raw_rows, const int64_t* utt_raw_start, const int32_t* keep,
                     const int64_t* utt_kept_start, int32_t num_utts, float* out) {
  if (!engine || !raw || !prestep || !out) return Fail(CVX_ERR_INVALID, "cvx_test_prestep: NULL argument");
  return Guard([&] {
    engine->impl->DebugPrestep(raw, dim, *prestep, total_raw_rows, utt_raw_start, keep, utt_kept_start, num_utts, out);
  });
}

int cvx_engine_last_chunk_embeddings(cvx_engine* engine, float* out, int64_t num_chunks) {
  if (!engine || !out) return Fail(CVX_ERR_INVALID, "cvx_engine_last_chunk_embeddings: NULL argument");
  return Guard([&] { engine->impl->LastChunkEmbeddings(out, num_chunks); });
}

int cvx_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return Fail(CVX_ERR_INVALID, "cvx_host_alloc: NULL argument");
  cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) return Fail(CVX_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  return CVX_OK;
}
void cvx_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}
int cvx_device_alloc(int32_t device, void** ptr, size_t bytes) {
  if (!ptr) return Fail(CVX_ERR_INVALID, "cvx_device_alloc: NULL argument");
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(ptr, bytes ? bytes : 1);
  if (e != cudaSuccess) return Fail(CVX_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  return CVX_OK;
}
void cvx_device_free(int32_t device, void* ptr) {
  if (ptr && cudaSetDevice(device) == cudaSuccess) cudaFree(ptr);
}
int cvx_memcpy_h2d(int32_t device, void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return Fail(CVX_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e));
  return CVX_OK;
}
int cvx_device_count(int32_t* count) {
  if (!count) return Fail(CVX_ERR_INVALID, "cvx_device_count: NULL argument");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  *count = (e == cudaSuccess) ? n : 0;
  if (e != cudaSuccess) return Fail(CVX_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  return CVX_OK;
}

int cvx_engine_get_profile(cvx_engine* engine, cvx_kernel_profile* records, int32_t max_records, int32_t* num_records) {
  if (!engine || !records || !num_records) return Fail(CVX_ERR_INVALID, "cvx_engine_get_profile: NULL argument");
  *num_records = engine->impl->GetProfile(records, max_records);
  return CVX_OK;
}
int cvx_engine_saturation_counts(cvx_engine* engine, int64_t* counts, int32_t max_counts, int32_t* num_counts) {
  return Guard([&] {
    if (!engine || !counts || !num_counts) throw cvx::InvalidArgument("cvx_engine_saturation_counts: NULL argument");
    *num_counts = engine->impl->SaturationCounts(counts, max_counts);
  });
}

int cvx_test_segment_affine(int32_t device, int32_t precision, const float* in, int64_t m, int32_t k, const float* w,
                            const float* bias, int32_t relu, const float* scale, const float* offset, int32_t n, float* out) {
  return Guard([&] { cvx::TestSegmentAffine(device, precision, in, m, k, w, bias, relu, scale, offset, n, out); });
}

void cvx_engine_reset_profile(cvx_engine* engine) {
  if (engine) engine->impl->ResetProfile();
}
int64_t cvx_engine_launch_count(const cvx_engine* engine) { return engine ? engine->impl->launch_count() : 0; }
double cvx_engine_last_extract_ms(const cvx_engine* engine) { return engine ? engine->impl->last_call_ms() : 0.0; }

int cvx_test_tdnn_layer(int32_t device, int32_t precision, const float* in, int64_t rows, int32_t in_dim,
                        const int32_t* offsets, int32_t num_offsets, const float* weight, const float* bias,
                        int32_t relu, const float* scale, const float* offset, int32_t out_dim, float* out) {
  if (!in || !offsets || !weight || !out || num_offsets <= 0) return Fail(CVX_ERR_INVALID, "cvx_test_tdnn_layer: bad argument");
  return Guard([&] {
    // A one-layer plan whose taps all read input stream 0; kDebugForceTapPath makes the
    // engine treat the stream like a previous layer's output (64-column K blocks per
    // tap, TMA row offsets) instead of the compact first-layer splice.
    cvx::Plan plan;
    plan.input_names = {"input"};
    plan.input_dims = {in_dim};
    cvx::Layer L;
    L.name = "test";
    for (int32_t i = 0; i < num_offsets; ++i) L.taps.push_back({-1, offsets[i]});
    L.in_dim = in_dim * num_offsets;
    L.out_dim = out_dim;
    L.weight.assign(weight, weight + static_cast<size_t>(L.in_dim) * out_dim);
    if (bias) L.bias.assign(bias, bias + out_dim);
    L.relu = relu != 0;
    if (scale && offset) {
      L.scale.assign(scale, scale + out_dim);
      L.offset.assign(offset, offset + out_dim);
    }
    plan.frame.push_back(L);
    plan.pool_src = 0;
    plan.pool_right_context = 1 << 30;
    plan.Finalize();
    cvx_engine_options o;
    cvx_engine_options_default(&o);
    o.device = device;
    o.precision = precision & 0xff;
    o.reserved[0] = (precision & 0x100) ? 0 : static_cast<int32_t>(cvx::kDebugForceTapPath);
    cvx::Engine eng(plan, o);
    const float* feats[1] = {in};
    eng.DebugForwardFrames(feats, rows, out);
  });
}

int cvx_test_tdnn_layer_bench(int32_t device, int32_t precision, int64_t rows, int32_t in_dim, const int32_t* offsets,
                              int32_t num_offsets, int32_t out_dim, int32_t iters, double* avg_ms) {
  if (!offsets || !avg_ms || num_offsets <= 0 || rows <= 0) return Fail(CVX_ERR_INVALID, "cvx_test_tdnn_layer_bench: bad argument");
  return Guard([&] {
    cvx::Plan plan;
    plan.input_names = {"input"};
    plan.input_dims = {in_dim};
    cvx::Layer L;
    L.name = "bench";
    for (int32_t i = 0; i < num_offsets; ++i) L.taps.push_back({-1, offsets[i]});
    L.in_dim = in_dim * num_offsets;
    L.out_dim = out_dim;
    L.weight.resize(static_cast<size_t>(L.in_dim) * out_dim);
    uint32_t s = 12345u;
    auto rnd = [&s]() {  // small LCG: values in [-0.5, 0.5)
      s = s * 1664525u + 1013904223u;
      return static_cast<float>(s >> 8) * (1.0f / 16777216.0f) - 0.5f;
    };
    for (float& w : L.weight) w = rnd() * 0.1f;
    L.bias.assign(out_dim, 0.1f);
    L.relu = true;
    L.scale.assign(out_dim, 1.0f);
    L.offset.assign(out_dim, -0.1f);
    plan.frame.push_back(L);
    plan.pool_src = 0;
    plan.pool_right_context = 1 << 30;
    plan.Finalize();
    cvx_engine_options o;
    cvx_engine_options_default(&o);
    o.device = device;
    o.precision = precision & 0xff;
    o.reserved[0] = (precision & 0x100) ? 0 : static_cast<int32_t>(cvx::kDebugForceTapPath);
    cvx::Engine eng(plan, o);
    std::vector<float> x(static_cast<size_t>(rows) * in_dim);
    for (float& v : x) v = rnd() * 4.0f;
    const float* feats[1] = {x.data()};
    *avg_ms = eng.DebugBenchFrames(feats, rows, iters);
  });
}

int cvx_test_stats_pool(int32_t device, int32_t precision, const float* in, int64_t rows, int32_t dim,
                        const int64_t* begin, const int64_t* end, int64_t num_chunks, float variance_floor,
                        float* out) {
  if (!in || !begin || !end || !out) return Fail(CVX_ERR_INVALID, "cvx_test_stats_pool: NULL argument");
  return Guard([&] { cvx::TestStatsPool(device, precision, in, rows, dim, begin, end, num_chunks, variance_floor, out); });
}

}  // extern "C"
