// Device-side executor of a Plan: owns weights, activation buffers, TMA
// descriptors, streams; runs packed ragged batches of chunks.
#ifndef CVX_ENGINE_H_
#define CVX_ENGINE_H_

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cvx.h"
#include "kernels/kernels.h"
#include "plan.h"

namespace cvx {

struct CudaError : public std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct InvalidArgument : public std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define CVX_CUDA_CHECK(expr)                                                                         \
  do {                                                                                               \
    cudaError_t cvx_e__ = (expr);                                                                    \
    if (cvx_e__ != cudaSuccess)                                                                      \
      throw ::cvx::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(cvx_e__) + " (" + \
                             __FILE__ + ":" + std::to_string(__LINE__) + ")");                       \
  } while (0)

constexpr uint32_t kDebugForceTapPath = 1u;   // cvx_engine_options.reserved[0] bits
constexpr uint32_t kDebugSingleCtaGemm = 2u;  // use the 1-CTA tcgen05 kernel instead of the CTA-pair one
constexpr uint32_t kDebugNoHaloGemm = 4u;     // use the plain CTA-pair kernel instead of the tap-halo one

// Segment-level affine layer (tdnn6 and later): fp32 in, fp32 out.  On the tensor path the fp32
// input and weights are split into 16-bit hi/lo parts -- x = hi + lo, W = Wh + Wl,
// y = hi*Wh + lo*Wh + hi*Wl (~2^-22 relative) -- and run through the CTA-pair GEMM with its fp32
// epilogue (engine_segment.cu).
struct SegmentOp {
  int32_t in_dim = 0, out_dim = 0;
  float *d_w = nullptr, *d_b = nullptr, *d_scale = nullptr, *d_offset = nullptr;
  bool relu = false;
  bool tensor = false;
  int32_t kp = 0, n_pad = 0;
  void* d_wcat = nullptr;      // [n_pad, 3*kp] = [Wh | Wh | Wl]
  float2* d_post = nullptr;    // [n_pad]
  std::vector<KBlockDesc> kb;  // hi blocks, lo blocks, hi blocks
};
// Uploads (and, for 16-bit `elem`, splits) the layer's parameters on the current device.
void BuildSegmentOp(const Layer& L, int32_t elem, SegmentOp* op);
void FreeSegmentOp(SegmentOp* op) noexcept;
// out[m, out_dim] = post(in[m, in_dim] W^T); hi / lo: scratch of split_rows x kp 16-bit values each
// (split_rows = m rounded up to 256), only used when op.tensor.
void RunSegmentOp(const SegmentOp& op, int32_t elem, int num_sms, const float* in, int32_t ld_in, int64_t m, float* out,
                  int32_t ld_out, void* hi, void* lo, int64_t split_rows, cudaStream_t stream);
// Test entry (cvx_test_segment_affine): one segment layer on host data through the production path.
void TestSegmentAffine(int32_t device, int32_t precision, const float* in, int64_t m, int32_t k, const float* w,
                       const float* bias, int32_t relu, const float* scale, const float* offset, int32_t n, float* out);

class Engine {
 public:
  Engine(const Plan& plan, const cvx_engine_options& opts);
  ~Engine();
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;

  void Extract(const float* const* feats, int32_t num_streams, int64_t total_rows, const int64_t* chunk_start,
               const int32_t* chunk_len, const int32_t* chunk_utt, const float* chunk_weight, int64_t num_chunks,
               int32_t num_utts, float* out, uint32_t flags);
  void LastChunkEmbeddings(float* out, int64_t num_chunks);
  // Frame-output models (Plan::frame_out >= 0): the output layer at every frame, edges replicated.
  void ComputeFrames(const float* const* feats, int32_t num_streams, int64_t total_rows, const int64_t* utt_start,
                     int32_t num_utts, float* out, uint32_t flags);
  // Extract() preceded by the device pre-step (sliding CMN + frame selection) on raw features.
  void ExtractRaw(const float* const* raw, const cvx_prestep* pre, int32_t num_streams, int64_t total_raw_rows,
                  const int64_t* utt_raw_start, const int32_t* keep, const int64_t* utt_kept_start, int32_t num_utts,
                  const int64_t* chunk_start, const int32_t* chunk_len, const int32_t* chunk_utt,
                  const float* chunk_weight, int64_t num_chunks, float* out, uint32_t flags);
  // Test hook: the pre-step alone; `out` receives the selected, normalised rows of stream 0.
  void DebugPrestep(const float* raw, int32_t dim, const cvx_prestep& pre, int64_t total_raw_rows,
                    const int64_t* utt_raw_start, const int32_t* keep, const int64_t* utt_kept_start, int32_t num_utts,
                    float* out);
  // Test hook: run only the frame layers on `rows` packed rows and return the last
  // frame layer's activations as fp32 [rows, out_dim].
  void DebugForwardFrames(const float* const* feats, int64_t rows, float* out);
  // Test hook: average device time (ms) of the frame-layer GEMM launches over `iters`
  // repetitions of the frame pass on `rows` rows (CUDA events around each launch).
  double DebugBenchFrames(const float* const* feats, int64_t rows, int32_t iters);

  // Post-step applied to the utterance embeddings on the device before they leave Extract()
  // (subtract mean -> transform -> length normalisation); chain == NULL removes it.
  void SetVectorChain(const cvx_vector_chain* chain);
  int32_t OutputDim() const { return chain_on_ ? chain_out_dim_ : plan_.EmbeddingDim(); }

  // check mode (opts.reserved[1]): values clipped at the fp16 range limit since the engine was built, for
  // the staged inputs (index 0) and every frame layer's output (index 1 + layer); names[] optional
  int SaturationCounts(int64_t* counts, int32_t max_counts);
  int GetProfile(cvx_kernel_profile* recs, int32_t max_recs);
  void ResetProfile();
  int64_t launch_count() const { return launches_; }
  float last_call_ms() const { return last_call_ms_; }
  const Plan& plan() const { return plan_; }

 private:
  struct Buffer {
    void* ptr = nullptr;
    int32_t ld = 0;   // elements per row, multiple of 64
    int32_t dim = 0;  // logical columns
  };
  struct Staging {
    std::vector<StageTap> taps;
    int32_t buffer = -1;
    uint32_t* d_colmap = nullptr;
  };
  struct FrameOp {
    int32_t layer = 0;
    int32_t out_buffer = -1;
    std::vector<int32_t> srcs;  // buffer ids
    std::vector<KBlockDesc> kb;
    int32_t n_pad = 0, k_pad = 0;
    void* d_w = nullptr;        // [n_pad, k_pad] of elem
    float2* d_post = nullptr;        // [n_pad] {b', o}: dense form, outputs are r + o
    float2* d_post_sparse = nullptr; // [n_pad] offset-forwarding form (engine.cu): outputs are r, sources' offsets in b'
    TdnnGemmArgs targs;
    TdnnGemmHaloArgs hargs;
    bool halo_ok = false;  // the layer fits the tap-halo kernel's limits
  };
  enum ProfClass { kProfStage = 0, kProfGemm, kProfPool, kProfPoolFin, kProfSegment, kProfAverage, kNumProf };
  struct PendingEvent {
    int cls;
    int layer = -1;
    cudaEvent_t a, b;
    double flops, bytes;
  };

  void Build();             // the constructor's work
  void Release() noexcept;  // frees every device / stream / event resource (destructor, failed constructor)
  int32_t StagingFor(const std::vector<Tap>& taps);
  void EnsureRows(int64_t rows);
  void EncodeMaps();
  void BuildHaloArgs(FrameOp& op);
  // pool_meta != NULL: the pooled layer runs with the fused-pooling epilogue (partials -> d_partials_)
  // sparse: run the offset-forwarding form (hidden activations stored without their BatchNorm offset)
  void RunFrames(const float* const* dev_feats, int64_t rows, int64_t src_row_min, int64_t src_row_max,
                 const PoolGroupMeta* pool_meta = nullptr, float* f32_out = nullptr, int32_t f32_ld = 0,
                 int32_t pool_slots = 0, bool sparse = false);
  void ProfBegin(int cls, double flops, double bytes, int layer = -1);
  void ProfEnd();
  void ProfResolve();
  template <typename T>
  T* DevAlloc(size_t n);
  void Grow(void** p, size_t* cap, size_t bytes);

  Plan plan_;
  cvx_engine_options opts_;
  int32_t elem_ = kF16;
  int num_sms_ = 148;
  uint32_t debug_ = 0;
  cudaStream_t compute_ = nullptr, copy_ = nullptr;
  cudaEvent_t feat_ready_[2] = {nullptr, nullptr}, feat_free_[2] = {nullptr, nullptr};

  std::vector<Buffer> buffers_;
  std::vector<Staging> stagings_;
  std::vector<FrameOp> frame_ops_;
  std::vector<SegmentOp> seg_ops_;
  std::vector<int32_t> layer_buffer_;
  int64_t rows_cap_ = 0;
  int64_t max_batch_rows_ = 0;

  // per-call scratch (grown on demand)
  void* d_feat_[2][kMaxASources] = {};
  size_t d_feat_cap_[2][kMaxASources] = {};
  void *d_slabs_ = nullptr, *d_chunk_slab_off_ = nullptr, *d_weights_ = nullptr, *d_utt_off_ = nullptr;
  size_t slabs_cap_ = 0, chunk_slab_off_cap_ = 0, weights_cap_ = 0, utt_off_cap_ = 0;
  void *d_partials_ = nullptr, *d_stats_ = nullptr, *d_seg_[2] = {nullptr, nullptr}, *d_chunk_emb_ = nullptr,
       *d_out_ = nullptr;
  size_t partials_cap_ = 0, stats_cap_ = 0, seg_cap_[2] = {0, 0}, chunk_emb_cap_ = 0, out_cap_ = 0;
  void *d_seg_hi_ = nullptr, *d_seg_lo_ = nullptr;  // 16-bit splits of a segment layer's fp32 input
  size_t seg_hi_cap_ = 0, seg_lo_cap_ = 0;
  int64_t seg_split_rows_ = 0;
  int64_t last_num_chunks_ = 0;
  bool fused_pool_ = false;  // pooled layer's activations are reduced in its GEMM epilogue
  bool forward_offsets_ = false;  // the sparse form exists (extraction plans; CVX_FORWARD_OFFSETS=0 turns it off)
  void *d_pool_meta_ = nullptr, *d_slot_n_ = nullptr, *d_slot_begin_ = nullptr, *d_chunk_rng_ = nullptr;
  size_t pool_meta_cap_ = 0, slot_n_cap_ = 0, slot_begin_cap_ = 0, chunk_rng_cap_ = 0;
  void RunPrestep(int32_t s, int32_t dim, const cvx_prestep& pre, int32_t u0, int32_t u1, int64_t run0, int64_t num_runs,
                  int64_t kept0, int64_t num_kept, bool has_keep);
  std::vector<int64_t> UploadPrestepIndex(const int64_t* utt_raw_start, const int32_t* keep, const int64_t* utt_kept_start,
                                          int32_t num_utts, const std::vector<int32_t>& kept_utt);
  int32_t pre_num_utts_ = 0;
  std::function<void(int64_t)> before_batch_;  // set by ExtractRaw: called with a device batch's last row + 1
  std::vector<cudaEvent_t> piece_events_;
  void *d_raw_[kMaxASources] = {}, *d_sel_[kMaxASources] = {};
  size_t raw_cap_[kMaxASources] = {}, sel_cap_[kMaxASources] = {};
  void *d_pre_aux_ = nullptr, *d_pre_utt_start_ = nullptr, *d_pre_kept_utt_ = nullptr, *d_pre_keep_ = nullptr;
  size_t pre_aux_cap_ = 0, pre_utt_start_cap_ = 0, pre_kept_utt_cap_ = 0, pre_keep_cap_ = 0;
  cudaEvent_t call_begin_ = nullptr, call_end_ = nullptr;
  float last_call_ms_ = 0.f;

  // frame-output mode: results leave on their own stream from two alternating buffers
  cudaStream_t d2h_ = nullptr;
  cudaEvent_t out_ready_[2] = {nullptr, nullptr}, out_free_[2] = {nullptr, nullptr};
  void* d_frame_out_[2] = {nullptr, nullptr};
  size_t frame_out_cap_[2] = {0, 0};

  bool chain_on_ = false;
  cvx_vector_chain chain_ = {};
  int32_t chain_out_dim_ = 0;
  float *d_chain_mean_ = nullptr, *d_chain_tf_ = nullptr;
  void* d_chain_out_ = nullptr;
  size_t chain_out_cap_ = 0;

  bool check_sat_ = false;
  unsigned long long* d_sat_ = nullptr;  // [1 + frame layers]
  std::vector<int64_t> sat_counts_;
  void CollectSaturation();               // after a synchronised call: device counters -> sat_counts_
  bool profile_ = false;
  std::vector<PendingEvent> pending_;
  std::vector<cudaEvent_t> event_pool_;
  double prof_ms_[kNumProf] = {}, prof_flops_[kNumProf] = {}, prof_bytes_[kNumProf] = {};
  int64_t prof_launches_[kNumProf] = {};
  std::vector<double> layer_ms_, layer_flops_;  // per frame layer (profile mode)
  std::vector<int64_t> layer_launches_;
  int64_t launches_ = 0;
};

template <typename T>
T* Engine::DevAlloc(size_t n) {
  void* p = nullptr;
  CVX_CUDA_CHECK(cudaMalloc(&p, (n > 0 ? n : 1) * sizeof(T)));
  return static_cast<T*>(p);
}

// Stand-alone post-step on host vectors (cvx_transform_vectors).
void TransformVectors(int32_t device, const float* in, int64_t num_vectors, int32_t dim, const cvx_vector_chain& chain,
                      float* out, float* ratios);
void ValidateVectorChain(const cvx_vector_chain& chain, int32_t dim);

// PLDA scoring of a trial list (cvx_plda_score).
void PldaScore(int32_t device, const cvx_plda& plda, const float* train, const int32_t* train_num_utts, int32_t num_train,
               const float* test, int32_t num_test, const int32_t* trial_train, const int32_t* trial_test, int64_t num_trials,
               float* scores);

// Stand-alone pooling run for tests (cvx_test_stats_pool).
void TestStatsPool(int32_t device, int32_t precision, const float* in, int64_t rows, int32_t dim, const int64_t* begin,
                   const int64_t* end, int64_t num_chunks, float variance_floor, float* out);

}  // namespace cvx
#endif  // CVX_ENGINE_H_
