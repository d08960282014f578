// Internal form of cvx_plan (include/cvx.h): the fused-layer program the
// engine executes.  Produced either by the nnet3 loader (nnet3_model.cc) or
// from a caller-built cvx_plan (cvx_api.cc).
#ifndef CVX_PLAN_H_
#define CVX_PLAN_H_

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace cvx {

// The graph uses something the packed-row execution model cannot represent (-> CVX_ERR_UNSUPPORTED).
struct UnsupportedError : public std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct Tap {
  int32_t src;     // >= 0: frame layer index; < 0: input stream (-1 - src)
  int32_t offset;  // value at t is the source at t + offset
};

struct Layer {
  std::string name;
  std::vector<Tap> taps;        // Append(...) order = column order of W
  int32_t in_dim = 0, out_dim = 0;
  std::vector<float> weight;    // [out_dim, in_dim]
  std::vector<float> bias;      // [out_dim] or empty
  bool relu = false;
  std::vector<float> scale, offset;  // test-mode BatchNorm, empty if none
  // frame t of this layer is computable iff left <= t <= T-1-right
  int32_t left = 0, right = 0;
};

struct Plan {
  std::vector<std::string> input_names;
  std::vector<int32_t> input_dims;
  std::vector<Layer> frame;     // topological order
  int32_t pool_src = -1;
  int32_t pool_left_context = 0, pool_right_context = 0;
  bool pool_output_stddevs = true;
  float pool_variance_floor = 1e-10f;
  std::vector<Layer> segment;   // chain; taps unused
  // Frame-output mode (nnet3-compute): no pooling; the result is frame layer
  // `frame_out` at every frame of the utterance, optionally through a row-wise log-softmax.
  int32_t frame_out = -1;
  bool frame_log_softmax = false;
  std::string description;

  int32_t SrcDim(int32_t src) const { return src >= 0 ? frame[src].out_dim : input_dims[-1 - src]; }
  bool FrameOutput() const { return frame_out >= 0; }
  int32_t PooledDim() const { return FrameOutput() ? 0 : frame[pool_src].out_dim * (pool_output_stddevs ? 2 : 1); }
  int32_t EmbeddingDim() const {
    if (FrameOutput()) return frame[frame_out].out_dim;
    return segment.empty() ? PooledDim() : segment.back().out_dim;
  }
  int64_t MacPerFrame() const;
  int64_t MacPerChunk() const;
  // Validates indices/dims, fills Layer::left/right and `description`; throws std::runtime_error, or
  // UnsupportedError for a graph with frames outside the input's time range (see plan.cc).
  void Finalize();
};

}  // namespace cvx
#endif  // CVX_PLAN_H_
