#include "plan.h"

#include <algorithm>
#include <climits>
#include <cstdint>
#include <sstream>
#include <stdexcept>

namespace cvx {

int64_t Plan::MacPerFrame() const {
  int64_t m = 0;
  for (const Layer& l : frame) m += static_cast<int64_t>(l.in_dim) * l.out_dim;
  return m;
}

int64_t Plan::MacPerChunk() const {
  int64_t m = 0;
  for (const Layer& l : segment) m += static_cast<int64_t>(l.in_dim) * l.out_dim;
  return m;
}

void Plan::Finalize() {
  auto fail = [](const std::string& m) { throw std::runtime_error("invalid plan: " + m); };
  if (input_dims.empty() || input_dims.size() != input_names.size()) fail("no inputs");
  for (int32_t d : input_dims)
    if (d <= 0) fail("input dim <= 0");
  if (frame.empty()) fail("no frame layers");
  const int32_t n_in = static_cast<int32_t>(input_dims.size());
  // nnet3's own contexts (a layer all of whose taps look forward is computable BEFORE t = 0, one that
  // only looks back AFTER t = T-1, and upstream computes such frames when a consumer asks for them)
  std::vector<int32_t> true_left(frame.size(), 0), true_right(frame.size(), 0);
  for (size_t i = 0; i < frame.size(); ++i) {
    Layer& l = frame[i];
    if (l.taps.empty()) fail("frame layer " + l.name + " has no taps");
    int32_t k = 0, left = 0, right = 0, tl = INT32_MIN, tr = INT32_MIN;
    for (const Tap& t : l.taps) {
      if (t.src >= static_cast<int32_t>(i) || t.src < -n_in) fail("frame layer " + l.name + ": bad tap source");
      k += SrcDim(t.src);
      int32_t sl = t.src >= 0 ? frame[t.src].left : 0, sr = t.src >= 0 ? frame[t.src].right : 0;
      // output t needs the source at t+offset: sl <= t+offset <= T-1-sr
      left = std::max(left, sl - t.offset);
      right = std::max(right, sr + t.offset);
      tl = std::max(tl, (t.src >= 0 ? true_left[t.src] : 0) - t.offset);
      tr = std::max(tr, (t.src >= 0 ? true_right[t.src] : 0) + t.offset);
    }
    true_left[i] = tl;
    true_right[i] = tr;
    // rows exist only for 0 <= t < T, so contexts never go negative
    l.left = std::max(left, 0);
    l.right = std::max(right, 0);
    if (k != l.in_dim) fail("frame layer " + l.name + ": taps give dim " + std::to_string(k) + ", in_dim " + std::to_string(l.in_dim));
    if (l.out_dim <= 0) fail("frame layer " + l.name + ": out_dim <= 0");
    if (l.weight.size() != static_cast<size_t>(l.in_dim) * l.out_dim) fail("frame layer " + l.name + ": weight size");
    if (!l.bias.empty() && static_cast<int32_t>(l.bias.size()) != l.out_dim) fail("frame layer " + l.name + ": bias size");
    if (l.scale.size() != l.offset.size() || (!l.scale.empty() && static_cast<int32_t>(l.scale.size()) != l.out_dim))
      fail("frame layer " + l.name + ": scale/offset size");
  }
  if (FrameOutput()) {
    if (frame_out != static_cast<int32_t>(frame.size()) - 1) fail("frame_out must be the last frame layer");
    if (pool_src >= 0 || !segment.empty()) fail("frame-output plans have no pooling or segment layers");
  } else {
    if (pool_src < 0 || pool_src >= static_cast<int32_t>(frame.size())) fail("pool_src out of range");
    if (pool_left_context < 0 || pool_right_context < 0) fail("negative pooling context");
    // The packed layout holds rows 0 <= t < T of every layer, so a layer's context is clamped at 0.  Nothing is
    // lost as long as the frames that reach the pooling are the ones upstream pools: the clamped chain must give
    // the pooled layer the context nnet3 itself computes (window [0, right-context]: a negative LEFT context is
    // harmless).  It does for every network of the recipes; it does not when a source whose taps all look
    // forward (computable before t = 0) is tapped again backwards, or when the pooled layer only looks back
    // (computable after the last input frame: upstream would pool those frames too).
    const Layer& pl = frame[pool_src];
    if (pl.left != std::max(true_left[pool_src], 0) || pl.right != true_right[pool_src])
      throw UnsupportedError("the pooled layer " + pl.name + " is computable on frames [" + std::to_string(true_left[pool_src]) +
                             ", T-1-(" + std::to_string(true_right[pool_src]) + ")] under nnet3's rules but only on [" +
                             std::to_string(pl.left) + ", T-1-" + std::to_string(pl.right) +
                             "] in a layout that holds frames 0..T-1 of every layer (some layer is needed before the first or "
                             "after the last input frame); the recipes' networks have no such taps");
  }
  int32_t d = PooledDim();
  for (Layer& l : segment) {
    if (l.in_dim != d) fail("segment layer " + l.name + ": in_dim " + std::to_string(l.in_dim) + " != " + std::to_string(d));
    if (l.weight.size() != static_cast<size_t>(l.in_dim) * l.out_dim) fail("segment layer " + l.name + ": weight size");
    if (!l.bias.empty() && static_cast<int32_t>(l.bias.size()) != l.out_dim) fail("segment layer " + l.name + ": bias size");
    if (l.scale.size() != l.offset.size() || (!l.scale.empty() && static_cast<int32_t>(l.scale.size()) != l.out_dim))
      fail("segment layer " + l.name + ": scale/offset size");
    d = l.out_dim;
  }

  std::ostringstream os;
  for (size_t i = 0; i < input_dims.size(); ++i)
    os << "input[" << i << "] " << input_names[i] << " dim=" << input_dims[i] << "\n";
  for (size_t i = 0; i < frame.size(); ++i) {
    const Layer& l = frame[i];
    os << "frame[" << i << "] " << l.name << " " << l.in_dim << "->" << l.out_dim << " taps={";
    for (size_t t = 0; t < l.taps.size(); ++t) {
      if (t) os << ",";
      if (l.taps[t].src >= 0) os << "L" << l.taps[t].src;
      else os << "in" << (-1 - l.taps[t].src);
      os << "@" << l.taps[t].offset;
    }
    os << "}" << (l.relu ? " relu" : "") << (l.scale.empty() ? "" : " bn") << " ctx=" << l.left << "/" << l.right << "\n";
  }
  if (FrameOutput())
    os << "frame output L" << frame_out << (frame_log_softmax ? " log-softmax" : "") << "\n";
  else
    os << "pool src=L" << pool_src << " window=[-" << pool_left_context << ",+" << pool_right_context << "] "
     << (pool_output_stddevs ? "mean+stddev" : "mean") << " floor=" << pool_variance_floor << "\n";
  for (size_t i = 0; i < segment.size(); ++i) {
    const Layer& l = segment[i];
    os << "segment[" << i << "] " << l.name << " " << l.in_dim << "->" << l.out_dim << (l.relu ? " relu" : "")
       << (l.scale.empty() ? "" : " bn") << "\n";
  }
  description = os.str();
}

}  // namespace cvx
