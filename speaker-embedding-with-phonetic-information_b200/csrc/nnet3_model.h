// nnet3 raw-model loader: file -> graph -> Plan.
#ifndef CVX_NNET3_MODEL_H_
#define CVX_NNET3_MODEL_H_

#include <stdexcept>
#include <string>

#include "plan.h"

namespace cvx {

struct FormatError : public std::runtime_error {
  using std::runtime_error::runtime_error;
};

// Reads `rxfilename` (path, "-", or "cmd |"; text or binary nnet3 raw model).
// If `output_node` is non-empty, output-node `output` is redefined to read that
// descriptor first (what `nnet3-copy --nnet-config=extract.config` does in
// egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:58-59).
// frame_output: lower for nnet3-compute (the output is a frame-level layer, optionally behind a
// LogSoftmaxComponent; egs/sre/v5/sid/nnet3_cvector/cvector/extract_am_embedding.sh:79-85).
// Throws KaldiError (I/O), FormatError, UnsupportedError.
Plan LoadNnet3Plan(const std::string& rxfilename, const std::string& output_node, bool frame_output = false);

}  // namespace cvx
#endif  // CVX_NNET3_MODEL_H_
