// Shared driver of the two drop-in programs (see extract_main.cc).
#ifndef CVX_BIN_EXTRACT_MAIN_H_
#define CVX_BIN_EXTRACT_MAIN_H_

namespace cvx {
// multiple_input = false: nnet3-xvector-compute       <nnet> <feats-rspecifier> <vector-wspecifier>
// multiple_input = true : nnet3-cvector-compute-multiple-input --name=a,b <n> <nnet> <feats1> .. <featsn> <vector-wspecifier>
int ExtractMain(int argc, const char* const* argv, bool multiple_input);
}  // namespace cvx
#endif
