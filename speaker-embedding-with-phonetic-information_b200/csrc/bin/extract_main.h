// Shared driver of the two drop-in programs (see extract_main.cc).
#ifndef CVX_BIN_EXTRACT_MAIN_H_
#define CVX_BIN_EXTRACT_MAIN_H_

#include <string>
#include <vector>

#include "../kaldi_io.h"

namespace cvx {
// The descriptor of "output-node name=output input=<descriptor>" in an extract.config-style file.
std::string ReadOutputNodeFromConfig(const std::string& path);
// nnet3-compute <raw-nnet-in> <features-rspecifier> <matrix-wspecifier>  (compute_main.cc)
int ComputeMain(int argc, const char* const* argv);
// Back-end vector programs (vector_tools_main.cc)
int IvectorMeanMain(int argc, const char* const* argv);
int SubtractGlobalMeanMain(int argc, const char* const* argv);
int TransformVecMain(int argc, const char* const* argv);
int NormalizeLengthMain(int argc, const char* const* argv);
int PldaScoringMain(int argc, const char* const* argv);
// Kaldi object files holding one vector / one matrix (mean.vec, transform.mat)
std::vector<float> ReadVectorFile(const std::string& rxfilename);
Matrix ReadMatrixFile(const std::string& rxfilename);
// multiple_input = false: nnet3-xvector-compute       <nnet> <feats-rspecifier> <vector-wspecifier>
// multiple_input = true : nnet3-cvector-compute-multiple-input --name=a,b <n> <nnet> <feats1> .. <featsn> <vector-wspecifier>
int ExtractMain(int argc, const char* const* argv, bool multiple_input);
}  // namespace cvx
#endif
