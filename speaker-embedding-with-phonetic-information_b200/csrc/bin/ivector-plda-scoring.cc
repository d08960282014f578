// Drop-in for upstream Kaldi's ivector-plda-scoring as called by egs/sre/v5/run_sre10.sh:122-129
// (see vector_tools_main.cc).
#include "extract_main.h"
int main(int argc, char** argv) { return cvx::PldaScoringMain(argc, argv); }
