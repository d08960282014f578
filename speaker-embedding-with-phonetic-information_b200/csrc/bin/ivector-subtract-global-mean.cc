// Drop-in for upstream Kaldi's ivector-subtract-global-mean as called by egs/sre/v5/run_sre10.sh:107-129 and
// egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:102-108 (see vector_tools_main.cc).
#include "extract_main.h"
int main(int argc, char** argv) { return cvx::SubtractGlobalMeanMain(argc, argv); }
