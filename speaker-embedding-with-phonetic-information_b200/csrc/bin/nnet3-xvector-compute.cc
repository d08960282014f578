// Drop-in for upstream Kaldi's src/nnet3bin/nnet3-xvector-compute.cc as called by
// egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:84-93.
#include "extract_main.h"
int main(int argc, char** argv) { return cvx::ExtractMain(argc, argv, false); }
