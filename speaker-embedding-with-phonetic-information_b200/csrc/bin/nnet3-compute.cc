// Drop-in for upstream Kaldi's src/nnet3bin/nnet3-compute.cc as called by
// egs/sre/v5/sid/nnet3_cvector/cvector/extract_am_embedding.sh:79-85 and extract_log_post.sh:80-85.
#include "extract_main.h"
int main(int argc, char** argv) { return cvx::ComputeMain(argc, argv); }
