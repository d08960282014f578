// Drop-in for `nnet3-cvector-compute-multiple-input`, referenced (not shipped) by
// egs/sre/v5/sid/nnet3_cvector/cvector/extract_cvectors_with_am.sh:99-110.
#include "extract_main.h"
int main(int argc, char** argv) { return cvx::ExtractMain(argc, argv, true); }
