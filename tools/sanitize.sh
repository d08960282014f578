#!/bin/bash
# compute-sanitizer passes over the GPU tests that finish quickly under instrumentation (run on a B200):
#   memcheck   out-of-bounds / misaligned accesses in every kernel of the extraction, frame and back-end paths
#   synccheck  barrier misuse
#   racecheck  shared-memory hazards (it reports the smem write of tcgen05.alloc against the read of the
#              allocated TMEM address in tmem_alloc_pair -- the instruction is warp-collective and the read
#              follows a tcgen05 fence + cluster barrier; no other site is reported)
set -x
K='golden or frames_match or plda_scoring_matches or transform_vectors or prestep_matches or fused_pooling_matches or segment_affine or sparse_and_dense or range_check or random_topology_matches_oracle[10'
for tool in memcheck synccheck racecheck; do
  compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_extract.py tests/test_gpu_frames.py \
    tests/test_gpu_backend.py tests/test_gpu_prestep.py tests/test_gpu_kernels.py tests/test_gpu_fuzz.py -q -m gpu --timeout=800 --timeout-method=thread -k "$K" 2>&1 | tail -6
done
