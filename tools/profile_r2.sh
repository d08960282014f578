#!/bin/bash
# ncu captures behind profiles/r2_*.md (run on the GPU box: gpurun -- 'bash tools/profile_r2.sh').
# tools/profile_step.py runs ONE 500 000-row device batch per call, so per call `-k regex:tdnn_gemm`
# matches exactly the 10 frame layers in plan order (the pooled layer = the kPool instance) + the
# segment GEMM, and the auxiliary kernels match once each.
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
# 1. launch list of bench.py's timed steps (shares only: cold-cache, serialised)
$NCU --metrics gpu__time_duration.sum -s 165 -c 110 --csv --log-file $O/r2_launches.csv \
    python bench.py --quick --steps 2 --warmup 3 --utts-per-gpu 120 > $O/r2_launches_bench.log 2>&1
# 2. all 10 frame layers of one batch of the headline net (4th call), full set + source
$NCU --set full --import-source on -k regex:tdnn_gemm -s 33 -c 10 -f -o $O/r2_prof_gemm \
    python tools/profile_step.py --tag r2 > $O/r2_prof_gemm.log 2>&1
# 3. the other kernels of that call
$NCU --set full -k 'regex:pool_meta|stage_input|stats_pool_finalize|split_f32|utt_average' -s 15 -c 5 -f -o $O/r2_prof_aux \
    python tools/profile_step.py --tag r2 > $O/r2_prof_aux.log 2>&1
# 4. the stand-alone HBM-bound pooling kernel
CVX_FUSED_POOL=0 $NCU --set full -k regex:stats_pool_partial -s 3 -c 1 -f -o $O/r2_prof_pool \
    python tools/profile_step.py --tag r2 > $O/r2_prof_pool.log 2>&1
# 5. short segments, x-vector net: 2 500 x 200-frame segments = one 500 000-row batch (5 frame layers + segment GEMM per call)
$NCU --set full -k regex:tdnn_gemm -s 18 -c 5 -f -o $O/r2x_prof_gemm \
    python tools/profile_step.py --topology v2_xvector --utts 2500 --frames 200 --chunk-size -1 --tag r2x > $O/r2x_prof_gemm.log 2>&1
ls -la $O/*.ncu-rep
