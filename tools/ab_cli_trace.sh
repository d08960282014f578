#!/bin/bash
# where the drop-in program's wall time goes: per-call host phases (CVX_TRACE) with the old and new buffer growth policy
mkdir -p gpurun_out
CVX_TRACE=1 CVX_GROW_EXACT=1 python tools/cli_throughput.py 2500 > gpurun_out/r2w_cli_exact.log 2>&1
CVX_TRACE=1 python tools/cli_throughput.py 2500 > gpurun_out/r2w_cli_full.log 2>&1
grep -c "cvx trace" gpurun_out/r2w_cli_exact.log gpurun_out/r2w_cli_full.log
grep "audio-s/s" gpurun_out/r2w_cli_exact.log gpurun_out/r2w_cli_full.log
