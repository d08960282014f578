#!/bin/bash
# same-box A/B of the drop-in program's wall time: orderly tear-down (CVX_FULL_TEARDOWN=1) vs leaving it to process exit
mkdir -p gpurun_out
python tools/cli_throughput.py 2500 > gpurun_out/r2w_cli_warm.log 2>&1   # warms the box (driver, page cache)
for i in 1 2; do
  CVX_FULL_TEARDOWN=1 python tools/cli_throughput.py 2500 > gpurun_out/r2w_cli_teardown_$i.log 2>&1
  python tools/cli_throughput.py 2500 > gpurun_out/r2w_cli_fastexit_$i.log 2>&1
done
grep "^scp (same\|^ark file\|^scp, auto" gpurun_out/r2w_cli_teardown_*.log gpurun_out/r2w_cli_fastexit_*.log | cut -c1-250
python -m pytest tests/test_gpu_jobs.py tests/test_gpu_extract.py tests/test_gpu_prestep.py tests/test_gpu_backend.py -x -q -m gpu 2>&1 | tail -2
