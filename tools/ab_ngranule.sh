#!/bin/bash
# A/B on one box: MMA width granularity 16 (default) vs 32 (CVX_N_GRANULE=32), alternating runs.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_headline.py tests/test_gpu_extract.py tests/test_gpu_frames.py -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1
tail -3 gpurun_out/r2v_tests.log
for i in 1 2 3; do
  for g in 16 32; do
    CVX_N_GRANULE=$g python bench.py --quick --steps 6 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2v_bench_g${g}_$i.json
    python - <<PY
import json
d=json.load(open("gpurun_out/r2v_bench_g${g}_$i.json"))
L=d["layer_ms_per_step"]
print("granule", $g, "run", $i, "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4), "sm", d["clocks"]["sm_mhz"],
      "tdnn2-4", [round(L[k]["ms"],2) for k in ("tdnn2.batchnorm","tdnn3.batchnorm","tdnn4.batchnorm")], "tdnn1", round(L["tdnn1.batchnorm"]["ms"],2), "parity", d.get("parity",{}).get("max_rel"))
PY
  done
done
