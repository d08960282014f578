#!/bin/bash
# same-box A/B of the device batch size (rows per packed batch), alternating runs
mkdir -p gpurun_out
for i in 1 2; do
  for r in 524288 786432 1048576 393216; do
    python bench.py --quick --steps 6 --warmup 3 --max-batch-rows $r 2>/dev/null | tail -1 > gpurun_out/r2z_rows_${r}_$i.json
    python - <<PY
import json
d=json.load(open("gpurun_out/r2z_rows_${r}_$i.json"))
print("rows", $r, "run", $i, "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "sm", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"])
PY
  done
done
