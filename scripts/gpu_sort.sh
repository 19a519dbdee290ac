#!/bin/bash
set -u
TAG=${1:-sort}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_traversal.py -x -q -m gpu 2>&1 | tail -6
for S in 0 1; do
  CFG4_SORT=$S timeout 900 python scripts/config4.py > $OUT/${TAG}_config4_sort$S.json 2> $OUT/${TAG}_config4_sort$S.err
  python - <<PY
import json
d=json.load(open("$OUT/${TAG}_config4_sort$S.json"))
print("config4 sort=$S", round(d["value"],1), "Mrays/s", d["ms_per_step"], d["parity_sample_bit_exact"], "build ms/Mtri", round(d["build_ms_per_mtri"],2))
PY
done
