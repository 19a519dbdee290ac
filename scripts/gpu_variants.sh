#!/bin/bash
# Bench every librtbvh_rs_<variant>.so (short runs, no CPU baseline) to pick kernel tuning parameters.
set -u
TAG=${1:-var}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu (default lib)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_gpu.txt
for LIB in rtbvh_b200/librtbvh_rs.so rtbvh_b200/librtbvh_rs_*.so; do
  NAME=$(basename $LIB .so)
  RTBVH_LIB=$PWD/$LIB timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 5 2> $OUT/${TAG}_$NAME.err > $OUT/${TAG}_$NAME.json
  python - <<PY
import json
d=json.load(open("$OUT/${TAG}_$NAME.json"))
print("$NAME", round(d["value"],1), "Mrays/s  e2e", round(d["e2e"]["value"],1), d["config"].get("build",{}).get("binned_sah_ms_per_mtri"))
PY
done
