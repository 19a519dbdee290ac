#!/bin/bash
# Parity tests + short bench for each single-ray kernel variant (RTBVH_TRACE_MODE).
set -u
TAG=${1:-modes}
MODES=${2:-"persistent coop"}
OUT=gpurun_out
mkdir -p $OUT
for MODE in $MODES; do
  echo "== pytest traversal ($MODE)"
  RTBVH_TRACE_MODE=$MODE timeout 900 python -m pytest tests/test_gpu_traversal.py -x -q -m gpu 2>&1 | tail -4
  echo "== bench ($MODE)"
  RTBVH_TRACE_MODE=$MODE timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 5 2> $OUT/${TAG}_$MODE.err > $OUT/${TAG}_$MODE.json
  python - <<PY
import json
d=json.load(open("$OUT/${TAG}_$MODE.json"))
print("$MODE", round(d["value"],1), "Mrays/s  e2e", round(d["e2e"]["value"],1))
PY
done
