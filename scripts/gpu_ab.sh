#!/bin/bash
# A/B of traversal kernel variants: parity tests, then the bench in each mode (short runs).
set -u
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_gpu.txt
for MODE in static persistent; do
  echo "== bench $MODE"
  RTBVH_TRACE_MODE=$MODE timeout 600 python bench.py --steps 40 --warmup 5 2> $OUT/${TAG}_bench_$MODE.err | tee $OUT/${TAG}_bench_$MODE.json
done
echo "== ncu full (persistent)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_single -s 3 -c 1 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log | cut -c1-300
