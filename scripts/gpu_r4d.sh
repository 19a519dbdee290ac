#!/bin/bash
# Round 4, session d: full GPU test suite incl. the full-size parity tests, bench for all five configs (first driver-style run).
set -u
TAG=${1:-r4d}
OUT=gpurun_out
mkdir -p $OUT
nproc > $OUT/${TAG}_nproc.txt
echo "== pytest -m gpu"; timeout 2000 python -m pytest tests -x -q -m gpu --durations=12 2>&1 | tail -30 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== bench (default = config 2)"
timeout 900 python bench.py --steps 40 --warmup 5 2> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-400
tail -3 $OUT/${TAG}_bench.err
for C in 1 3 4 5; do
  echo "== bench --config $C"
  timeout 1200 python bench.py --config $C 2> $OUT/${TAG}_bench_c$C.err | tee $OUT/${TAG}_bench_c$C.json | cut -c1-500
  tail -3 $OUT/${TAG}_bench_c$C.err
  echo "== bench --config $C --impl reference"
  timeout 900 python bench.py --config $C --impl reference --steps 3 --warmup 1 2> $OUT/${TAG}_ref_c$C.err | tee $OUT/${TAG}_ref_c$C.json | cut -c1-300
done
