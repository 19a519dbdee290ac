#!/bin/bash
# Last verification of the round on the final tree: GPU suite, smoke, default bench (both arms).
set -u
TAG=${1:-r6c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 900 python bench.py 2> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-300
tail -2 $OUT/${TAG}_bench.err | cut -c1-200
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2> $OUT/${TAG}_ref.err | tee $OUT/${TAG}_ref.json | cut -c1-200
