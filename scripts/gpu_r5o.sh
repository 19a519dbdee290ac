#!/bin/bash
# Round 5, session o: full GPU suite on the current code, Q3 count (rays that a conservative node format would answer differently).
set -u
TAG=${1:-r5o}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -16 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== q3 count"; timeout 900 python scripts/q3_count.py 2> $OUT/${TAG}_q3.err | tee $OUT/${TAG}_q3.json
tail -3 $OUT/${TAG}_q3.err
