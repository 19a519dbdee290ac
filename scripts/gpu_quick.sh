#!/bin/bash
# Quick GPU check: full gpu test-suite + one short bench line.
set -u
TAG=${1:-quick}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_gpu.txt
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 5 2> $OUT/${TAG}_bench.err > $OUT/${TAG}_bench.json
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench.json"))
print(round(d["value"],1), "Mrays/s  e2e", round(d["e2e"]["value"],1), json.dumps(d["config"].get("build")))
PY
tail -3 $OUT/${TAG}_bench.err
